// rg_device.cuh -- device side of the B200-native RG hot path (sm_100a).
//
// One CTA owns one Hamiltonian for a whole `converge` (RG.py:278-290): the Lie-series operands live in
// shared memory as per-mode records, the coefficient array f and the generator y live in an
// L2-resident per-CTA workspace, every loop of the reference (series, Lie steps, rg_map iterates)
// runs on the device.  Compile with -fmad=false: only the explicit fma() of the convolution loops
// fuse, the elementwise arithmetic rounds like the reference's numpy expressions.
#pragma once
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdint.h>

namespace rgdev {

struct Sched {
    int n_warps, n_boxes, n_support, half, n_stages;
    int general;                    // 1: caller-supplied generator (schedule kind 1): no product-rule split (BIG kernels)
    int maxp;                       // piece slots per warp and stage (rgplan::Schedule::maxp)
    int fin_begin[4], fin_end[4];   // boxes owned by each CTA of a cluster
    int stage0_pieces[4];           // pieces each CTA combines in the storing stage
    const int* lane_out;         // [NB*32] centred mode of each lane, -1 idle
    const uint32_t* lane_coord;  // [NB*32] packed biased coords
    const int4* pieces;          // [n_stages][NW*maxp] {box, begin, end, rank}; box = -1 unused; heads first
    const int2* piece_runs;      // same shape: {end of the piece's run of complete products, end of its run truncated to LIM_B}
    const uint2* entries;        // {lanes of the box in range for nu1 (bit l = lane l), centred index of nu1 | rec<<16}
    const int* supp_rec;         // [M] record of centred mode or -1
};

struct DevPlan {
    int dim, L, J, JP, n1, M, vecjl, off2, rec_units, idx_y1J;
    int g_units;               // BIG kernels: 16-byte units of a g record (g[0..J] only, product-rule form); else = rec_units
    int g_smem;                // BIG kernels: 1 = the g records live in shared memory (and `out` in the global workspace), 0 = in the workspace
    int stride0, stride1, stride2;
    int norm_choice;
    double norm_analytic;
    double tol_min, tol_max;
    int max_lie, ct;
    double min_step, abs_tol, rel_tol, eigenvalue;
    int max_series;
    long long max_liouville;   // work budget per Hamiltonian (the reference has none, quirk Q6)
    int sym_tables;            // 1: mask[-nu] == mask[nu], omega0_nu[-nu] == -omega0_nu[nu] exactly (checked by the host)
    int rec_global;            // 1: the operand records live in the CTA's global workspace (array too large for shared memory)
    long long ws_stride;       // doubles of workspace per CTA (WS_COUNT arrays + the global records, if any); even
    long long rec_offset;      // doubles from a CTA's workspace base to its global records (16-byte aligned): Y records first
    long long rec2_offset;     // BIG kernels: doubles to the second global block -- the g records (g_smem = 0) or `out` (g_smem = 1)
    int debug;   // RG_B200_DEBUG experiments (0 in production): bit0 gather always hits the zero record,
                 // bit1 every entry broadcasts y record 0
    const double* omega0_nu;  // [M] centred
    const uint16_t* mask;     // [M] I^- bits per Taylor level
    const int* reidx;         // [M] centred -> centred source, -1 masked
    const double* normw;      // [M] exp(NormAnalytic*|nu|_1)
    const int* c2w;           // [M] centred -> wrapped index
    // wrapped-order tables of the stand-alone HBM-bound operators (rg_norm, rg_sym, rg_reindex_rescale)
    const uint16_t* mask_w;   // [M] I^- bits per Taylor level
    const double* normw_w;    // [M] exp(NormAnalytic*|nu|_1)
    const int* neg_w;         // [M] wrapped index of -nu
    const int* reidx_w;       // [M] wrapped index of sign*N*nu, -1 masked
    const double* theta;
    int n_theta;
    // schedule kind: 0 = I^- support, all modes; 1 = dense (caller-supplied y); 2 = I^- support, half space only
    // (exactly symmetric operands: out[-nu] = out[nu]).  One set per cluster size 1, 2, 4: index kind + 3*log2(csize).
    Sched sched[9];
};

// workspace arrays of one CTA (each vecjl doubles, layout [j][centred nu])
enum { WS_FA = 0, WS_FB, WS_Y1, WS_Y2, WS_Y3, WS_H1, WS_H2, WS_T0, WS_T1, WS_T2, WS_T3, WS_T4, WS_T5, WS_T6, WS_T7, WS_T8,
       WS_A0, WS_A1, WS_AM,   // RG.approach on the device: the bracketing pair and the midpoint
       WS_COUNT };

// Plans live in constant memory (uniform-datapath loads inside the device functions); a launch names
// its plan by slot.  Shared-memory pointers are always re-derived from the __shared__ symbol so that
// the compiler emits LDS/STS, never generic loads.
#define RG_MAX_PLANS 16
// One plan table per translation unit (= per instantiated kernel), under a name unique to the unit: with internal
// linkage (`static`) ptxas allocates the hot loop of conv_all differently (82 instead of 90 operand-reuse hits per
// complete Taylor product at J = 10, +2.5 % kernel time).
#ifndef RG_PLAN_TABLE
#define RG_PLAN_TABLE c_plan_host
#endif
#define c_plan RG_PLAN_TABLE
__constant__ DevPlan c_plan[RG_MAX_PLANS];
extern __shared__ double2 smem_dyn[];

struct Ctx {
    int slot;     // index into c_plan
    int sidx;     // schedule index (0 sparse, 1 dense, 2 sparse half-space)
    int symm;     // 1 while the coefficient array is exactly symmetric under nu -> -nu (after RG.sym)
    int cbase;    // 3*log2(cluster size): offset of the schedule set
    int csize, crank;  // thread-block cluster size (1, 2, 4) and this CTA's rank in it
    double* ws;   // this CTA's global workspace
    double2* rec; // global operand records (BIG kernels): the Y records
    double* rec2; // BIG kernels: the g records (g_smem = 0) or `out` (g_smem = 1) in the global workspace
    const double* Omega_nu;  // current Omega.nu, centred, global
    long long n_liou, n_lie;
    long long n_liou_full;   // of n_liou: brackets evaluated on the full (not half-space) schedule: the 1-norm estimator's probes
    uint64_t rng;  // state of the 1-norm estimator's +-1 columns (seeded per Hamiltonian)
    long long t_conv, t_epi;  // SM clocks spent in conv_all / series_epilogue (thread 0's view; profiling aid)
    long long t_busy;         // of t_conv: clocks warp 0 spent computing its own pieces (the rest is barrier wait)
    long long budget_base;   // n_liou when the current converge started: max_liouville is a budget per converge (RG.approach runs many)
};

// Shared memory of a CTA: [red: 64 doubles of reduction scratch + 16 bytes for the work ticket | G | Y | out].
// BIG kernels (arrays too large for one SM, e.g. n = 3 at L = J = 6: 246 KB of g records alone) keep G and Y in
// the CTA's global workspace instead -- private to the CTA, read through L1 -- and only `red` and `out` on chip.
struct Smem {
    double2* G;   // M records followed by RG_ZERO_RECS all-zero ones (out-of-range gathers)
    double2* Y;   // n_support records
    double* out;  // [M][JP]
};
#define RG_RED_UNITS 33   // 16-byte units of the `red` block
#define RG_CTL_UNITS 128  // 16-byte units of the convolution control block: 512 ints (per-box rank counters + 1)
#define RG_MAX_BOXES 500
// An out-of-range lane gathers the zero record whose index is congruent (mod 8) to the record it would have
// read: a quarter-warp's eight addresses then stay in eight distinct 16-byte bank groups whatever the mix of
// in-range and out-of-range lanes (with a single zero record 30 % of the shared-memory wavefronts were replays).
#define RG_ZERO_RECS 8
__device__ __forceinline__ double* smem_red() { return reinterpret_cast<double*>(smem_dyn); }
__device__ __forceinline__ int* smem_ctl() { return reinterpret_cast<int*>(smem_dyn + RG_RED_UNITS); }
// BIG kernels (Hamiltonians whose two-derivative records do not fit one SM: n = 3 at L = J = 6, 2-D at L >= 14) use
// the product-rule form of the bracket, whose g records hold g[0..J] only (5 instead of 7 units at J = 6): they go to
// shared memory when they fit (176 KB at n = 3, L = J = 6) with `out` in the CTA's global workspace, else they stay
// in the workspace and `out` is on chip.  The Y records (warp-uniform loads) are always in the workspace.  All three
// are reached through generic pointers in these kernels.
template <bool BIG>
__device__ __forceinline__ Smem smem_view(const DevPlan& P, const Sched& S, const Ctx& c) {
    Smem v;
    if (BIG) {
        v.Y = c.rec;
        if (P.g_smem) {
            v.G = smem_dyn + RG_RED_UNITS + RG_CTL_UNITS;
            v.out = c.rec2;
        } else {
            v.G = reinterpret_cast<double2*>(c.rec2);
            v.out = reinterpret_cast<double*>(smem_dyn + RG_RED_UNITS + RG_CTL_UNITS);
        }
    } else {
        v.G = smem_dyn + RG_RED_UNITS + RG_CTL_UNITS;
        v.Y = v.G + (P.M + RG_ZERO_RECS) * P.rec_units;
        v.out = reinterpret_cast<double*>(v.Y + S.n_support * P.rec_units);
    }
    return v;
}
#define RG_PLAN(c) c_plan[(c).slot]
#define RG_SCHED(c) c_plan[(c).slot].sched[(c).sidx + (c).cbase]

__device__ __forceinline__ double pmax(double a, double b) { return (a > b || a != a) ? a : b; }

__device__ __forceinline__ void norm_acc(int type, double& p, double x, double w) {
    double ax = fabs(x);
    if (type == 0) p += ax;
    else if (type == 1) p = pmax(p, ax);
    else if (type == 2) p += ax * ax;
    else p = pmax(p, ax * w);
}
__device__ __forceinline__ double norm_fin(int type, double p) { return type == 2 ? sqrt(p) : p; }

// deterministic two-value block reduction (fixed shuffle tree, fixed cross-warp order)
__device__ __forceinline__ double2 block_reduce2(double* red, int type, double a, double b) {
    const bool ismax = (type == 1 || type == 3);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double a2 = __shfl_down_sync(0xffffffffu, a, o), b2 = __shfl_down_sync(0xffffffffu, b, o);
        if (ismax) { a = pmax(a, a2); b = pmax(b, b2); } else { a += a2; b += b2; }
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (lane == 0) { red[w] = a; red[32 + w] = b; }
    __syncthreads();
    double ra = red[0], rb = red[32];
    for (int i = 1; i < nw; ++i) {
        if (ismax) { ra = pmax(ra, red[i]); rb = pmax(rb, red[32 + i]); } else { ra += red[i]; rb += red[32 + i]; }
    }
    __syncthreads();
    return make_double2(norm_fin(type, ra), norm_fin(type, rb));
}

// ------------------------------------------------------------------------------------------------
// Shared-memory kernels (BIG = false).  Taylor product of one (nu1, nu2) pair: acc[j] += sum_{j1+j2=j} dAy[j1]*Omg[j2] - Omy[j1]*dAg[j2]
// record layout (doubles): [0..J-1] d-type, [OFF2..OFF2+J] Omega-type, y1[J] (caller-supplied y only) at IDXJ.
//
// Everything the inner loop reads from shared memory is in flight long before it is consumed:
//   * the two halves of the gathered record G[nu-nu1] live in separate registers; the d-half is loaded
//     while the Omega-half is consumed (first convolution) and the next entry's Omega-half while the
//     d-half is consumed (second convolution);
//   * the broadcast y values go through a rolling window of YW double2 registers that runs across entry
//     boundaries: pair k+YW is requested the moment pair k is taken out of the window.
template <int J>
struct RecGeom {
    static constexpr int JP = J + 1, OFF2 = J + (J & 1), TOT = OFF2 + JP, NU2 = (TOT + 1) / 2;
    static constexpr int ND2 = OFF2 / 2;      // double2 units of the d-type half
    static constexpr int NO2 = NU2 - ND2;     // double2 units of the Omega-type half
    static constexpr int IDXJ = (J & 1) ? J : OFF2 + JP;
    static constexpr int NP1 = (J + 1) / 2;   // y pairs of conv(dA y, Omega g): y1[0..J-1]
    static constexpr int NP2 = J / 2 + 1;     // y pairs of conv(Omega y, dA g): y2[0..J]
    static constexpr int NP = NP1 + NP2;
    // prefetch window of the broadcast y values (pairs).  Which window is best is a property of the register allocation
    // ptxas settles on for the whole kernel, not of the loop: with 12 warps x 168 registers and the coordinate
    // test in the decode 1 pair won (50.0 ms on the uniform batch against 50.4 / 50.8 / 52.7 ms with 2 / 3 / 4), with the
    // shipped 8 warps x 255 registers and the planner's range masks 3 pairs do (57.4 ms against 60.0 with 1, 58.4 with 2):
    // the candidates were ranked on the CPU by the measured DFMA cost model (2 cycles with an adjacent operand-reuse hit,
    // 3 without, +1 per integer instruction; tools/dfma_lab) and the best three confirmed on the GPU.
#ifndef RG_YW
#define RG_YW ((J >= 8) ? 3 : 1)
#endif
    static constexpr int YW = RG_YW;
    // Every entry runs a branch-free, software-pipelined Taylor product on its zero-padded y record (the absent
    // levels are stored as exact zeros): entries with jm <= LIM_A levels take the product truncated to LIM_A
    // levels of y, jm <= LIM_B the one truncated to LIM_B, the rest the complete one (J = 10: 3 / 6 / 11 levels,
    // J = 6: 3 / 5 / 7).  rgplan::level_class mirrors this.
    static constexpr int LIM_A = (J >= 8) ? 3 : (J / 2 > 2 ? J / 2 : 2), LIM_B = (J >= 8) ? 6 : J - 1;
    static constexpr int NPAD = ((NP + YW - 1) / YW) * YW;   // the window must line up across entries
    // double2 unit of the k-th y pair in consumption order (indices >= NP are padding: re-read the last pair)
    __host__ __device__ static constexpr int unit(int k) { return k < NP1 ? k : (k < NP ? ND2 + (k - NP1) : ND2 + NP2 - 1); }
};

template <int J>
__device__ __forceinline__ void load_omega_half(double (&gO)[2 * RecGeom<J>::NO2], const double2* __restrict__ gp) {
#pragma unroll
    for (int u = 0; u < RecGeom<J>::NO2; ++u) { const double2 t = gp[RecGeom<J>::ND2 + u]; gO[2 * u] = t.x; gO[2 * u + 1] = t.y; }
}
template <int J>
__device__ __forceinline__ void load_d_half(double (&gD)[2 * RecGeom<J>::ND2], const double2* __restrict__ gp) {
#pragma unroll
    for (int u = 0; u < RecGeom<J>::ND2; ++u) { const double2 t = gp[u]; gD[2 * u] = t.x; gD[2 * u + 1] = t.y; }
}

// Per-lane constants of the gather: the lane's bit in an entry's range mask, the linear index of its output mode.
struct LaneGeo {
    uint32_t lanebit;     // 1 << lane
    int lin;              // sum_d v_d * stride_d : centred index of (nu - nu1) = lin - index(nu1)
    int zero_rec;         // index of the first all-zero record (M)
    uint32_t g_base, y_base;  // shared-memory byte addresses of G and Y
    int rec_bytes;
};

// next entry: form the two record addresses.  w = the entry's two list words, read by every lane from the same address
// two entries ahead (a warp-wide block of words broadcast by SHFL put the shuffle's latency on the critical path of the
// short, level-truncated products: 6 % of the stall samples).  Whether nu - nu1 is on the grid is the planner's test:
// word 0 carries one bit per lane of the box (an integer instruction costs the fp64 stream an issue cycle and the DFMA
// behind it its operand-reuse hit, tools/dfma_lab: the three-coordinate test in the kernel was 9 of them per entry).
__device__ __forceinline__ void decode_entry(const LaneGeo& q, const uint2& w, uint32_t& gp, uint32_t& yp) {
    const uint32_t ei = w.y;
    const bool ok = (w.x & q.lanebit) != 0;
    const int nat = q.lin - (int)(ei & 0xffff);
    const int gi = ok ? nat : q.zero_rec + ((nat - q.zero_rec) & (RG_ZERO_RECS - 1));
    gp = q.g_base + (uint32_t)(gi * q.rec_bytes);
    yp = q.y_base + (uint32_t)((int)(ei >> 16) * q.rec_bytes);
}

__device__ __forceinline__ double2 ldrec(const LaneGeo& q, uint32_t addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}

// One list entry with every Taylor level of y present (jm == J+1): branch-free; the decode of the next entry
// sits inside the DFMA stream (same basic block) so that its latency chain is hidden.
// yq: window holding pairs 0..YW-1 of this entry on entry, pairs 0..YW-1 of the next entry on exit.
// gO: Omega-half of this entry's gathered record on entry, of the next entry's on exit.
// LIM: Taylor levels of y the entry may have (its record is zero beyond its own jm <= LIM): LIM = J+1 is the
// complete product; smaller LIM drop the blocks of the absent levels at compile time and keep the load schedule.
template <int J, int LIM>
__device__ __forceinline__ void taylor_full(double (&acc)[J + 1], double (&gO)[2 * RecGeom<J>::NO2],
                                            double2 (&yq)[RecGeom<J>::YW], const LaneGeo& q, const uint2& wn,
                                            uint32_t& gp, uint32_t& yp) {
    typedef RecGeom<J> R;
    double gD[2 * R::ND2];
#pragma unroll
    for (int u = 0; u < R::ND2; ++u) { const double2 t = ldrec(q, gp + 16 * u); gD[2 * u] = t.x; gD[2 * u + 1] = t.y; }
    uint32_t gp_n = gp, yp_n = yp;
#pragma unroll
    for (int k = 0; k < R::NPAD; ++k) {
        const double2 t = yq[k % R::YW];
        if (k == (R::NP1 > 1 ? 1 : 0)) decode_entry(q, wn, gp_n, yp_n);   // after the first block has issued; before the next record is requested (J = 2: NP1 = 1)
        if (k + R::YW < R::NPAD) yq[k % R::YW] = ldrec(q, yp + 16 * R::unit(k + R::YW));
        else yq[k % R::YW] = ldrec(q, yp_n + 16 * R::unit(k + R::YW - R::NPAD));
        if (k >= R::NP) continue;   // padding pair
        if (k < R::NP1) {
            // conv(dA y, Omega g): y1[j1] times B1[j2] = gO[j2]
            const int j1 = 2 * k;
            if (j1 + 1 < LIM) {     // y1[j1] = (j1+1) y[j1+1]
#pragma unroll
                for (int j2 = 0; j2 <= J - j1; ++j2) acc[j1 + j2] = fma(t.x, gO[j2], acc[j1 + j2]);
            }
            if (j1 + 1 < J && j1 + 2 < LIM) {
#pragma unroll
                for (int j2 = 0; j2 <= J - j1 - 1; ++j2) acc[j1 + 1 + j2] = fma(t.y, gO[j2], acc[j1 + 1 + j2]);
            }
            if (k == R::NP1 - 1) {   // next entry's Omega-half lands while the d-half is consumed
#pragma unroll
                for (int u = 0; u < R::NO2; ++u) { const double2 g2 = ldrec(q, gp_n + 16 * (R::ND2 + u)); gO[2 * u] = g2.x; gO[2 * u + 1] = g2.y; }
            }
        } else {
            // conv(Omega y, dA g): y2[j1] times B2[j2] = gD[j2], j2 < J ; subtracted
            const int j1 = 2 * (k - R::NP1);
            if (j1 < LIM) {
#pragma unroll
                for (int j2 = 0; j2 < J && j1 + j2 <= J; ++j2) acc[j1 + j2] = fma(-t.x, gD[j2], acc[j1 + j2]);
            }
            if (j1 + 1 <= J && j1 + 1 < LIM) {
#pragma unroll
                for (int j2 = 0; j2 < J && j1 + 1 + j2 <= J; ++j2) acc[j1 + 1 + j2] = fma(-t.y, gD[j2], acc[j1 + 1 + j2]);
            }
        }
    }
    gp = gp_n; yp = yp_n;
}

// One list entry with only the first jm Taylor levels of y present (or, jm = J+2, a caller-supplied generator):
// blocks run in ascending j1 and stop at the first absent level; y values are read on demand.  Keeps the
// hand-over protocol of taylor_full (window and gO hold the next entry's data on exit).
template <int J>
__device__ __forceinline__ void taylor_partial(double (&acc)[J + 1], double (&gO)[2 * RecGeom<J>::NO2],
                                               double2 (&yq)[RecGeom<J>::YW], const LaneGeo& q, const uint2& wn,
                                               uint32_t& gp, uint32_t& yp, const int jm) {
    typedef RecGeom<J> R;
    constexpr int JP = J + 1;
    double gD[2 * R::ND2];
#pragma unroll
    for (int u = 0; u < R::ND2; ++u) { const double2 t = ldrec(q, gp + 16 * u); gD[2 * u] = t.x; gD[2 * u + 1] = t.y; }
    uint32_t gp_n = gp, yp_n = yp;
    decode_entry(q, wn, gp_n, yp_n);
    // conv(dA y, Omega g): y1[j1] = (j1+1) y[j1+1] is non-zero iff j1+1 < jm
#pragma unroll
    for (int k = 0; k < R::NP1; ++k) {
        const int j1 = 2 * k;
        if (!(j1 + 1 < jm)) break;
        const double2 t = (k < R::YW) ? yq[k] : ldrec(q, yp + 16 * R::unit(k));
#pragma unroll
        for (int j2 = 0; j2 <= J - j1; ++j2) acc[j1 + j2] = fma(t.x, gO[j2], acc[j1 + j2]);
        if (j1 + 1 < J && j1 + 2 < jm) {
#pragma unroll
            for (int j2 = 0; j2 <= J - j1 - 1; ++j2) acc[j1 + 1 + j2] = fma(t.y, gO[j2], acc[j1 + 1 + j2]);
        }
    }
    if (JP < jm) {  // caller-supplied y with y1[J] != 0 (never on the fused path)
        double a;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(a) : "r"(yp + 8 * R::IDXJ));
        acc[J] = fma(a, gO[0], acc[J]);
    }
#pragma unroll
    for (int u = 0; u < R::NO2; ++u) { const double2 g2 = ldrec(q, gp_n + 16 * (R::ND2 + u)); gO[2 * u] = g2.x; gO[2 * u + 1] = g2.y; }
    // conv(Omega y, dA g): y2[j1] non-zero iff j1 < jm ; subtracted
#pragma unroll
    for (int k = 0; k < R::NP2; ++k) {
        const int j1 = 2 * k;
        if (!(j1 < jm)) break;
        const double2 t = (R::NP1 + k < R::YW) ? yq[R::NP1 + k] : ldrec(q, yp + 16 * R::unit(R::NP1 + k));
#pragma unroll
        for (int j2 = 0; j2 < J && j1 + j2 <= J; ++j2) acc[j1 + j2] = fma(-t.x, gD[j2], acc[j1 + j2]);
        if (j1 + 1 <= J && j1 + 1 < jm) {
#pragma unroll
            for (int j2 = 0; j2 < J && j1 + 1 + j2 <= J; ++j2) acc[j1 + 1 + j2] = fma(-t.y, gD[j2], acc[j1 + 1 + j2]);
        }
    }
#pragma unroll
    for (int k = 0; k < R::YW; ++k) yq[k] = ldrec(q, yp_n + 16 * R::unit(k));
    gp = gp_n; yp = yp_n;
}

// One piece of a box's source list.  The planner sorts a list by level class, so a piece is at most three runs (ends
// e_full <= e_mid <= e_end: complete products, products truncated to LIM_B levels of y, to LIM_A): one counted loop per
// class.  A caller-supplied generator (schedule kind 1) takes the on-demand product for every entry.
template <int J>
__device__ __forceinline__ void conv_piece(const Ctx& c, int box, int e_begin, int e_full, int e_mid, int e_end, double (&acc)[J + 1]) {
    typedef RecGeom<J> R;
    const DevPlan& P = RG_PLAN(c);
    const Sched& S = RG_SCHED(c);
    const Smem sm = smem_view<false>(P, S, c);
    const uint2* const entries = S.entries;
    const int lane = threadIdx.x & 31;
    const uint32_t mc = S.lane_coord[box * 32 + lane];
    LaneGeo q;
    q.lanebit = 1u << lane;
    q.lin = ((int)(mc & 0xff) + P.L) * P.stride0 + ((int)((mc >> 8) & 0xff) + P.L) * P.stride1 + ((int)((mc >> 16) & 0xff) + P.L) * P.stride2;
    q.zero_rec = P.M;
    q.rec_bytes = P.rec_units * 16;
    q.g_base = (uint32_t)__cvta_generic_to_shared(sm.G);
    q.y_base = (uint32_t)__cvta_generic_to_shared(sm.Y);
    double gO[2 * R::NO2];
    double2 yq[R::YW];
#pragma unroll
    for (int j = 0; j <= J; ++j) acc[j] = 0.0;
    if (e_begin >= e_end) return;
    // list words of the current entry, the next one (decoded inside the current product) and the one after
    // (in flight); past the end the last entry is read again (harmless, keeps the loop uniform)
    const int last = e_end - 1;
    uint2 wn = entries[min(e_begin + 1, last)];
    uint32_t gp, yp;
    decode_entry(q, entries[e_begin], gp, yp);
#pragma unroll
    for (int u = 0; u < R::NO2; ++u) { const double2 t = ldrec(q, gp + 16 * (R::ND2 + u)); gO[2 * u] = t.x; gO[2 * u + 1] = t.y; }
#pragma unroll
    for (int k = 0; k < R::YW; ++k) yq[k] = ldrec(q, yp + 16 * R::unit(k));
    int i = e_begin;
    if (S.general) {
#pragma unroll 1
        for (; i < e_end; ++i) {
            const uint2 wnn = entries[min(i + 2, last)];
            taylor_partial<J>(acc, gO, yq, q, wn, gp, yp, J + 2);
            wn = wnn;
        }
        return;
    }
#pragma unroll 1
    for (; i < e_full; ++i) {
        const uint2 wnn = entries[min(i + 2, last)];
        taylor_full<J, J + 1>(acc, gO, yq, q, wn, gp, yp);
        wn = wnn;
    }
#pragma unroll 1
    for (; i < e_mid; ++i) {
        const uint2 wnn = entries[min(i + 2, last)];
        taylor_full<J, R::LIM_B>(acc, gO, yq, q, wn, gp, yp);
        wn = wnn;
    }
#pragma unroll 1
    for (; i < e_end; ++i) {
        const uint2 wnn = entries[min(i + 2, last)];
        taylor_full<J, R::LIM_A>(acc, gO, yq, q, wn, gp, yp);
        wn = wnn;
    }
}

// ------------------------------------------------------------------------------------------------
// BIG kernels: the bracket in product-rule form.  RG.liouville (RG.py:115-117) convolves dA y with Omega_nu g and Omega_nu y with
// dA g; a lane would have to gather BOTH derivatives of g[nu - nu1] (2J+1 doubles per source mode).  Because
// Omega.(nu - nu1) = Omega.nu - Omega.nu1 and (a + b) y[a] g[b] = (a y[a]) g[b] + y[a] (b g[b]), the same bracket is
//
//     out[nu][j] = (Omega.nu) S1[j] - (j + 1) S2[j],
//     S1[j] = sum_{nu1} sum_{a+b=j+1, a>=1}  y1[nu1][a-1] g[nu-nu1][b]       y1[a-1] = a y[a]           (= dA y)
//     S2[j] = sum_{nu1} sum_{a+b=j+1}        y2[nu1][a]   g[nu-nu1][b]       y2[a]   = (Omega.nu1) y[a] (= Omega_nu y)
//
// so the gathered record is g itself (J+1 doubles, 6 LDS.128 at J = 10 instead of 11), the y side is what
// generate_y already produces, and an entry costs 140 DFMAs instead of 130 with 2(J+1) accumulators per lane.
// (Holds for a generator whose y1 and y2 come from ONE y, as generate_y's do; a caller-supplied pair (y1, y2) --
// RG.liouville / expm_* through the drop-in API, schedule kind 1 -- takes taylor_general below, which forms
// Omega_nu g and dA g in registers and multiplies exactly like the reference.)
//
// record layouts (doubles):  Y record: [0..J-1] y1 = dA y, [OFF2..OFF2+J] y2 = Omega_nu y, y1[J] (caller-supplied y
// only) at IDXJ;  G record: g[0..J].  Both padded to an odd number of 16-byte units (conflict-free LDS.128 over a
// quarter-warp whose record indices are distinct mod 8).
template <int J>
struct RecGeomPR {
    static constexpr int JP = J + 1, OFF2 = J + (J & 1), TOT = OFF2 + JP;
    static constexpr int ND2 = OFF2 / 2;              // double2 units of the y1 part of a Y record
    static constexpr int NY2 = (TOT + 1) / 2 - ND2;   // double2 units of the y2 part
    static constexpr int GU = (JP + 1) / 2;           // double2 units of a G record that hold data
    static constexpr int IDXJ = (J & 1) ? J : OFF2 + JP;
    // Every entry runs a branch-free, software-pipelined product on its zero-padded y record (the absent levels are
    // stored as exact zeros): entries with jm <= LIM_A levels take the product truncated to LIM_A levels of y,
    // jm <= LIM_B the one truncated to LIM_B, the rest the complete one (J = 10: 3 / 6 / 11 levels, J = 6: 3 / 5 / 7).
    // rgplan::level_class mirrors this.
    static constexpr int LIM_A = (J >= 8) ? 3 : (J / 2 > 2 ? J / 2 : 2), LIM_B = (J >= 8) ? 6 : J - 1;
    // y pairs (double2 units) a product truncated to LIM levels of y reads: y1[i] != 0 iff i + 1 < LIM, y2[a] iff a < LIM
    __host__ __device__ static constexpr int np1(int LIM) { return LIM / 2 < (J + 1) / 2 ? LIM / 2 : (J + 1) / 2; }
    __host__ __device__ static constexpr int np2(int LIM) { return (LIM + 1) / 2 < J / 2 + 1 ? (LIM + 1) / 2 : J / 2 + 1; }
    // double2 unit of the k-th y pair of such a product, in consumption order; past the end: the last pair again
    __host__ __device__ static constexpr int unit(int LIM, int k) {
        return k < np1(LIM) ? k : (k < np1(LIM) + np2(LIM) ? ND2 + (k - np1(LIM)) : ND2 + np2(LIM) - 1);
    }
    // The y pairs come from the CTA's global workspace (through L1, L2 behind it).  Up to J = 8 the registers allow
    // ALL pairs of the NEXT entry to be requested at the top of the current one (a whole product of latency cover:
    // ncu showed 4.7 long-scoreboard stalls per issue with a window of one pair); beyond that a rolling window of one.
    static constexpr int N1MAX = (J + 1) / 2, N2MAX = J / 2 + 1, NPMAX = N1MAX + N2MAX;
    static constexpr bool YALL = (J <= 8);
    static constexpr int YREGS = YALL ? NPMAX : 1;
};

// Per-lane constants of the gather: the lane's bit in an entry's range mask, the linear index of its output mode.
struct LaneGeoPR {
    uint32_t lanebit;     // 1 << lane
    int lin;              // sum_d v_d * stride_d : centred index of (nu - nu1) = lin - index(nu1)
    int zero_rec;         // index of the first all-zero record (M)
    int g_bytes, y_bytes;     // record strides
    const char* gg;           // the g records: a generic pointer (shared memory when they fit, else the CTA's workspace)
    const char* gy;           // the Y records: the CTA's global workspace
};

// next entry: form the two record addresses (word 0 of an entry: the lanes whose nu - nu1 is on the grid; decode_entry).
// An out-of-range lane gathers the zero record congruent (mod 8) to the record it would have read.
__device__ __forceinline__ void decode_entry_pr(const LaneGeoPR& q, const uint2& w, uint32_t& gp, uint32_t& yp, int& nat_out) {
    const uint32_t ei = w.y;
    const bool ok = (w.x & q.lanebit) != 0;
    const int nat = q.lin - (int)(ei & 0xffff);
    const int gi = ok ? nat : q.zero_rec + ((nat - q.zero_rec) & (RG_ZERO_RECS - 1));
    gp = (uint32_t)(gi * q.g_bytes);
    yp = (uint32_t)((int)(ei >> 16) * q.y_bytes);
    nat_out = ok ? nat : -1;
}

__device__ __forceinline__ double2 ldg_pr(const LaneGeoPR& q, uint32_t off) {     // gathered g record: generic
    double2 v;
    asm volatile("ld.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(q.gg + off));
    return v;
}
__device__ __forceinline__ double2 ldy_pr(const LaneGeoPR& q, uint32_t off) {     // warp-uniform y pair: global, through L1
    double2 v;
    asm volatile("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(q.gy + off));
    return v;
}

// One list entry of the product-rule bracket, truncated at compile time to LIM levels of y (LIM = J+1: all of them).
// On entry:  g  = this entry's gathered record, yq = its first y pair.
// On exit:   gn = the next entry's gathered record (requested right after the decode that sits inside the first rows
//            of the DFMA stream, so it has the whole product to land), yq = the next entry's first y pair,
//            (gp, yp) = the next entry's addresses.  The caller alternates (g, gn).
template <int J, int LIM>
__device__ __forceinline__ void taylor_leib(double (&S1)[J + 1], double (&S2)[J + 1], const double (&g)[2 * RecGeomPR<J>::GU],
                                            double (&gn)[2 * RecGeomPR<J>::GU], double2 (&y)[RecGeomPR<J>::YREGS],
                                            double2 (&yn)[RecGeomPR<J>::YREGS], const LaneGeoPR& q, const uint2& wn,
                                            uint32_t& gp, uint32_t& yp) {
    typedef RecGeomPR<J> R;
    constexpr int N1 = R::np1(LIM), N2 = R::np2(LIM), NP = N1 + N2;
    uint32_t gp_n = gp, yp_n = yp;
#pragma unroll
    for (int k = 0; k < NP; ++k) {
        // YALL: pair slots are fixed whatever the class -- y1 pair p in slot p, y2 pair p in slot N1MAX + p -- because the
        // next entry (requested with THIS entry's class, a superset) may run a lower one
        const double2 t = R::YALL ? y[k < N1 ? k : R::N1MAX + (k - N1)] : y[0];
        if (k == 0) {   // decode the next entry and request its record (and, registers permitting, all its y pairs)
            int nat;
            decode_entry_pr(q, wn, gp_n, yp_n, nat);
#pragma unroll
            for (int u = 0; u < R::GU; ++u) { const double2 v = ldg_pr(q, gp_n + 16 * u); gn[2 * u] = v.x; gn[2 * u + 1] = v.y; }
            if (R::YALL) {
#pragma unroll
                for (int kk = 0; kk < N1; ++kk) yn[kk] = ldy_pr(q, yp_n + 16 * kk);
#pragma unroll
                for (int kk = 0; kk < N2; ++kk) yn[R::N1MAX + kk] = ldy_pr(q, yp_n + 16 * (R::ND2 + kk));
            }
        }
        if (!R::YALL) y[0] = (k + 1 < NP) ? ldy_pr(q, yp + 16 * R::unit(LIM, k + 1)) : ldy_pr(q, yp_n + 16 * R::unit(LIM, 0));
        if (k < N1) {
            // S1[i + b] += y1[i] g[b],  i = a - 1
            const int i = 2 * k;
            if (i + 1 < LIM) {
#pragma unroll
                for (int b = 0; b <= J - i; ++b) S1[i + b] = fma(t.x, g[b], S1[i + b]);
            }
            if (i + 1 < J && i + 2 < LIM) {
#pragma unroll
                for (int b = 0; b <= J - i - 1; ++b) S1[i + 1 + b] = fma(t.y, g[b], S1[i + 1 + b]);
            }
        } else {
            // S2[a + b - 1] += y2[a] g[b]
            const int a = 2 * (k - N1);
            if (a < LIM) {
#pragma unroll
                for (int b = (a == 0 ? 1 : 0); b <= J && a + b - 1 <= J; ++b) S2[a + b - 1] = fma(t.x, g[b], S2[a + b - 1]);
            }
            if (a + 1 <= J && a + 1 < LIM) {
#pragma unroll
                for (int b = 0; b <= J && a + b <= J; ++b) S2[a + b] = fma(t.y, g[b], S2[a + b]);
            }
        }
    }
    gp = gp_n; yp = yp_n;
}

// One list entry of a caller-supplied generator (y1 and y2 are independent arrays: RG.liouville / expm_* through the
// drop-in API).  acc[j] += sum_{j1+j2=j} y1[j1] (Omega_nu2 g[j2]) - y2[j1] ((j2+1) g[j2+1]) with the two derivatives of
// g formed in registers, rounded exactly like RG.derivs (RG.py:76).  Not on the fused path; loads on demand.
template <int J>
__device__ __forceinline__ void taylor_general(double (&acc)[J + 1], const LaneGeoPR& q, const uint2& w, const double* __restrict__ Omega_nu) {
    typedef RecGeomPR<J> R;
    uint32_t gp, yp;
    int nat;
    decode_entry_pr(q, w, gp, yp, nat);
    double g[2 * R::GU];
#pragma unroll
    for (int u = 0; u < R::GU; ++u) { const double2 v = ldg_pr(q, gp + 16 * u); g[2 * u] = v.x; g[2 * u + 1] = v.y; }
    const double om2 = nat >= 0 ? Omega_nu[nat] : 0.0;
    double gO[J + 1], gD[J + 1];
#pragma unroll
    for (int j = 0; j <= J; ++j) { gO[j] = om2 * g[j]; gD[j] = (j < J) ? g[j + 1] * (double)(j + 1) : 0.0; }
#pragma unroll
    for (int k = 0; k < R::ND2; ++k) {       // conv(y1, Omega g)
        const double2 t = ldy_pr(q, yp + 16 * k);
        const int j1 = 2 * k;
#pragma unroll
        for (int j2 = 0; j2 <= J - j1; ++j2) acc[j1 + j2] = fma(t.x, gO[j2], acc[j1 + j2]);
        if (j1 + 1 < J) {
#pragma unroll
            for (int j2 = 0; j2 <= J - j1 - 1; ++j2) acc[j1 + 1 + j2] = fma(t.y, gO[j2], acc[j1 + 1 + j2]);
        }
    }
    {   // y1[J] of a caller-supplied generator (zero for generate_y's)
        double a;
        asm volatile("ld.global.f64 %0, [%1];" : "=d"(a) : "l"(q.gy + yp + 8 * R::IDXJ));
        acc[J] = fma(a, gO[0], acc[J]);
    }
#pragma unroll
    for (int k = 0; k < R::NY2; ++k) {       // - conv(y2, dA g)
        const double2 t = ldy_pr(q, yp + 16 * (R::ND2 + k));
        const int j1 = 2 * k;
        if (j1 <= J) {
#pragma unroll
            for (int j2 = 0; j2 < J && j1 + j2 <= J; ++j2) acc[j1 + j2] = fma(-t.x, gD[j2], acc[j1 + j2]);
        }
        if (j1 + 1 <= J) {
#pragma unroll
            for (int j2 = 0; j2 < J && j1 + 1 + j2 <= J; ++j2) acc[j1 + 1 + j2] = fma(-t.y, gD[j2], acc[j1 + 1 + j2]);
        }
    }
}

__device__ __forceinline__ LaneGeoPR lane_geometry_pr(const Ctx& c, int box) {
    const DevPlan& P = RG_PLAN(c);
    const Sched& S = RG_SCHED(c);
    const Smem sm = smem_view<true>(P, S, c);
    const int lane = threadIdx.x & 31;
    const uint32_t mc = S.lane_coord[box * 32 + lane];
    LaneGeoPR q;
    q.lanebit = 1u << lane;
    q.lin = ((int)(mc & 0xff) + P.L) * P.stride0 + ((int)((mc >> 8) & 0xff) + P.L) * P.stride1 + ((int)((mc >> 16) & 0xff) + P.L) * P.stride2;
    q.zero_rec = P.M;
    q.g_bytes = P.g_units * 16;
    q.y_bytes = P.rec_units * 16;
    q.gg = reinterpret_cast<const char*>(sm.G);
    q.gy = reinterpret_cast<const char*>(sm.Y);
    return q;
}

// One piece of a box's source list on the fused path: S1, S2 of the product-rule form, summed from zero.  The
// planner sorts a list by level class, so a piece is at most three runs (ends e_full <= e_mid <= e_end): one tight
// loop per class, two entries per trip so that the gathered record alternates between two register sets.
template <int J>
__device__ __forceinline__ void conv_piece_pr(const Ctx& c, int box, int e_begin, int e_full, int e_mid, int e_end,
                                           double (&S1)[J + 1], double (&S2)[J + 1]) {
    typedef RecGeomPR<J> R;
    const uint2* const entries = RG_SCHED(c).entries;
    const LaneGeoPR q = lane_geometry_pr(c, box);
#pragma unroll
    for (int j = 0; j <= J; ++j) { S1[j] = 0.0; S2[j] = 0.0; }
    if (e_begin >= e_end) return;
    double ga[2 * R::GU], gb[2 * R::GU];
    double2 ya[R::YREGS], yb[R::YREGS];
    uint32_t gp, yp;
    {
        int nat;
        decode_entry_pr(q, entries[e_begin], gp, yp, nat);
    }
#pragma unroll
    for (int u = 0; u < R::GU; ++u) { const double2 v = ldg_pr(q, gp + 16 * u); ga[2 * u] = v.x; ga[2 * u + 1] = v.y; }
    if (R::YALL) {
#pragma unroll
        for (int k = 0; k < R::N1MAX; ++k) ya[k] = ldy_pr(q, yp + 16 * k);
#pragma unroll
        for (int k = 0; k < R::N2MAX; ++k) ya[R::N1MAX + k] = ldy_pr(q, yp + 16 * (R::ND2 + k));
    } else {
        ya[0] = ldy_pr(q, yp);
    }
    // list words of the next entry (decoded inside the current product) and of the one after (in flight); past the
    // end the last entry is read again (harmless, keeps the loop uniform)
    const int last = e_end - 1;
    int i = e_begin;
    uint2 wn = entries[min(i + 1, last)];
#define RG_RUN(LIM, RUN_END)                                                                      \
    _Pragma("unroll 1")                                                                           \
    for (; i + 2 <= (RUN_END); i += 2) {                                                          \
        const uint2 w2 = entries[min(i + 2, last)], w3 = entries[min(i + 3, last)];               \
        taylor_leib<J, LIM>(S1, S2, ga, gb, ya, yb, q, wn, gp, yp);                               \
        if (R::YALL) taylor_leib<J, LIM>(S1, S2, gb, ga, yb, ya, q, w2, gp, yp);                  \
        else taylor_leib<J, LIM>(S1, S2, gb, ga, ya, yb, q, w2, gp, yp);                          \
        wn = w3;                                                                                  \
    }                                                                                             \
    if (i < (RUN_END)) {                                                                          \
        const uint2 w2 = entries[min(i + 2, last)];                                               \
        taylor_leib<J, LIM>(S1, S2, ga, gb, ya, yb, q, wn, gp, yp);                               \
        _Pragma("unroll")                                                                         \
        for (int u = 0; u < 2 * R::GU; ++u) ga[u] = gb[u];                                        \
        if (R::YALL) {                                                                            \
            _Pragma("unroll")                                                                     \
            for (int u = 0; u < R::YREGS; ++u) ya[u] = yb[u];                                     \
        }                                                                                         \
        wn = w2; ++i;                                                                             \
    }
    RG_RUN(J + 1, e_full)
    RG_RUN(R::LIM_B, e_mid)
    RG_RUN(R::LIM_A, e_end)
#undef RG_RUN
}

// One piece of a caller-supplied generator's list (schedule kind 1): the bracket itself in acc.
template <int J>
__device__ __noinline__ void conv_piece_general(const Ctx& c, int box, int e_begin, int e_end, double (&acc)[J + 1]) {
    const uint2* const entries = RG_SCHED(c).entries;
    const LaneGeoPR q = lane_geometry_pr(c, box);
#pragma unroll
    for (int j = 0; j <= J; ++j) acc[j] = 0.0;
#pragma unroll 1
    for (int i = e_begin; i < e_end; ++i) taylor_general<J>(acc, q, entries[i], c.Omega_nu);
}

// out[nu][j] = conv(dA y, Omega g) - conv(Omega y, dA g)  (G, Y in shared memory).  Ends with a barrier.
// Every warp owns up to S.maxp pieces of the boxes' source lists per stage: heads (rank 0) store their sums,
// the single continuation a warp may hold (rank r > 0) adds once the box's counter says r, stage-1 pieces once
// every stage-0 piece is in -- a fixed order of every sum, so the result does not depend on timing, and no
// barrier until the end of the bracket.  On the half-space schedule only modes nu <= 0 are produced.
//
// Cluster mode (csize = 2 or 4 CTAs per Hamiltonian): every CTA of the cluster runs the whole program on
// its own copy of the state (bit-identical control flow); only this function is split.  The planner gives
// each CTA a contiguous group of whole boxes, CTA r computes them exactly as above with its own warps
// (virtual warps [r*W, (r+1)*W)) and then broadcasts the finished rows into the `out` buffer of every peer
// through distributed shared memory: two cluster barriers and (csize-1) x its share of `out` per term.
#ifndef RG_CONV_ATTR
#define RG_CONV_ATTR __noinline__
#endif
#ifndef RG_EPI_ATTR
#define RG_EPI_ATTR __noinline__
#endif
template <int J, bool BIG>
__device__ RG_CONV_ATTR void conv_all(Ctx& c) {
    constexpr int JP = J + 1;
    namespace cg = cooperative_groups;
    const Sched& S = RG_SCHED(c);
    double* const out = smem_view<BIG>(RG_PLAN(c), S, c).out;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, nwc = (int)(blockDim.x >> 5);
    const int vw = c.crank * nwc + w;
    const long long t_begin = clock64();
    if (c.csize > 1) cg::this_cluster().sync();   // every peer has finished reading its `out` of the previous term
    // Ordered combination without barriers: ctl[b] = number of pieces of box b already combined into its rows,
    // ctl[n_boxes] = stage-0 pieces combined.  A piece of rank r waits until its box's counter says r (the head
    // stores, the others add), a stage-1 piece also until every stage-0 piece is in; the order of every sum is
    // that of the plan, and no warp waits for the slowest one before the end of the bracket.
    int* const ctl = smem_ctl();
    volatile int* const vctl = ctl;
    const int nb_all = S.n_boxes, need0 = S.stage0_pieces[c.crank];
    for (int i = threadIdx.x; i <= nb_all; i += blockDim.x) ctl[i] = 0;
    __syncthreads();
    double acc[JP];
    const long long t_stage = clock64();
    if (vw < S.n_warps) {
        for (int st = 0; st < S.n_stages; ++st) {
            const int4* pieces = S.pieces + (size_t)st * S.n_warps * S.maxp;
            {
            const int pw = vw;
#pragma unroll 1
            for (int q = 0; q < S.maxp; ++q) {
                const int4 p = pieces[pw * S.maxp + q];      // heads first, the continuation (if any) last
                if (p.x < 0) continue;
                if constexpr (BIG) {
                    // product-rule form: the piece's contribution is (Omega.nu) S1 - (j+1) S2 of its own partial sums
                    // (the map is linear, so pieces add); a caller-supplied generator takes the general product
                    if (S.general) {
                        conv_piece_general<J>(c, p.x, p.y, p.z, acc);
                    } else {
                        double S1[JP], S2[JP];
                        const int2 r = S.piece_runs[((size_t)st * S.n_warps + pw) * S.maxp + q];
                        conv_piece_pr<J>(c, p.x, p.y, r.x, r.y, p.z, S1, S2);
                        const int nu_ = S.lane_out[p.x * 32 + lane];
                        const double om = nu_ >= 0 ? c.Omega_nu[nu_] : 0.0;
#pragma unroll
                        for (int j = 0; j < JP; ++j) acc[j] = om * S1[j] - (double)(j + 1) * S2[j];
                    }
                } else {
                    const int2 r = S.piece_runs[((size_t)st * S.n_warps + pw) * S.maxp + q];
                    conv_piece<J>(c, p.x, p.y, r.x, r.y, p.z, acc);
                }
                if (lane == 0) {
                    if (st) while (vctl[nb_all] < need0) {}
                    while (vctl[p.x] != p.w) {}
                }
                __syncwarp();
                __threadfence_block();
                const int nu = S.lane_out[p.x * 32 + lane];
                if (nu >= 0) {
                    if (st == 0 && p.w == 0) {
#pragma unroll
                        for (int j = 0; j < JP; ++j) out[nu * JP + j] = acc[j];
                    } else {
#pragma unroll
                        for (int j = 0; j < JP; ++j) out[nu * JP + j] += acc[j];
                    }
                }
                __threadfence_block();   // the rows before the counters
                __syncwarp();
                if (lane == 0) {
                    vctl[p.x] = p.w + 1;
                    if (st == 0) atomicAdd(&ctl[nb_all], 1);
                }
            }
            }
        }
    }
    c.t_busy += clock64() - t_stage;
    __syncthreads();
    if (c.csize > 1) {
        cg::cluster_group cl = cg::this_cluster();
        for (int b = S.fin_begin[c.crank] + w; b < S.fin_end[c.crank]; b += nwc) {
            const int nu = S.lane_out[b * 32 + lane];
            if (nu < 0) continue;
            double v[JP];
#pragma unroll
            for (int j = 0; j < JP; ++j) v[j] = out[nu * JP + j];
            for (int r = 0; r < c.csize; ++r) {
                if (r == c.crank) continue;
                // peers are consecutive CTAs: their `out` is at the same shared-memory offset (DSMEM), or -- BIG kernels
                // with the g records on chip -- in their slice of the global workspace
                double* o = ((BIG && RG_PLAN(c).g_smem) ? out + (long long)(r - c.crank) * RG_PLAN(c).ws_stride : cl.map_shared_rank(out, r)) + nu * JP;
#pragma unroll
                for (int j = 0; j < JP; ++j) o[j] = v[j];
            }
        }
        cl.sync();
    }
    c.n_liou++;
    if (S.half == 0) c.n_liou_full++;
    c.t_conv += clock64() - t_begin;
}

// ------------------------------------------------------------------------------------------------
// records of g: d-type = dA g (RG.py:76 roll(f*J_,-1)), Omega-type = Omega_nu*g.  src: global [j][nu].
// (The per-Lie-step helpers below are templated on J so that their level loops unroll and the J+1 loads of a
// mode are in flight together: with a run-time J every level waited for its own L2 round trip.)
template <int J, bool BIG>
__device__ __noinline__ void build_G(const Ctx& c, const double* __restrict__ src) {
    const DevPlan& P = RG_PLAN(c);
    const int M = P.M;
    const Smem sm = smem_view<BIG>(P, RG_SCHED(c), c);
    // exactly symmetric src (c.symm): the thread of nu <= 0 also writes the mirror record (one pass over 221 of the
    // 441 modes at L = 10 in one pass of the CTA's threads)
    const bool mirror = c.symm != 0;
    const int nmax = mirror ? (M + 1) / 2 : M;
    for (int nu = threadIdx.x; nu < nmax; nu += blockDim.x) {
        const int mu = M - 1 - nu;
        double g[J + 1];
#pragma unroll
        for (int j = 0; j <= J; ++j) g[j] = src[j * M + nu];
        if constexpr (BIG) {   // product-rule form: the record is g itself (and so is the mirror's: g is symmetric)
            double* r = reinterpret_cast<double*>(sm.G + nu * P.g_units);
#pragma unroll
            for (int j = 0; j <= J; ++j) r[j] = g[j];
            if (mirror && mu != nu) {
                double* r2 = reinterpret_cast<double*>(sm.G + mu * P.g_units);
#pragma unroll
                for (int j = 0; j <= J; ++j) r2[j] = g[j];
            }
        } else {
            double* r = reinterpret_cast<double*>(sm.G + nu * P.rec_units);
            const double om = c.Omega_nu[nu];
#pragma unroll
            for (int j = 0; j <= J; ++j) {
                r[P.off2 + j] = om * g[j];
                if (j < J) r[j] = g[j + 1] * (double)(j + 1);
            }
            if (mirror && mu != nu) {
                double* r2 = reinterpret_cast<double*>(sm.G + mu * P.rec_units);
#pragma unroll
                for (int j = 0; j <= J; ++j) {
                    r2[P.off2 + j] = -(om * g[j]);
                    if (j < J) r2[j] = g[j + 1] * (double)(j + 1);
                }
            }
        }
    }
}

// records of y scaled by `step` (item*step, RG.py:176) or divided by `scale` (item/scale, RG.py:156)
template <int J, bool BIG>
__device__ __noinline__ void build_Y(const Ctx& c, double step, double divisor, bool use_div) {
    const DevPlan& P = RG_PLAN(c);
    const int M = P.M;
    const double* y1 = c.ws + (size_t)WS_Y1 * P.vecjl;
    const double* y2 = c.ws + (size_t)WS_Y2 * P.vecjl;
    const Sched& S = RG_SCHED(c);
    const Smem sm = smem_view<BIG>(P, S, c);
    // y of an exactly symmetric f: y1 = dA y is exactly antisymmetric, y2 = Omega y exactly symmetric
    const bool mirror = c.symm != 0 && P.sym_tables != 0;
    const int nmax = mirror ? (M + 1) / 2 : M;
    for (int nu = threadIdx.x; nu < nmax; nu += blockDim.x) {
        const int rec = S.supp_rec[nu];
        if (rec < 0) continue;
        double* r = reinterpret_cast<double*>(sm.Y + rec * P.rec_units);
        const int mu = M - 1 - nu;
        double* r2 = (mirror && mu != nu) ? reinterpret_cast<double*>(sm.Y + S.supp_rec[mu] * P.rec_units) : nullptr;
#pragma unroll
        for (int j = 0; j <= J; ++j) {
            double a = y1[j * M + nu], b = y2[j * M + nu];
            a = use_div ? a / divisor : a * step;
            b = use_div ? b / divisor : b * step;
            r[j < J ? j : P.idx_y1J] = a;
            r[P.off2 + j] = b;
            if (r2) { r2[j < J ? j : P.idx_y1J] = -a; r2[P.off2 + j] = b; }
        }
    }
}

// One series term: sh = [y3 +] (a0*dA g + conv) [/ m]; f += sh; G <- records of sh.
// Returns (norm(sh), norm(f)).  Thread <-> mode nu.  On the half-space schedule the thread of nu <= 0 also
// writes the mirror mode -nu (sh, f are exactly symmetric, Omega_nu exactly antisymmetric).
// SUM: NormChoice = 'sum' (the default) known at compile time -- no norm dispatch, no weights in the item loop.
template <int J, bool BIG, bool SUM>
__device__ RG_EPI_ATTR double2 series_epilogue_impl(Ctx& c, double* __restrict__ f, double a0, bool first, double y3mul,
                                                     double y3div, bool use_div, double m) {
    constexpr int JP = J + 1;
    constexpr int CH = 8;   // items a thread holds between the read and the write phase of a chunk
    const DevPlan& P = RG_PLAN(c);
    const int M = P.M, type = SUM ? 0 : P.norm_choice;
    const bool half = RG_SCHED(c).half != 0;
    const int nmax = half ? (M + 1) / 2 : M;
    const int total = nmax * JP;
    const double* y3 = c.ws + (size_t)WS_Y3 * P.vecjl;
    const Smem sm = smem_view<BIG>(P, RG_SCHED(c), c);
    const long long t_begin = clock64();
    double ps = 0.0, pf = 0.0;
    // Work item = one coefficient (j, nu), j-major, dealt round-robin to ALL threads of the CTA (a thread per mode
    // would leave 40 % of the warps idle on the half-space schedule and serialise 11 levels per thread).  The new
    // d-type entry r[j-1] = j*sh[j] overwrites what item (j-1, nu) reads as dA g: items are processed in chunks of
    // CH per thread, reads first, a barrier, then the writes; (j-1, nu) is in the same or an earlier chunk.
    // (j, nu) of item i = base + q*blockDim + tid, advanced without divisions
    const int dj = (int)blockDim.x / nmax, dnu = (int)blockDim.x - dj * nmax;
    int ja = (int)threadIdx.x / nmax, nua = (int)threadIdx.x - ja * nmax;
    for (int base = 0; base < total; base += CH * (int)blockDim.x) {
        double v_[CH], fo_[CH], om_[CH], w_[CH];
        int jb = ja, nub = nua;      // the write phase walks the same items again
#pragma unroll
        for (int q = 0; q < CH; ++q) {
            const int i = base + q * (int)blockDim.x + (int)threadIdx.x;
            const int j = ja, nu = nua;
            ja += dj; nua += dnu;
            if (nua >= nmax) { nua -= nmax; ++ja; }
            v_[q] = 0.0; fo_[q] = 0.0; om_[q] = 0.0; w_[q] = 1.0;
            if (i < total) {
                // every global load of the item belongs to the read phase: behind the first store to f the compiler must
                // assume aliasing, and a load issued between the stores waits out its own L2 round trip (the write phase
                // used to load Omega_nu[nu] per item: 7 serialised round trips, 4 k of the epilogue's 6.8 k clocks)
                if (!BIG) om_[q] = c.Omega_nu[nu];
                if (!SUM && type == 3) w_[q] = P.normw[nu];
                double dg;      // dA g  (RG.py:76)
                if constexpr (BIG) {
                    const double* r = reinterpret_cast<const double*>(sm.G + nu * P.g_units);
                    dg = (j < J) ? r[j + 1] * (double)(j + 1) : 0.0;
                } else {
                    const double* r = reinterpret_cast<const double*>(sm.G + nu * P.rec_units);
                    dg = (j < J) ? r[j] : 0.0;
                }
                double v = a0 * dg + sm.out[nu * JP + j];
                if (first) {
                    double t = y3[j * M + nu];
                    t = use_div ? t / y3div : t * y3mul;
                    v = t + v;
                } else {
                    v = v / m;
                }
                v_[q] = v;
                fo_[q] = f[j * M + nu];
            }
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < CH; ++q) {
            const int i = base + q * (int)blockDim.x + (int)threadIdx.x;
            const int j = jb, nu = nub;
            jb += dj; nub += dnu;
            if (nub >= nmax) { nub -= nmax; ++jb; }
            if (i < total) {
                const int mu = M - 1 - nu;
                const bool twin = half && (mu != nu);
                const double om = om_[q];
                const double w = w_[q];
                const double v = v_[q], fn = fo_[q] + v;
                f[j * M + nu] = fn;
                norm_acc(type, ps, v, w);
                norm_acc(type, pf, fn, w);
                if (twin) {
                    f[j * M + mu] = fn;
                    norm_acc(type, ps, v, w);
                    norm_acc(type, pf, fn, w);
                }
                if constexpr (BIG) {
                    // item (j, nu) read the OLD g[j+1] above and writes the new g[j]: (j+1, nu), which overwrites g[j+1],
                    // is in the same or a later chunk (behind this chunk's barrier)
                    reinterpret_cast<double*>(sm.G + nu * P.g_units)[j] = v;
                    if (twin) reinterpret_cast<double*>(sm.G + mu * P.g_units)[j] = v;
                } else {
                    double* r = reinterpret_cast<double*>(sm.G + nu * P.rec_units);
                    r[P.off2 + j] = om * v;
                    if (j > 0) r[j - 1] = v * (double)j;
                    if (twin) {
                        double* r2 = reinterpret_cast<double*>(sm.G + mu * P.rec_units);
                        r2[P.off2 + j] = -(om * v);
                        if (j > 0) r2[j - 1] = v * (double)j;
                    }
                }
            }
        }
    }
    const double2 res = block_reduce2(smem_red(), type, ps, pf);
    c.t_epi += clock64() - t_begin;
    return res;
}
template <int J, bool BIG>
__device__ __forceinline__ double2 series_epilogue(Ctx& c, double* __restrict__ f, double a0, bool first, double y3mul,
                                                   double y3div, bool use_div, double m) {
    if (RG_PLAN(c).norm_choice == 0) return series_epilogue_impl<J, BIG, true>(c, f, a0, first, y3mul, y3div, use_div, m);
    return series_epilogue_impl<J, BIG, false>(c, f, a0, first, y3mul, y3div, use_div, m);
}

static __device__ __noinline__ void copy_arr(const Ctx& c, double* dst, const double* src) {
    const int n = RG_PLAN(c).vecjl;
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}

// RG.expm_onestep (RG.py:173-192) in place on f (global).  y in WS_Y1..3, a_unscaled = y[0].
template <int J, bool BIG>
__device__ __noinline__ int expm_series(Ctx& c, double* f, double a_unscaled, double step, double divisor, bool use_div, int m_star) {
    const DevPlan& P = RG_PLAN(c);
    const double tol = 1.1102230246251565e-16;
    const double a0 = use_div ? a_unscaled / divisor : a_unscaled * step;
    int err = 0;
    const int sidx_saved = c.sidx;
    if (c.sidx == 0 && c.symm) c.sidx = 2;   // exactly symmetric f and antisymmetric y: compute nu <= 0 only
    build_Y<J, BIG>(c, step, divisor, use_div);
    build_G<J, BIG>(c, f);
    __syncthreads();
    // one loop for all terms (the first one adds y3 and is not divided): conv_all and the epilogue appear once
    double c1 = 0.0;
    for (int m = 1; m == 1 || m_star <= 0 || m <= m_star; ++m) {
        const bool first = (m == 1);
        conv_all<J, BIG>(c);
        const double2 r = series_epilogue<J, BIG>(c, f, a0, first, first ? step : 0.0, first ? divisor : 1.0, first && use_div, (double)m);
        if (first) { c1 = r.x; continue; }
        const double c2 = r.x;
        if (c1 + c2 < tol * r.y) break;
        else if (c1 + c2 > P.tol_max) { err = 1; break; }
        c1 = c2;
        if (m >= P.max_series || c.n_liou - c.budget_base > P.max_liouville) { err = 1; break; }  // safety: the reference would spin (quirk Q6)
    }
    c.sidx = sidx_saved;
    return err;
}

template <int J, bool BIG>
__device__ __forceinline__ int expm_onestep(Ctx& c, double* f, double a_unscaled, double step) {
    return expm_series<J, BIG>(c, f, a_unscaled, step, 1.0, false, 0);
}

// norm(h1 - h2), norm(h1); skip_centre: the difference is taken with its nu = 0 column zeroed (RG.norm_int, RG.py:244-247)
template <int J>
__device__ __noinline__ double2 diff_norms(const Ctx& c, const double* h1, const double* h2, bool skip_centre = false) {
    const DevPlan& P = RG_PLAN(c);
    const int M = P.M, type = P.norm_choice, c0 = (M - 1) / 2;
    double pd = 0.0, p1 = 0.0;
    for (int nu = threadIdx.x; nu < M; nu += blockDim.x) {
        const double w = (type == 3) ? P.normw[nu] : 1.0;
#pragma unroll
        for (int j = 0; j <= J; ++j) {
            const double a = h1[j * M + nu], b = h2[j * M + nu];
            if (!(skip_centre && nu == c0)) norm_acc(type, pd, a - b, w);
            norm_acc(type, p1, a, w);
        }
    }
    return block_reduce2(smem_red(), type, pd, p1);
}

// RG.expm_adapt (RG.py:194-205): the recursion unrolled into a depth-first stack of sub-steps.
template <int J, bool BIG>
__device__ __noinline__ int expm_adapt(Ctx& c, double* f, double a_unscaled, double step0, int err_in) {
    const DevPlan& P = RG_PLAN(c);
    double* h1 = c.ws + (size_t)WS_H1 * P.vecjl;
    double* h2 = c.ws + (size_t)WS_H2 * P.vecjl;
    double stack[64];
    int sp = 0, err = err_in;
    stack[sp++] = step0;
    while (sp > 0) {
        const double s = stack[--sp];
        if (s < P.min_step || sp >= 60) {  // RG.py:196-198 (error=4 is dead, quirk Q3)
            err = expm_onestep<J, BIG>(c, f, a_unscaled, s);
            __syncthreads();
            continue;
        }
        copy_arr(c, h1, f);
        copy_arr(c, h2, f);
        __syncthreads();
        expm_onestep<J, BIG>(c, h1, a_unscaled, s);
        __syncthreads();
        expm_onestep<J, BIG>(c, h2, a_unscaled, 0.5 * s);
        __syncthreads();
        expm_onestep<J, BIG>(c, h2, a_unscaled, 0.5 * s);
        __syncthreads();
        const double2 r = diff_norms<J>(c, h1, h2);
        if (r.x < P.abs_tol + P.rel_tol * r.y) {
            for (int i = threadIdx.x; i < P.vecjl; i += blockDim.x) f[i] = 0.75 * h1[i] + 0.25 * h2[i];
            __syncthreads();
        } else {
            stack[sp++] = 0.5 * s;
            stack[sp++] = 0.5 * s;
        }
    }
    return err;
}

// RG.generate_y (RG.py:108-113): writes unscaled y1,y2,y3 to WS_Y1..3; returns y[0]
template <int J>
__device__ __noinline__ double generate_y(const Ctx& c, const double* __restrict__ f) {
    const DevPlan& P = RG_PLAN(c);
    const int M = P.M, c0 = (M - 1) / 2;
    double* y1 = c.ws + (size_t)WS_Y1 * P.vecjl;
    double* y2 = c.ws + (size_t)WS_Y2 * P.vecjl;
    double* y3 = c.ws + (size_t)WS_Y3 * P.vecjl;
    const double f2 = f[2 * M + c0], f1 = f[1 * M + c0];
    // exactly symmetric f and antisymmetric omega0_nu, Omega_nu: y[-nu] = -y[nu] exactly (every operation below is
    // odd or even in the sign flips), so the thread of nu <= 0 writes both modes
    const bool mirror = c.symm != 0 && P.sym_tables != 0;
    const int nmax = mirror ? c0 + 1 : M;
    for (int nu = threadIdx.x; nu < nmax; nu += blockDim.x) {
        const unsigned mk = P.mask[nu];
        const int mu = M - 1 - nu;
        const bool twin = mirror && mu != nu;
        const double om0 = P.omega0_nu[nu], om = c.Omega_nu[nu];
        double fv[J + 1];
#pragma unroll
        for (int j = 0; j <= J; ++j) fv[j] = f[j * M + nu];
        double yprev = 0.0;
#pragma unroll
        for (int j = 0; j <= J; ++j) {
            double y = 0.0;
            if ((mk >> j) & 1u) {
                if (j == 0) y = fv[0] / om0;
                else y = (fv[j] - 2 * f2 * om * yprev) / om0;
            }
            const double v1 = y * (double)j, v2 = om * y, v3 = -om0 * y;
            if (j >= 1) y1[(j - 1) * M + nu] = v1;
            y2[j * M + nu] = v2;
            y3[j * M + nu] = v3;
            if (twin) {
                if (j >= 1) y1[(j - 1) * M + mu] = -v1;
                y2[j * M + mu] = v2;
                y3[j * M + mu] = v3;
            }
            yprev = y;
        }
        y1[J * M + nu] = 0.0;
        if (twin) y1[J * M + mu] = 0.0;
    }
    return -f1 / (2 * f2);
}

// RG.sym (RG.py:249-253) in place, fused with the I^- norm (RG.py:217-218,221) and norm_int (RG.py:244-247)
template <int J>
__device__ __noinline__ double2 sym_and_norms(Ctx& c, double* f, bool do_sym) {
    const DevPlan& P = RG_PLAN(c);
    const int M = P.M, type = P.norm_choice, c0 = (M - 1) / 2;
    double pim = 0.0, pint = 0.0;
    int asym = 0;
    for (int nu = threadIdx.x; nu <= c0; nu += blockDim.x) {
        const int mu = M - 1 - nu;
        const unsigned mk_a = P.mask[nu], mk_b = P.mask[mu];
        const double wa = (type == 3) ? P.normw[nu] : 1.0, wb = (type == 3) ? P.normw[mu] : 1.0;
        double fa_[J + 1], fb_[J + 1];
#pragma unroll
        for (int j = 0; j <= J; ++j) { fa_[j] = f[j * M + nu]; fb_[j] = f[j * M + mu]; }
#pragma unroll
        for (int j = 0; j <= J; ++j) {
            double a = fa_[j], b = fb_[j];
            if (do_sym) {
                const double s1 = (a + b) / 2, s2 = (b + a) / 2;
                a = s1; b = s2;
                if (nu == c0 && j == 0) { a = 0.0; b = 0.0; }
                f[j * M + nu] = a;
                if (mu != nu) f[j * M + mu] = b;
            }
            asym |= !(a == b);
            if ((mk_a >> j) & 1u) norm_acc(type, pim, a, wa);
            if (nu != c0) {
                if ((mk_b >> j) & 1u) norm_acc(type, pim, b, wb);
                norm_acc(type, pint, a, wa);
                norm_acc(type, pint, b, wb);
            }
        }
    }
    c.symm = __syncthreads_or(asym) ? 0 : 1;
    return block_reduce2(smem_red(), type, pim, pint);
}

// head of RG.rg_map (RG.py:210-216): dst = ren[j] * src[j][sign*N*nu] on the mask, else 0
template <int J>
__device__ __noinline__ void reindex_rescale(const Ctx& c, const double* __restrict__ src, double* __restrict__ dst, double nrm) {
    const DevPlan& P = RG_PLAN(c);
    const int M = P.M, c0 = (M - 1) / 2;
    const double f2 = src[2 * M + c0];
    const double base = 2 * nrm / P.eigenvalue * f2;
    double* red = smem_red();
    if (threadIdx.x <= J) red[threadIdx.x] = pow(base, (double)(2 - (int)threadIdx.x)) / (2 * f2);
    __syncthreads();
    for (int nu = threadIdx.x; nu < M; nu += blockDim.x) {
        const int s = P.reidx[nu];
#pragma unroll
        for (int j = 0; j <= J; ++j) dst[j * M + nu] = (s >= 0 ? src[j * M + s] : 0.0) * red[j];
    }
    __syncthreads();
}

struct Tables {
    int n_tab;
    const double* nrm;       // [n_tab]
    const double* Omega_nu;  // [(n_tab+1)][M] centred
};

template <int J, bool BIG> __device__ int expm_multiply(Ctx& c, double* f, double a_unscaled);  // rg_multiply.cuh

// RG.rg_map (RG.py:207-228).  f points to the current array (WS_FA or WS_FB); returns the new current.
template <int J, bool BIG>
__device__ __noinline__ double* rg_map_dev(Ctx& c, double* f, const Tables& T, int k, int& err_out, double& norm_int_out) {
    const DevPlan& P = RG_PLAN(c);
    double* fa = c.ws + (size_t)WS_FA * P.vecjl;
    double* fb = c.ws + (size_t)WS_FB * P.vecjl;
    c.Omega_nu = T.Omega_nu + (size_t)(k + 1) * P.M;
    double* dst = (f == fa) ? fb : fa;
    reindex_rescale<J>(c, f, dst, T.nrm[k]);
    f = dst;
    int err = 0, kl = 0;
    double2 nn = sym_and_norms<J>(c, f, false);
    while (P.tol_max > nn.x && nn.x > P.tol_min && kl < P.max_lie && c.n_liou - c.budget_base <= P.max_liouville) {
        double a;
        a = generate_y<J>(c, f);
        __syncthreads();
        if (P.ct == 0) err = expm_onestep<J, BIG>(c, f, a, 1.0);
        else if (P.ct == 1) err = expm_adapt<J, BIG>(c, f, a, 1.0, err);
        else err = expm_multiply<J, BIG>(c, f, a);
        __syncthreads();
        nn = sym_and_norms<J>(c, f, true);
        ++kl;
        c.n_lie++;
    }
    if (nn.x > P.tol_max) err = 2;
    else if (kl >= P.max_lie) err = -2;
    err_out = err;
    norm_int_out = nn.y;
    return f;
}

// RG.converge (RG.py:278-290) on the array f (WS_FA or WS_FB): iterates rg_map from table row k0 until the
// interior norm leaves (TolMin, TolMax).  Returns k >= 0 and sets conv; status: 0 ok, 1 table exhausted, 2 work budget.
template <int J, bool BIG>
__device__ __forceinline__ int converge_dev(Ctx& c, double*& f, const Tables& T, int k0, bool& conv, int& err, int& status,
                                            long long& n_maps) {
    const DevPlan& P = RG_PLAN(c);
    double2 nn = sym_and_norms<J>(c, f, false);
    double ni = nn.y;
    int k = 0;
    err = 0; status = 0;
    c.budget_base = c.n_liou;
    while (P.tol_max > ni && ni > P.tol_min) {
        if (k0 + k >= T.n_tab) { status = 1; break; }
        if (c.n_liou - c.budget_base > P.max_liouville) { status = 2; break; }
        f = rg_map_dev<J, BIG>(c, f, T, k0 + k, err, ni);
        ++k; ++n_maps;
    }
    if (status == 0 && c.n_liou - c.budget_base > P.max_liouville) status = 2;   // cut short inside the last rg_map
    conv = (ni < P.tol_min) && status == 0;
    return k;
}

}  // namespace rgdev
