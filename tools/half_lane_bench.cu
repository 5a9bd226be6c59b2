// half_lane_bench.cu -- design study for the bracket's inner loop (rg_kernel<10,false>): the shipped product
// (taylor_full: one lane = one output mode, both convolutions, 130 DFMA per list entry, 168 registers, 12 warps per
// SM) against a "half-lane" product: a lane PAIR owns an output mode, lanes 0-15 of a warp run conv(dA y, Omega g)
// and lanes 16-31 run -conv(Omega y, dA g) for the same 16 modes with ONE uniform instruction stream (66 DFMA per
// entry, each lane gathers only its half of the record), fewer registers, more resident warps.
// Both variants run the real record layout, gather pattern (conflict-free quarter-warps), list-word loads and entry
// decode out of shared memory; F calls the product code of rg_device.cuh itself.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -lineinfo -o tools/half_lane_bench tools/half_lane_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../rg_b200/csrc/rg_device.cuh"
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
using namespace rgdev;

constexpr int J = 10, M = 441, NSUP = 252, REC_UNITS = 11;

// ---------------------------------------------------------------- F: the shipped product
__global__ void __launch_bounds__(384, 1) full_kernel(const uint2* __restrict__ entries, int n_entries, int iters, double* out,
                                                      long long* clk) {
    typedef RecGeom<J> R;
    double2* G = smem_dyn;
    double2* Y = G + (M + RG_ZERO_RECS) * REC_UNITS;
    for (int i = threadIdx.x; i < (M + RG_ZERO_RECS) * REC_UNITS; i += blockDim.x)
        G[i] = i < M * REC_UNITS ? make_double2(1e-3 * (i % 7), -1e-3 * (i % 5)) : make_double2(0.0, 0.0);
    for (int i = threadIdx.x; i < NSUP * REC_UNITS; i += blockDim.x) Y[i] = make_double2(1e-3 * (i % 3), 1e-3 * (i % 11));
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    LaneGeo q;
    q.v0 = 10 + (lane >> 3); q.v1 = 10 + (lane & 7) + (warp & 1); q.v2 = 10;
    q.lin = q.v0 * 21 + q.v1;
    q.lim = 20u; q.zero_rec = M; q.rec_bytes = REC_UNITS * 16;
    q.g_base = (uint32_t)__cvta_generic_to_shared(G);
    q.y_base = (uint32_t)__cvta_generic_to_shared(Y);
    double acc[J + 1], gO[2 * R::NO2];
    double2 yq[R::YW];
#pragma unroll
    for (int j = 0; j <= J; ++j) acc[j] = 0.0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const uint2* const ent = entries + (size_t)warp * n_entries;
        const int last = n_entries - 1;
        uint2 wn = ent[1];
        uint32_t gp, yp; int jm;
        decode_entry(q, ent[0], gp, yp, jm);
#pragma unroll
        for (int u = 0; u < R::NO2; ++u) { const double2 t = ldrec(q, gp + 16 * (R::ND2 + u)); gO[2 * u] = t.x; gO[2 * u + 1] = t.y; }
#pragma unroll
        for (int k = 0; k < R::YW; ++k) yq[k] = ldrec(q, yp + 16 * R::unit(k));
        int i = 0;
#pragma unroll 1
        while (i < n_entries && jm > R::LIM_B) {
            const uint2 wnn = ent[min(i + 2, last)];
            taylor_full<J, J + 1>(acc, gO, yq, q, wn, gp, yp, jm);
            wn = wnn; ++i;
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int j = 0; j <= J; ++j) s += acc[j];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

// ---------------------------------------------------------------- H: half-lane product
// G record (11 units): u0..u4 = dA g pairs (g2[0..9]); u5..u10 = Omega g pairs (g1[0..10], 0).
//   A lanes (0-15) read u5..u10; B lanes (16-31) read u0..u4 and, as sixth unit, a zero unit.
// Y record (12 units): u0..u5 = dA y[0..9], 0, 0 (A lanes); u6..u11 = -Omega y[0..10], 0 (B lanes).
constexpr int YH_UNITS = 12;
struct LaneGeoH {
    int v0, v1, v2, lin;
    unsigned lim;
    int zero_rec;
    uint32_t g_base, y_base;   // already offset by the lane's half
    uint32_t zero_unit;        // B lanes: address of an all-zero unit; A lanes: 0 (use the record's own sixth unit)
    int rec_bytes;
};
__device__ __forceinline__ double2 lds128(uint32_t addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void decode_h(const LaneGeoH& q, const uint2& w, uint32_t& gp, uint32_t& g5, uint32_t& yp) {
    const uint32_t ec = w.x, ei = w.y;
    const unsigned u0 = (unsigned)(q.v0 - (int)(ec & 0xff)), u1 = (unsigned)(q.v1 - (int)((ec >> 8) & 0xff)),
                   u2 = (unsigned)(q.v2 - (int)((ec >> 16) & 0xff));
    const bool ok = max(max(u0, u1), u2) <= q.lim;
    const int nat = q.lin - (int)(ei & 0xffff);
    const int gi = ok ? nat : q.zero_rec + ((nat - q.zero_rec) & (RG_ZERO_RECS - 1));
    gp = q.g_base + (uint32_t)(gi * q.rec_bytes);
    g5 = q.zero_unit ? q.zero_unit : gp + 80;
    yp = q.y_base + (uint32_t)((int)(ei >> 16) * (YH_UNITS * 16));
}

// one entry; g = units 1..5 of this entry's half record (10 doubles: g[2..11]), g0 = unit 0 of this entry, g0n <- unit 0 of the next;
// yq = first y pair of this entry on entry, of the next on exit.
template <int PIPE>
__device__ __forceinline__ void taylor_half(double (&acc)[J + 1], double (&g)[10], const double2& g0, double2& g0n, double2& yq,
                                            const LaneGeoH& q, const uint2& wn, uint32_t& gp, uint32_t& g5, uint32_t& yp) {
    uint32_t gp_n = gp, g5_n = g5, yp_n = yp;
    auto G = [&](int j2) -> double { return j2 == 0 ? g0.x : (j2 == 1 ? g0.y : g[j2 - 2]); };
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const double2 t = yq;
        if (k == 0) { decode_h(q, wn, gp_n, g5_n, yp_n); }
        yq = (k < 5) ? lds128(yp + 16 * (k + 1)) : lds128(yp_n);
        if (k == 1) g0n = lds128(gp_n);
        const int j1 = 2 * k;
#pragma unroll
        for (int j2 = 0; j2 <= J - j1; ++j2) acc[j1 + j2] = fma(t.x, G(j2), acc[j1 + j2]);
        if (PIPE) {   // unit u_m (g[2m], g[2m+1]) is free once row J-2m is done: reload it with the next entry's
            const int m = (J - j1) / 2;   // row j1 = J - 2m  ->  m = 5 at j1 = 0, 4 at j1 = 2, ...
            if (m >= 1) { const double2 n = (m == 5) ? lds128(g5_n) : lds128(gp_n + 16 * m); g[2 * m - 2] = n.x; g[2 * m - 1] = n.y; }
        }
        if (j1 + 1 <= J) {
#pragma unroll
            for (int j2 = 0; j2 <= J - j1 - 1; ++j2) acc[j1 + 1 + j2] = fma(t.y, G(j2), acc[j1 + 1 + j2]);
        }
    }
    if (!PIPE) {
#pragma unroll
        for (int m = 1; m <= 5; ++m) { const double2 n = (m == 5) ? lds128(g5_n) : lds128(gp_n + 16 * m); g[2 * m - 2] = n.x; g[2 * m - 1] = n.y; }
    }
    gp = gp_n; g5 = g5_n; yp = yp_n;
}

template <int NT, int MINB, int PIPE>
__global__ void __launch_bounds__(NT, MINB) half_kernel(const uint2* __restrict__ entries, int n_entries, int iters, double* out,
                                                        long long* clk) {
    double2* G = smem_dyn;
    double2* Y = G + (M + RG_ZERO_RECS) * REC_UNITS + 1;
    for (int i = threadIdx.x; i < (M + RG_ZERO_RECS) * REC_UNITS + 1; i += blockDim.x)
        G[i] = i < M * REC_UNITS ? make_double2(1e-3 * (i % 7), -1e-3 * (i % 5)) : make_double2(0.0, 0.0);
    for (int i = threadIdx.x; i < NSUP * YH_UNITS; i += blockDim.x) Y[i] = make_double2(1e-3 * (i % 3), 1e-3 * (i % 11));
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int o = lane & 15, isB = lane >> 4;
    LaneGeoH q;
    q.v0 = 10 + (o >> 3); q.v1 = 10 + (o & 7) + (warp & 1); q.v2 = 10;
    q.lin = q.v0 * 21 + q.v1;
    q.lim = 20u; q.zero_rec = M; q.rec_bytes = REC_UNITS * 16;
    q.g_base = (uint32_t)__cvta_generic_to_shared(G) + (isB ? 0 : 80);
    q.y_base = (uint32_t)__cvta_generic_to_shared(Y) + (isB ? 96 : 0);
    q.zero_unit = isB ? (uint32_t)__cvta_generic_to_shared(G + (M + RG_ZERO_RECS) * REC_UNITS) : 0u;
    double acc[J + 1], g[10];
    double2 g0a, g0b, yq;
#pragma unroll
    for (int j = 0; j <= J; ++j) acc[j] = 0.0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const uint2* const ent = entries + (size_t)warp * n_entries;
        const int last = n_entries - 1;
        uint2 wn = ent[1];
        uint32_t gp, g5, yp;
        decode_h(q, ent[0], gp, g5, yp);
        g0a = lds128(gp);
#pragma unroll
        for (int m = 1; m <= 5; ++m) { const double2 n = (m == 5) ? lds128(g5) : lds128(gp + 16 * m); g[2 * m - 2] = n.x; g[2 * m - 1] = n.y; }
        yq = lds128(yp);
#pragma unroll 1
        for (int i = 0; i < n_entries; i += 2) {
            const uint2 w2 = ent[min(i + 2, last)];
            taylor_half<PIPE>(acc, g, g0a, g0b, yq, q, wn, gp, g5, yp);
            const uint2 w3 = ent[min(i + 3, last)];
            taylor_half<PIPE>(acc, g, g0b, g0a, yq, q, w2, gp, g5, yp);
            wn = w3;
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int j = 0; j <= J; ++j) s += acc[j];
    s += __shfl_xor_sync(0xffffffffu, s, 16);
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <typename K>
static void run(const char* name, K kern, int nthreads, size_t smem, const uint2* d_ent, int n_entries, int iters, double* d_out,
                long long* d_clk, int sms, double dfma_per_entry, double outputs_per_entry) {
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, kern));
    int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, nthreads, smem));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    kern<<<sms, nthreads, smem>>>(d_ent, n_entries, 2, d_out, d_clk);
    CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
    float best = 1e30f; long long bestclk = 0;
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(e0));
        kern<<<sms, nthreads, smem>>>(d_ent, n_entries, iters, d_out, d_clk);
        CK(cudaEventRecord(e1)); CK(cudaGetLastError()); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        std::vector<long long> clk(sms); CK(cudaMemcpy(clk.data(), d_clk, sms * 8, cudaMemcpyDeviceToHost));
        long long mx = 0; for (auto c : clk) mx = c > mx ? c : mx;
        if (ms < best) { best = ms; bestclk = mx; }
    }
    std::vector<unsigned long long> bits((size_t)sms * nthreads);
    CK(cudaMemcpy(bits.data(), d_out, bits.size() * 8, cudaMemcpyDeviceToHost));
    unsigned long long sum = 0;
    for (size_t i = 0; i < bits.size(); ++i) sum += bits[i] * (2 * i + 1);
    const int warps = nthreads / 32;
    const double entries_sm = (double)warps * n_entries * iters;        // warp-entries per SM
    const double dfma_smsp = entries_sm * dfma_per_entry / 4.0;          // warp-DFMAs per sub-partition
    const double pairs = entries_sm * outputs_per_entry;                 // (output, source) pairs per SM
    printf("{\"variant\": \"%s\", \"threads\": %d, \"regs\": %d, \"occupancy_ctas\": %d, \"ms\": %.3f, \"clk\": %lld, "
           "\"clk_per_dfma_per_smsp\": %.3f, \"clk_per_pair_per_sm\": %.4f, \"issued_tflops\": %.2f, \"checksum\": \"%016llx\"}\n",
           name, nthreads, fa.numRegs, occ, best, bestclk, bestclk / dfma_smsp, bestclk / pairs,
           2.0 * 32 * entries_sm * dfma_per_entry * sms / (best * 1e-3) * 1e-12, sum);
}

int main(int argc, char** argv) {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    const int n_entries = 240, iters = argc > 1 ? atoi(argv[1]) : 40, max_warps = 32;
    std::vector<uint2> ent((size_t)max_warps * n_entries);
    uint32_t seed = 12345u;
    for (auto& e : ent) {
        seed = seed * 1664525u + 1013904223u;
        const int c0 = 8 + (seed >> 8) % 5, c1 = 4 + (seed >> 16) % 12;           // biased coords of nu1: most lanes in range
        const int idx = (c0 - 8) * 21 + (c1 - 4) + (int)((seed >> 24) % 8) * 8;   // record index of nu1 (lin - idx stays in [0, M))
        const int rec = (seed >> 4) % NSUP;
        e.x = (uint32_t)c0 | ((uint32_t)c1 << 8) | (0u << 16) | (11u << 24);
        e.y = (uint32_t)(idx & 0xffff) | ((uint32_t)rec << 16);
    }
    uint2* d_ent; double* d_out; long long* d_clk;
    CK(cudaMalloc(&d_ent, ent.size() * sizeof(uint2))); CK(cudaMemcpy(d_ent, ent.data(), ent.size() * sizeof(uint2), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_out, (size_t)sms * 1024 * 8)); CK(cudaMalloc(&d_clk, sms * 8));
    const size_t smemF = ((size_t)(M + RG_ZERO_RECS) * REC_UNITS + (size_t)NSUP * REC_UNITS) * 16;
    const size_t smemH = ((size_t)(M + RG_ZERO_RECS) * REC_UNITS + 1 + (size_t)NSUP * YH_UNITS) * 16;
    run("full_12warps", full_kernel, 384, smemF, d_ent, n_entries, iters, d_out, d_clk, sms, 130.0, 32.0);
    run("half_pipe_12warps", half_kernel<384, 1, 1>, 384, smemH, d_ent, n_entries, iters, d_out, d_clk, sms, 66.0, 16.0);
    run("half_pipe_16warps", half_kernel<512, 1, 1>, 512, smemH, d_ent, n_entries, iters, d_out, d_clk, sms, 66.0, 16.0);
    run("half_pipe_20warps", half_kernel<640, 1, 1>, 640, smemH, d_ent, n_entries, iters, d_out, d_clk, sms, 66.0, 16.0);
    run("half_pipe_24warps", half_kernel<768, 1, 1>, 768, smemH, d_ent, n_entries, iters, d_out, d_clk, sms, 66.0, 16.0);
    run("half_nopipe_20warps", half_kernel<640, 1, 0>, 640, smemH, d_ent, n_entries, iters, d_out, d_clk, sms, 66.0, 16.0);
    run("half_nopipe_24warps", half_kernel<768, 1, 0>, 768, smemH, d_ent, n_entries, iters, d_out, d_clk, sms, 66.0, 16.0);
    run("half_nopipe_32warps", half_kernel<1024, 1, 0>, 1024, smemH, d_ent, n_entries, iters, d_out, d_clk, sms, 66.0, 16.0);
    return 0;
}
