// rg_kernel.cuh -- the fused kernel of the RG hot path: one persistent CTA (or cluster) per Hamiltonian, every loop of
// RG.converge / rg_map / expm_* on the device.  Instantiated once per Taylor order J and record placement (BIG) in its
// own translation unit (rg_kernel_inst.cu), so that the orders compile in parallel and any J in 2..12 has a kernel.
#pragma once
#include "../../include/rg_b200.h"
#include "rg_device.cuh"
#include "rg_multiply.cuh"

namespace rgdev {

struct Job {
    int mode;  // 0 region (amps), 1 converge (f given), 2 rg_map x n_iter, 3 expm, 4 liouville, 5 generate_y, 6 compute_m_s,
               // 7 approach (bisection of the pair f_in / f1_io)
    int n;
    int sched;  // 0 sparse, 1 dense (2 = sparse half-space is selected on the device)
    int csize;  // thread-block cluster size: 1, 2 or 4 CTAs per Hamiltonian (modes 0-2 only)
    int* ticket;
    // inputs
    const double* amps; const int* K_flat; int nK;   // mode 0 (K_flat: j*M + centred mode)
    const double* f_in;                               // [n][vecjl] wrapped layout
    const int* k0;                                    // may be null
    int n_iter;
    const double* y0; const double* y1; const double* y2; const double* y3;  // mode 3/4 (wrapped layout)
    const double* Omega_nu;                           // [M] centred (mode 3/4/5)
    double step; int ct;
    // outputs
    int* conv; int* count; int* err;
    double* f_out; double* y0_out; double* y1_out; double* y2_out; double* y3_out;
    long long* n_liou; long long* n_lie; long long* n_maps;
    int* m_out; int* s_out; double* norm1_out;   // mode 6
    int pnorm_p; double ms_n0, ms_ell;            // mode 6: pnorm_p > 0 -> RG.pnorm(y, p) only; else compute_m_s(y, n0, ell)
    long long* liou_out;                          // optional per-item liouville count
    double* ws_base;
    // mode 7: the pair is (f_in -> f_out, f1_io); count/err hold the first end, count1/err1 the second
    double* f1_io; int* count1; int* err1; int* levels; double reldist; int max_levels;
};

__device__ __forceinline__ void load_wrapped(const DevPlan& P, const double* __restrict__ src, double* __restrict__ dst) {
    for (int i = threadIdx.x; i < P.vecjl; i += blockDim.x) {
        const int j = i / P.M, c = i - j * P.M;
        dst[i] = src[j * P.M + P.c2w[c]];
    }
}
__device__ __forceinline__ void store_wrapped(const DevPlan& P, const double* __restrict__ src, double* __restrict__ dst) {
    for (int i = threadIdx.x; i < P.vecjl; i += blockDim.x) {
        const int j = i / P.M, c = i - j * P.M;
        dst[j * P.M + P.c2w[c]] = src[i];
    }
}

// ---- work distribution: an atomic ticket over the items in array order (persistent CTAs; cost per item varies from
// one to tens of thousands of liouville evaluations).  Results do not depend on the order in which items are processed.
__device__ __forceinline__ int fetch_item(const Job& job) { return atomicAdd(job.ticket, 1); }

// Threads and registers.  Orders J >= 8 on the shared-memory path run 8 warps of 255 registers, one CTA per SM: with the
// 168 registers of 12 warps ptxas folds the two halves of the gathered record and the next entry's operands onto the same
// registers, so their loads can only be issued a few cycles before the data is needed (11 % of the hot loop's stall
// samples were shared-memory scoreboard waits); with 255 registers the loads sit half a list entry ahead and two warps
// per sub-partition keep the fp64 pipe busier than three did (L = J = 10, uniform batch: bracket 114.2 k -> 107.2 k
// clocks, kernel -4.5 %; 4 warps: +17 %).  BIG kernels keep 12 warps x 168: their y records come through L1 and need the
// extra warps to cover it (8 warps: bracket +15 %).  Small orders: 8 warps x 128 registers, two CTAs per SM.
template <int J, bool BIG> struct KernelShape {
    static constexpr int threads = BIG ? 384 : 256;
    static constexpr int min_ctas = (J >= 8 || BIG) ? 1 : 2;
};
template <int J, bool BIG>
__global__ void __launch_bounds__(KernelShape<J, BIG>::threads, KernelShape<J, BIG>::min_ctas) rg_kernel(const int slot, const __grid_constant__ Tables T, const __grid_constant__ Job job) {
    namespace cg = cooperative_groups;
    Ctx c;
    c.slot = slot;
    c.sidx = job.sched;
    c.csize = job.csize;
    c.cbase = job.csize == 4 ? 6 : (job.csize == 2 ? 3 : 0);
    c.crank = job.csize > 1 ? (int)cg::this_cluster().block_rank() : 0;
    const DevPlan& P = RG_PLAN(c);
    int* sh_item = reinterpret_cast<int*>(smem_red() + 64);
    c.ws = job.ws_base + (size_t)blockIdx.x * (size_t)P.ws_stride;
    c.rec = reinterpret_cast<double2*>(c.ws + (size_t)P.rec_offset);
    c.rec2 = c.ws + (size_t)P.rec2_offset;
    const Smem sm = smem_view<BIG>(P, RG_SCHED(c), c);
    c.Omega_nu = T.Omega_nu;
    // zero all records once (pads and the all-zero records stay zero for the life of the CTA)
    if (BIG) {
        const int ng = (P.M + RG_ZERO_RECS) * P.g_units, ny = P.M * P.rec_units;
        for (int i = threadIdx.x; i < ng; i += blockDim.x) sm.G[i] = make_double2(0.0, 0.0);
        for (int i = threadIdx.x; i < ny; i += blockDim.x) sm.Y[i] = make_double2(0.0, 0.0);
    } else {
        const int nd2 = (P.M + RG_ZERO_RECS + RG_SCHED(c).n_support) * P.rec_units;
        for (int i = threadIdx.x; i < nd2; i += blockDim.x) sm.G[i] = make_double2(0.0, 0.0);
    }
    double* fa = c.ws + (size_t)WS_FA * P.vecjl;
    const int M = P.M, c0 = (M - 1) / 2;
    for (;;) {
        if (c.csize > 1) {
            // one ticket per cluster: rank 0 draws it and posts it into every peer's shared memory
            cg::cluster_group cl = cg::this_cluster();
            cl.sync();
            if (c.crank == 0 && threadIdx.x == 0) {
                const int it = fetch_item(job);
                for (int r = 0; r < c.csize; ++r) *cl.map_shared_rank(sh_item, r) = it;
            }
            cl.sync();
        } else {
            __syncthreads();
            if (threadIdx.x == 0) *sh_item = fetch_item(job);
            __syncthreads();
        }
        const int item = *sh_item;
        if (item >= job.n) break;
        c.n_liou = 0; c.n_liou_full = 0; c.n_lie = 0; c.budget_base = 0; c.symm = 0; c.t_conv = 0; c.t_epi = 0; c.t_busy = 0;
        const long long t_item = clock64();
        c.rng = 0x243F6A8885A308D3ull ^ ((uint64_t)(item + 1) * 0x9E3779B97F4A7C15ull);
        long long n_maps = 0;
        double* f = fa;
        const int k0 = job.k0 ? job.k0[item] : 0;
        if (job.mode == 0) {
            // generate_1Hamiltonian(amps, symmetric=True)  RG.py:255-264
            for (int i = threadIdx.x; i < P.vecjl; i += blockDim.x) f[i] = 0.0;
            __syncthreads();
            if (threadIdx.x == 0) {
                for (int q = 0; q < job.nK; ++q) f[job.K_flat[q]] = job.amps[(size_t)item * job.nK + q];
                f[2 * M + c0] = 0.5;
            }
            __syncthreads();
            sym_and_norms<J>(c, f, true);
        } else if (job.mode != 6) {
            load_wrapped(P, job.f_in + (size_t)item * P.vecjl, f);
            __syncthreads();
        }
        if (job.mode == 0 || job.mode == 1) {
            // RG.converge  RG.py:278-290
            double2 nn = sym_and_norms<J>(c, f, false);
            double ni = nn.y;
            int k = 0, err = 0;
            bool overflow = false, budget = false;
            while (P.tol_max > ni && ni > P.tol_min) {
                if (k0 + k >= T.n_tab) { overflow = true; break; }
                if (c.n_liou - c.budget_base > P.max_liouville) { budget = true; break; }
                f = rg_map_dev<J, BIG>(c, f, T, k0 + k, err, ni);
                ++k; ++n_maps;
            }
            // a Lie series or Lie loop cut short by the work budget leaves an array that means nothing: report the
            // budget, whatever the norm of that array says
            if (c.n_liou - c.budget_base > P.max_liouville) budget = true;
            if (threadIdx.x == 0 && c.crank == 0) {
                const bool cv = (ni < P.tol_min) && !overflow && !budget;
                job.conv[item] = cv ? 1 : 0;
                job.count[item] = cv ? -k : k;
                if (budget) job.err[item] = RG_ERR_WORK_BUDGET;
                else if (overflow) job.err[item] = RG_ERR_TABLE_EXHAUSTED;
                else if (!cv) job.err[item] = err;
                else if (job.mode == 0) job.err[item] = 0;
            }
        } else if (job.mode == 2) {
            int err = 0; double ni;
            for (int k = 0; k < job.n_iter; ++k) {
                if (k0 + k >= T.n_tab) { err = RG_ERR_TABLE_EXHAUSTED; break; }
                f = rg_map_dev<J, BIG>(c, f, T, k0 + k, err, ni);
                ++n_maps;
            }
            __syncthreads();
            if (c.crank == 0) {
                store_wrapped(P, f, job.f_out + (size_t)item * P.vecjl);
                if (threadIdx.x == 0) job.err[item] = err;
            }
        } else if (job.mode == 3 || job.mode == 4) {
            // caller-supplied y (wrapped) -> WS_Y1..3
            c.Omega_nu = job.Omega_nu;
            load_wrapped(P, job.y1 + (size_t)item * P.vecjl, c.ws + (size_t)WS_Y1 * P.vecjl);
            load_wrapped(P, job.y2 + (size_t)item * P.vecjl, c.ws + (size_t)WS_Y2 * P.vecjl);
            if (job.mode == 3) load_wrapped(P, job.y3 + (size_t)item * P.vecjl, c.ws + (size_t)WS_Y3 * P.vecjl);
            __syncthreads();
            const double a = job.y0[item];
            if (job.mode == 3) {
                int err = job.err[item];
                if (job.ct == 0) err = expm_onestep<J, BIG>(c, f, a, job.step);
                else if (job.ct == 1) err = expm_adapt<J, BIG>(c, f, a, job.step, err);
                else err = expm_multiply<J, BIG>(c, f, a);
                __syncthreads();
                store_wrapped(P, f, job.f_out + (size_t)item * P.vecjl);
                if (threadIdx.x == 0) job.err[item] = err;
            } else {
                // RG.liouville  RG.py:115-117
                double* dst = c.ws + (size_t)WS_T0 * P.vecjl;
                build_Y<J, BIG>(c, 1.0, 1.0, false);
                liouville_apply<J, BIG>(c, f, dst, a, false);
                __syncthreads();
                store_wrapped(P, dst, job.f_out + (size_t)item * P.vecjl);
            }
        } else if (job.mode == 6) {
            // RG.compute_m_s  RG.py:130-150
            c.Omega_nu = job.Omega_nu;
            load_wrapped(P, job.y1 + (size_t)item * P.vecjl, c.ws + (size_t)WS_Y1 * P.vecjl);
            load_wrapped(P, job.y2 + (size_t)item * P.vecjl, c.ws + (size_t)WS_Y2 * P.vecjl);
            __syncthreads();
            build_Y<J, BIG>(c, 1.0, 1.0, false);
            __syncthreads();
            int m_star = 0; double scale = 0.0, n1;
            if (job.pnorm_p > 0) {   // RG.pnorm(y, p)  RG.py:119-123
                const double e = onenormest_pow<J, BIG>(c, job.y0[item], job.pnorm_p);
                n1 = job.pnorm_p == 1 ? e : pow(e, 1.0 / (double)job.pnorm_p);
            } else {
                compute_m_s<J, BIG>(c, job.y0[item], m_star, scale, n1, job.ms_n0, job.ms_ell);
            }
            if (threadIdx.x == 0) {
                job.m_out[item] = m_star;
                job.s_out[item] = scale < 2147483647.0 ? (int)scale : 2147483647;
                job.norm1_out[item] = n1;
            }
        } else if (job.mode == 7) {
            // RG.approach (RG.py:292-307): the whole bisection of one pair, no host round trip per level
            double* a0 = c.ws + (size_t)WS_A0 * P.vecjl;
            double* a1 = c.ws + (size_t)WS_A1 * P.vecjl;
            double* am = c.ws + (size_t)WS_AM * P.vecjl;
            load_wrapped(P, job.f_in + (size_t)item * P.vecjl, a0);
            load_wrapped(P, job.f1_io + (size_t)item * P.vecjl, a1);
            __syncthreads();
            int cnt0 = job.count[item], cnt1 = job.count1[item], e1 = job.err1[item];
            int levels = 0, bad = 0;
            for (;;) {
                const double2 dn = diff_norms<J>(c, a0, a1, true);
                if (!(dn.x > job.reldist * dn.y) || levels >= job.max_levels) break;
                f = fa;
                for (int i = threadIdx.x; i < P.vecjl; i += blockDim.x) { const double m = (a0[i] + a1[i]) * 0.5; am[i] = m; f[i] = m; }
                __syncthreads();
                bool cv; int err, status;
                const int k = converge_dev<J, BIG>(c, f, T, k0, cv, err, status, n_maps);
                ++levels;
                if (status) { bad = status == 1 ? RG_ERR_TABLE_EXHAUSTED : RG_ERR_WORK_BUDGET; break; }
                __syncthreads();
                copy_arr(c, cv ? a0 : a1, am);
                if (cv) cnt0 = -k; else { cnt1 = k; e1 = err; }
                __syncthreads();
            }
            if (c.crank == 0) {
                store_wrapped(P, a0, job.f_out + (size_t)item * P.vecjl);
                store_wrapped(P, a1, job.f1_io + (size_t)item * P.vecjl);
                if (threadIdx.x == 0) {
                    job.count[item] = cnt0; job.count1[item] = cnt1;
                    job.err1[item] = bad ? bad : e1;
                    if (bad) job.err[item] = bad;
                    job.levels[item] = levels;
                }
            }
        } else if (job.mode == 5) {
            c.Omega_nu = job.Omega_nu;
            const double a = generate_y<J>(c, f);
            __syncthreads();
            store_wrapped(P, c.ws + (size_t)WS_Y1 * P.vecjl, job.y1_out + (size_t)item * P.vecjl);
            store_wrapped(P, c.ws + (size_t)WS_Y2 * P.vecjl, job.y2_out + (size_t)item * P.vecjl);
            store_wrapped(P, c.ws + (size_t)WS_Y3 * P.vecjl, job.y3_out + (size_t)item * P.vecjl);
            if (threadIdx.x == 0) job.y0_out[item] = a;
        }
        if (threadIdx.x == 0 && c.crank == 0 && job.liou_out) job.liou_out[item] = c.n_liou;
        if (threadIdx.x == 0 && c.crank == 0 && job.n_liou) {
            job.n_liou[item] = c.n_liou; job.n_lie[item] = c.n_lie; job.n_maps[item] = n_maps;
            job.n_maps[(size_t)job.n + item] = c.t_conv; job.n_maps[2 * (size_t)job.n + item] = c.t_epi;
            job.n_maps[3 * (size_t)job.n + item] = clock64() - t_item;
            job.n_maps[4 * (size_t)job.n + item] = c.t_busy;
            job.n_maps[5 * (size_t)job.n + item] = c.n_liou_full;
        }
    }
}

}  // namespace rgdev
