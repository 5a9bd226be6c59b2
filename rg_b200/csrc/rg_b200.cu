// rg_b200.cu -- kernels, host runtime and C ABI (include/rg_b200.h) of the B200-native RG hot path.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -shared -Xcompiler -fPIC
#include <cuda_runtime.h>
#include <stdint.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/rg_b200.h"
#include "rg_plan.hpp"
#include "rg_kernel.cuh"   // Job, DevPlan, Tables and the small device helpers; the fused kernel itself is instantiated elsewhere

using namespace rgdev;

// The fused kernel rg_kernel<J, BIG> is instantiated once per Taylor order J (its Taylor loops are fully unrolled) and
// record placement (BIG = operand records in global memory, for arrays beyond one SM's shared memory), each in its own
// translation unit (rg_kernel_inst.cu, compiled in parallel by the Makefile for J = 2..12).  The host side finds them
// through weak symbols: an order the Makefile did not build (`make JS="10"` for quick experiments) is simply absent.
#define RG_J_ALL(X) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12)
#define RG_DECL(j)                                                                                         \
    extern "C" const void* rg_inst_kernel_##j##_0(void) __attribute__((weak));                             \
    extern "C" const void* rg_inst_kernel_##j##_1(void) __attribute__((weak));                             \
    extern "C" int rg_inst_upload_##j##_0(const rgdev::DevPlan*, int) __attribute__((weak));               \
    extern "C" int rg_inst_upload_##j##_1(const rgdev::DevPlan*, int) __attribute__((weak));
RG_J_ALL(RG_DECL)
#undef RG_DECL
struct KernelInst { const void* kernel; int (*upload)(const rgdev::DevPlan*, int); };
static KernelInst pick_inst(int J, bool big) {
    KernelInst k{nullptr, nullptr};
    switch (J) {
#define RG_CASE(j)                                                                                         \
    case j:                                                                                                \
        if (big) { if (rg_inst_kernel_##j##_1) { k.kernel = rg_inst_kernel_##j##_1(); k.upload = rg_inst_upload_##j##_1; } } \
        else { if (rg_inst_kernel_##j##_0) { k.kernel = rg_inst_kernel_##j##_0(); k.upload = rg_inst_upload_##j##_0; } }     \
        break;
        RG_J_ALL(RG_CASE)
#undef RG_CASE
        default: break;
    }
    return k;
}

// ------------------------------------------------------------------------------------------------ errors
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CU(x)                                                                                         \
    do {                                                                                              \
        cudaError_t e_ = (x);                                                                         \
        if (e_ != cudaSuccess) {                                                                      \
            return fail((int)e_, std::string(#x) + ": " + cudaGetErrorString(e_));                   \
        }                                                                                             \
    } while (0)

extern "C" const char* rg_last_error(void) { return g_err.c_str(); }
extern "C" int rg_device_count(int* n) {
    cudaError_t e = cudaGetDeviceCount(n);
    if (e != cudaSuccess) { *n = 0; return fail((int)e, cudaGetErrorString(e)); }
    return 0;
}

// ---- small HBM-bound batched operators: one CTA per Hamiltonian, the array streamed in memory order
// (reference layout [j][wrapped nu]); the per-mode tables are in wrapped order too.  Algorithmic traffic:
// 8*vecjl bytes per Hamiltonian for a norm, 16*vecjl for sym and re-index (read + write).
// Every thread keeps RG_SMALL_BATCH independent loads in flight (16 CTAs x 128 threads x 4 x 8 B = 64 KB per SM).
#define RG_SMALL_THREADS 128
#define RG_SMALL_BATCH 4
#define RG_SMALL_STEP(w, j, M) { w += RG_SMALL_THREADS; while (w >= M) { w -= M; ++j; } }
// RG.norm / norm_int / I^- norm (RG.py:92-97,244-247,217-218).  One double out.
__global__ void __launch_bounds__(RG_SMALL_THREADS, 16) norm_kernel(const DevPlan P, int n, const double* __restrict__ f, int which, double* __restrict__ out) {
    __shared__ double red[64];
    const int type = P.norm_choice, M = P.M, V = P.vecjl;
    int j0 = 0, w0 = threadIdx.x;
    while (w0 >= M) { w0 -= M; ++j0; }
    for (int item = blockIdx.x; item < n; item += gridDim.x) {
        const double* fi = f + (size_t)item * V;
        double p = 0.0, q = 0.0;
        int j = j0, w = w0;
        for (int i = threadIdx.x; i < V; i += RG_SMALL_BATCH * RG_SMALL_THREADS) {
            double v[RG_SMALL_BATCH];
#pragma unroll
            for (int u = 0; u < RG_SMALL_BATCH; ++u) { const int k = i + u * RG_SMALL_THREADS; v[u] = k < V ? fi[k] : 0.0; }
#pragma unroll
            for (int u = 0; u < RG_SMALL_BATCH; ++u) {
                if (i + u * RG_SMALL_THREADS < V) {
                    bool use = true;
                    if (which == RG_WHICH_NORM_INT) use = (w != 0);
                    else if (which == RG_WHICH_NORM_IMINUS) use = (P.mask_w[w] >> j) & 1u;
                    if (use) norm_acc(type, p, v[u], type == 3 ? P.normw_w[w] : 1.0);
                }
                RG_SMALL_STEP(w, j, M)
            }
        }
        double2 r = block_reduce2(red, type, p, q);
        if (threadIdx.x == 0) out[item] = r.x;
    }
}
// RG.sym (RG.py:249-253)
__global__ void __launch_bounds__(RG_SMALL_THREADS, 16) sym_kernel(const DevPlan P, int n, const double* __restrict__ f, double* __restrict__ out) {
    const int M = P.M, V = P.vecjl;
    int j0 = 0, w0 = threadIdx.x;
    while (w0 >= M) { w0 -= M; ++j0; }
    for (int item = blockIdx.x; item < n; item += gridDim.x) {
        const double* fi = f + (size_t)item * V;
        double* oi = out + (size_t)item * V;
        int j = j0, w = w0;
        for (int i = threadIdx.x; i < V; i += RG_SMALL_BATCH * RG_SMALL_THREADS) {
            double a[RG_SMALL_BATCH], b[RG_SMALL_BATCH];
#pragma unroll
            for (int u = 0; u < RG_SMALL_BATCH; ++u) {
                const int k = i + u * RG_SMALL_THREADS;
                const bool ok = k < V;
                a[u] = ok ? fi[k] : 0.0;
                b[u] = ok ? fi[j * M + P.neg_w[w]] : 0.0;
                RG_SMALL_STEP(w, j, M)
            }
#pragma unroll
            for (int u = 0; u < RG_SMALL_BATCH; ++u) {
                const int k = i + u * RG_SMALL_THREADS;
                if (k < V) oi[k] = (k == 0) ? 0.0 : (a[u] + b[u]) / 2;
            }
        }
    }
}
// head of RG.rg_map (RG.py:210-216)
__global__ void __launch_bounds__(RG_SMALL_THREADS, 16) reindex_kernel(const DevPlan P, const Tables T, int n, const double* __restrict__ f, const int* __restrict__ k0, double* __restrict__ out) {
    __shared__ double ren[32];
    const int M = P.M, V = P.vecjl;
    int j0 = 0, w0 = threadIdx.x;
    while (w0 >= M) { w0 -= M; ++j0; }
    for (int item = blockIdx.x; item < n; item += gridDim.x) {
        const double* fi = f + (size_t)item * V;
        double* oi = out + (size_t)item * V;
        const int k = k0 ? k0[item] : 0;
        const double f2 = fi[2 * M];
        const double base = 2 * T.nrm[k] / P.eigenvalue * f2;
        __syncthreads();
        if (threadIdx.x <= P.J) ren[threadIdx.x] = pow(base, (double)(2 - (int)threadIdx.x)) / (2 * f2);
        int j = j0, w = w0;
        // the gathers do not depend on the rescaling factors: issue the first batch before waiting for them
        for (int i0 = 0; i0 < V; i0 += RG_SMALL_BATCH * RG_SMALL_THREADS) {   // CTA-uniform trip count (barrier inside)
            const int i = i0 + threadIdx.x;
            double a[RG_SMALL_BATCH];
            int jj[RG_SMALL_BATCH];
#pragma unroll
            for (int u = 0; u < RG_SMALL_BATCH; ++u) {
                const int kk = i + u * RG_SMALL_THREADS;
                const int s = P.reidx_w[w];
                a[u] = (kk < V && s >= 0) ? fi[j * M + s] : 0.0;
                jj[u] = j;
                RG_SMALL_STEP(w, j, M)
            }
            if (i0 == 0) __syncthreads();
#pragma unroll
            for (int u = 0; u < RG_SMALL_BATCH; ++u) {
                const int kk = i + u * RG_SMALL_THREADS;
                if (kk < V) oi[kk] = a[u] * ren[min(jj[u], 31)];
            }
        }
    }
}
// RG.conv_product (RG.py:102-106) for arbitrary operands: operands staged in shared memory (centred),
// one thread per output element.  Not on the fused path (the Lie series uses conv_all).
__global__ void __launch_bounds__(256) conv_kernel(const DevPlan P, int n, const double* __restrict__ f1, const double* __restrict__ f2, double* __restrict__ out) {
    extern __shared__ double2 sm2[];
    double* A = reinterpret_cast<double*>(sm2);
    double* B = A + P.vecjl;
    const int M = P.M, L = P.L, n1 = P.n1;
    for (int item = blockIdx.x; item < n; item += gridDim.x) {
        __syncthreads();
        for (int i = threadIdx.x; i < P.vecjl; i += blockDim.x) {
            const int j = i / M, cidx = i - j * M;
            A[i] = f1[(size_t)item * P.vecjl + j * M + P.c2w[cidx]];
            B[i] = f2[(size_t)item * P.vecjl + j * M + P.c2w[cidx]];
        }
        __syncthreads();
        for (int i = threadIdx.x; i < P.vecjl; i += blockDim.x) {
            const int j = i / M, cidx = i - j * M;
            int v[3] = {0, 0, 0};
            { int r = cidx; for (int d = P.dim - 1; d >= 0; --d) { v[d] = r % n1; r /= n1; } }
            double acc = 0.0;
            for (int j1 = 0; j1 <= j; ++j1) {
                const double* Aj = A + j1 * M; const double* Bj = B + (j - j1) * M;
                if (P.dim == 2) {
                    for (int a1 = max(0, v[0] - L); a1 <= min(n1 - 1, v[0] + L); ++a1) {
                        const int a2 = v[0] + L - a1;
                        for (int b1 = max(0, v[1] - L); b1 <= min(n1 - 1, v[1] + L); ++b1)
                            acc = fma(Aj[a1 * n1 + b1], Bj[a2 * n1 + (v[1] + L - b1)], acc);
                    }
                } else if (P.dim == 3) {
                    for (int a1 = max(0, v[0] - L); a1 <= min(n1 - 1, v[0] + L); ++a1)
                        for (int b1 = max(0, v[1] - L); b1 <= min(n1 - 1, v[1] + L); ++b1)
                            for (int c1 = max(0, v[2] - L); c1 <= min(n1 - 1, v[2] + L); ++c1)
                                acc = fma(Aj[(a1 * n1 + b1) * n1 + c1],
                                          Bj[((v[0] + L - a1) * n1 + (v[1] + L - b1)) * n1 + (v[2] + L - c1)], acc);
                } else {
                    for (int a1 = max(0, v[0] - L); a1 <= min(n1 - 1, v[0] + L); ++a1) acc = fma(Aj[a1], Bj[v[0] + L - a1], acc);
                }
            }
            out[(size_t)item * P.vecjl + j * M + P.c2w[cidx]] = acc;
        }
    }
}

// ------------------------------------------------------------------------------------------------ host runtime
struct DevSched {
    int* lane_out = nullptr; uint32_t* lane_coord = nullptr; int4* pieces = nullptr; int2* piece_runs = nullptr; uint2* entries = nullptr; int* supp_rec = nullptr;
};

struct rg_plan {
    rgplan::Geometry g;
    rgplan::Schedule hs[9];   // kind + 3*log2(csize)
    DevSched ds[9];
    int cluster_pref = 0;     // 0 = automatic, else forced cluster size for the fused modes
    int max_cluster = 1;      // largest cluster size the device can co-schedule for this kernel
    DevPlan dp;
    int device = 0;
    int slot = -1;  // index into the constant-memory plan table
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int threads = 0, sm_count = 0;
    size_t smem[2] = {0, 0};
    int ctas_per_sm[2] = {0, 0};
    int grid_max[2] = {0, 0};
    double* ws = nullptr; size_t ws_ctas = 0;
    int* ticket = nullptr;
    std::vector<void*> owned;  // device allocations
    std::vector<int> w2c;      // wrapped -> centred
    void* kernel = nullptr;
    int (*upload)(const rgdev::DevPlan*, int) = nullptr;   // into the plan table of the kernel's translation unit
    // staging arena: device memory for host-side arguments of an ABI call.  Grow-only; a call bumps through it, the
    // next call starts from the top again (calls are synchronous), so the steady state has no cudaMalloc/cudaFree.
    std::vector<std::pair<char*, size_t>> arena;   // chunks (more than one only until the next call consolidates them)
    size_t arena_used = 0;                          // bytes taken from the last chunk
    size_t arena_want = 0;                          // bytes requested by the current call (all chunks)
};

static int arena_reset(rg_plan* p) {
    if (p->arena.size() > 1 || (p->arena.size() == 1 && p->arena_want > p->arena[0].second)) {
        size_t cap = 0;
        for (auto& c : p->arena) { cap += c.second; cudaFree(c.first); }
        p->arena.clear();
        cap = std::max(cap, p->arena_want);
        char* d = nullptr;
        CU(cudaMalloc(&d, cap));
        p->arena.push_back({d, cap});
    }
    p->arena_used = 0; p->arena_want = 0;
    return 0;
}
static int arena_alloc(rg_plan* p, size_t bytes, void** out) {
    bytes = (std::max<size_t>(bytes, 8) + 255) & ~(size_t)255;
    p->arena_want += bytes;
    if (p->arena.empty() || p->arena_used + bytes > p->arena.back().second) {
        const size_t cap = std::max<size_t>(bytes, std::max<size_t>((size_t)1 << 20, p->arena.empty() ? 0 : 2 * p->arena.back().second));
        char* d = nullptr;
        CU(cudaMalloc(&d, cap));
        p->arena.push_back({d, cap});
        p->arena_used = 0;
    }
    *out = p->arena.back().first + p->arena_used;
    p->arena_used += bytes;
    return 0;
}

struct rg_tables {
    rg_plan* plan;
    Tables t;
    double* d_nrm = nullptr; double* d_onu = nullptr;
};

static bool g_slot_used[RG_MAX_PLANS] = {false};
static int alloc_slot() { for (int i = 0; i < RG_MAX_PLANS; ++i) if (!g_slot_used[i]) { g_slot_used[i] = true; return i; } return -1; }

template <typename T>
static int dev_upload(rg_plan* p, const T* host, size_t n, T** out) {
    T* d = nullptr;
    CU(cudaMalloc(&d, std::max<size_t>(n, 1) * sizeof(T)));
    if (n) CU(cudaMemcpy(d, host, n * sizeof(T), cudaMemcpyHostToDevice));
    p->owned.push_back(d);
    *out = d;
    return 0;
}

// warps per CTA: one lane per output mode -> ceil(M/32) boxes; rounded up to a multiple of 4 so that the four
// SM sub-partitions carry the same number of warps (14 warps ran 12 % slower than 16: 4+4+3+3), capped at 16
// (512 threads x 128 registers = the whole register file).
static int rg_warps_for(int M, int J, bool big = false) {
    int w = (((M + 31) / 32 + 3) / 4) * 4;
    if (w < 4) w = 4;
    const int wmax = big ? 12 : 8;   // must match rgdev::KernelShape: BIG 384 x 168 registers; else 256 x 255 (J >= 8) or 256 x 128
    if (w > wmax) w = wmax;
    const char* e = getenv("RG_B200_WARPS");
    if (e && atoi(e) >= 4 && atoi(e) <= wmax && atoi(e) % 4 == 0) w = atoi(e);
    return w;
}

// Shared memory of one CTA.  Small kernels: two-derivative records of g and y plus `out`.  BIG kernels (product-rule
// form): the g-only records when they fit (g_smem), else `out`.
static size_t smem_bytes(const rgplan::Geometry& g, int n_support, bool big, bool g_smem = false) {
    size_t b;
    if (!big) b = (size_t)(g.M + RG_ZERO_RECS + n_support) * g.rec_units * 16 + (size_t)g.M * g.JP * 8;
    else if (g_smem) b = (size_t)(g.M + RG_ZERO_RECS) * g.g_units * 16;
    else b = (size_t)g.M * g.JP * 8;
    b += 64 * 8 + 16;            // reduction scratch + work ticket
    b += RG_CTL_UNITS * 16;      // the convolution's control block
    return b;
}

// Does the Hamiltonian need the BIG kernel?  (records + out + scratch beyond the opt-in shared memory of an SM)
static const size_t RG_SMEM_OPTIN = 232448;   // sm_100: 227 KB per CTA
static bool needs_global_records(const rgplan::Geometry& g, const std::vector<uint16_t>& mask, size_t optin) {
    const char* e = getenv("RG_B200_FORCE_GLOBAL");
    if (e && atoi(e)) return true;
    int n_support = 0;
    for (int c = 0; c < g.M; ++c) n_support += mask[c] != 0;
    return smem_bytes(g, n_support, false) > optin;
}

static std::vector<uint16_t> masks_from_iminus(const rgplan::Geometry& g, const uint8_t* iminus_wrapped) {
    std::vector<uint16_t> mask(g.M, 0);
    for (int c = 0; c < g.M; ++c) {
        int w = g.wrapped_of_centred(c);
        uint16_t m = 0;
        for (int j = 0; j <= g.J; ++j) if (iminus_wrapped[(size_t)j * g.M + w]) m |= (uint16_t)(1u << j);
        mask[c] = m;
    }
    return mask;
}

extern "C" int rg_plan_create(const rg_params* prm, rg_plan** out) {
    if (!prm || !out) return fail(-1, "null argument");
    if (prm->dim < 1 || prm->dim > 3) return fail(-2, "dim must be 1, 2 or 3");
    if (prm->J < 2 || prm->J > 12) return fail(-2, "J must be in 2..12 (RG.py:61 uses f[2]; beyond 12 the 2(J+1) accumulators of a lane do not fit the register file)");
    if (prm->L < 1 || prm->L > 60) return fail(-2, "L must be in 1..60");
    if (prm->ct < 0 || prm->ct > 2) return fail(-2, "unknown canonical transformation");
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (prm->device < 0 || prm->device >= ndev) return fail(-4, "no such CUDA device");
    CU(cudaSetDevice(prm->device));
    rg_plan* p = new rg_plan();
    p->device = prm->device;
    p->g.init(prm->dim, prm->L, prm->J);
    const rgplan::Geometry& g = p->g;
    std::vector<uint16_t> mask = masks_from_iminus(g, prm->iminus);
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, p->device));
    const bool big = needs_global_records(g, mask, (size_t)prop.sharedMemPerBlockOptin);
    // BIG kernels: the g-only records of the product-rule form go to shared memory when they fit one SM
    const char* gse = getenv("RG_B200_G_SMEM");
    const bool g_smem = big && (!gse || atoi(gse)) && smem_bytes(g, 0, true, true) <= (size_t)prop.sharedMemPerBlockOptin;
    const KernelInst inst = pick_inst(prm->J, big);
    if (!inst.kernel) {
        delete p;
        return fail(-3, std::string("this build of librg_b200.so has no ") + (big ? "global-record " : "") + "kernel for J=" +
                            std::to_string(prm->J) + " (rg_b200/csrc/Makefile: JS lists the Taylor orders to instantiate)");
    }
    const int n_warps = rg_warps_for(g.M, g.J, big);
    for (int ci = 0; ci < 3; ++ci) {
        int vw = n_warps << ci;   // virtual warps of a cluster of 1, 2, 4 CTAs
        p->hs[3 * ci + 0] = rgplan::make_schedule(g, mask, false, false, vw, 1 << ci, big);
        p->hs[3 * ci + 1] = rgplan::make_schedule(g, mask, true, false, vw, 1 << ci, big);
        p->hs[3 * ci + 2] = rgplan::make_schedule(g, mask, false, true, vw, 1 << ci, big);
    }
    p->sm_count = prop.multiProcessorCount;
    p->threads = n_warps * 32;
    for (int s = 0; s < 9; ++s)
        if (p->hs[s].n_warps != (n_warps << (s / 3)) || p->hs[s].n_boxes > RG_MAX_BOXES) {
            delete p;
            return fail(-5, "too many output modes for one CTA's convolution schedule");
        }
    DevPlan& dp = p->dp;
    memset(&dp, 0, sizeof(dp));
    dp.dim = g.dim; dp.L = g.L; dp.J = g.J; dp.JP = g.JP; dp.n1 = g.n1; dp.M = g.M; dp.vecjl = g.vecjl;
    dp.off2 = g.off2; dp.rec_units = g.rec_units;
    dp.g_units = big ? g.g_units : g.rec_units; dp.g_smem = g_smem ? 1 : 0;
    dp.idx_y1J = (g.J & 1) ? g.J : g.off2 + g.JP;
    dp.stride0 = g.stride[0]; dp.stride1 = g.stride[1]; dp.stride2 = g.stride[2];
    dp.norm_choice = prm->norm_choice; dp.norm_analytic = prm->norm_analytic;
    dp.tol_min = prm->tol_min; dp.tol_max = prm->tol_max; dp.max_lie = prm->max_lie; dp.ct = prm->ct;
    dp.min_step = prm->min_step; dp.abs_tol = prm->abs_tol; dp.rel_tol = prm->rel_tol; dp.eigenvalue = prm->eigenvalue;
    dp.max_series = prm->max_series > 0 ? prm->max_series : 100000;
    dp.max_liouville = prm->max_liouville > 0 ? prm->max_liouville : 2000000LL;
    dp.rec_global = big ? 1 : 0;
    {   // mirror shortcuts of generate_y / build_Y need exactly (anti)symmetric tables
        bool ok = true;
        for (int c = 0; c < g.M && ok; ++c) {
            const int m = g.M - 1 - c;
            ok = mask[c] == mask[m] && prm->omega0_nu[g.wrapped_of_centred(c)] == -prm->omega0_nu[g.wrapped_of_centred(m)];
        }
        dp.sym_tables = ok ? 1 : 0;
    }
    dp.rec_offset = ((long long)WS_COUNT * g.vecjl + 1) & ~1LL;   // records are read as double2: keep them 16-byte aligned
    // BIG kernels: [Y records: M x rec_units] then [g records: (M + 8) x g_units, or `out`: M x JP doubles when the g records are on chip]
    dp.rec2_offset = dp.rec_offset + (big ? (long long)g.M * g.rec_units * 2 : 0LL);
    dp.ws_stride = dp.rec2_offset + (big ? (g_smem ? (((long long)g.M * g.JP + 1) & ~1LL) : (long long)(g.M + RG_ZERO_RECS) * g.g_units * 2) : 0LL);
    dp.debug = getenv("RG_B200_DEBUG") ? atoi(getenv("RG_B200_DEBUG")) : 0;
    // centred tables
    std::vector<double> om0(g.M), normw(g.M);
    std::vector<int> reidx(g.M), c2w(g.M);
    p->w2c.resize(g.M);
    for (int c = 0; c < g.M; ++c) {
        int w = g.wrapped_of_centred(c);
        c2w[c] = w; p->w2c[w] = c;
    }
    for (int c = 0; c < g.M; ++c) {
        int w = c2w[c];
        om0[c] = prm->omega0_nu[w];
        int src = prm->reidx[w];
        reidx[c] = src >= 0 ? p->w2c[src] : -1;
        int v[4]; g.coords(c, v);
        int l1 = 0; for (int d = 0; d < g.dim; ++d) l1 += std::abs(v[d] - g.L);
        normw[c] = std::exp(prm->norm_analytic * (double)l1);
    }
    // the same tables in wrapped (memory) order for the stand-alone streaming operators
    std::vector<uint16_t> mask_w(g.M);
    std::vector<double> normw_w(g.M);
    std::vector<int> neg_w(g.M), reidx_w(g.M);
    for (int c = 0; c < g.M; ++c) {
        const int w = c2w[c];
        mask_w[w] = mask[c]; normw_w[w] = normw[c];
        neg_w[w] = c2w[g.M - 1 - c];
        reidx_w[w] = prm->reidx[w];
    }
    int rc;
    uint16_t* d_mask_w; double* d_normw_w; int* d_neg_w; int* d_reidx_w;
    if ((rc = dev_upload(p, mask_w.data(), mask_w.size(), &d_mask_w))) return rc;
    if ((rc = dev_upload(p, normw_w.data(), normw_w.size(), &d_normw_w))) return rc;
    if ((rc = dev_upload(p, neg_w.data(), neg_w.size(), &d_neg_w))) return rc;
    if ((rc = dev_upload(p, reidx_w.data(), reidx_w.size(), &d_reidx_w))) return rc;
    dp.mask_w = d_mask_w; dp.normw_w = d_normw_w; dp.neg_w = d_neg_w; dp.reidx_w = d_reidx_w;
    double* d_om0; uint16_t* d_mask; int* d_reidx; double* d_normw; int* d_c2w; double* d_theta;
    if ((rc = dev_upload(p, om0.data(), om0.size(), &d_om0))) return rc;
    if ((rc = dev_upload(p, mask.data(), mask.size(), &d_mask))) return rc;
    if ((rc = dev_upload(p, reidx.data(), reidx.size(), &d_reidx))) return rc;
    if ((rc = dev_upload(p, normw.data(), normw.size(), &d_normw))) return rc;
    if ((rc = dev_upload(p, c2w.data(), c2w.size(), &d_c2w))) return rc;
    if ((rc = dev_upload(p, prm->theta, (size_t)prm->n_theta, &d_theta))) return rc;
    dp.omega0_nu = d_om0; dp.mask = d_mask; dp.reidx = d_reidx; dp.normw = d_normw; dp.c2w = d_c2w;
    dp.theta = d_theta; dp.n_theta = prm->n_theta;
    for (int s = 0; s < 9; ++s) {
        const rgplan::Schedule& hs = p->hs[s];
        DevSched& ds = p->ds[s];
        if ((rc = dev_upload(p, hs.lane_out.data(), hs.lane_out.size(), &ds.lane_out))) return rc;
        if ((rc = dev_upload(p, hs.lane_coord.data(), hs.lane_coord.size(), &ds.lane_coord))) return rc;
        if ((rc = dev_upload(p, reinterpret_cast<const int4*>(hs.pieces.data()), hs.pieces.size(), &ds.pieces))) return rc;
        if ((rc = dev_upload(p, reinterpret_cast<const int2*>(hs.piece_runs.data()), hs.piece_runs.size() / 2, &ds.piece_runs))) return rc;
        if ((rc = dev_upload(p, reinterpret_cast<const uint2*>(hs.entries.data()), hs.entries.size() / 2, &ds.entries))) return rc;
        if ((rc = dev_upload(p, hs.supp_rec.data(), hs.supp_rec.size(), &ds.supp_rec))) return rc;
        dp.sched[s].n_warps = hs.n_warps; dp.sched[s].n_support = hs.n_support;
        dp.sched[s].maxp = hs.maxp;
        dp.sched[s].general = hs.general ? 1 : 0;
        dp.sched[s].n_boxes = hs.n_boxes; dp.sched[s].half = hs.half ? 1 : 0; dp.sched[s].n_stages = hs.n_stages;
        for (int r = 0; r < 4; ++r) {
            dp.sched[s].fin_begin[r] = hs.fin_begin[r]; dp.sched[s].fin_end[r] = hs.fin_end[r];
            dp.sched[s].stage0_pieces[r] = hs.stage0_pieces[r];
        }
        dp.sched[s].lane_out = ds.lane_out; dp.sched[s].lane_coord = ds.lane_coord; dp.sched[s].pieces = ds.pieces; dp.sched[s].piece_runs = ds.piece_runs;
        dp.sched[s].entries = ds.entries; dp.sched[s].supp_rec = ds.supp_rec;
    }
    p->kernel = (void*)inst.kernel;
    p->upload = inst.upload;
    for (int s = 0; s < 2; ++s) {
        p->smem[s] = smem_bytes(g, p->hs[s].n_support, big, g_smem);
        if (p->smem[s] > (size_t)prop.sharedMemPerBlockOptin) {
            if (s == 0) { rg_plan_destroy(p); return fail(-5, "Hamiltonian too large for one CTA's shared memory (" + std::to_string(p->smem[s]) + " B)"); }
            p->ctas_per_sm[s] = 0; p->grid_max[s] = 0;
            continue;
        }
    }
    size_t smax = std::max(p->smem[0], p->smem[1] <= (size_t)prop.sharedMemPerBlockOptin ? p->smem[1] : 0);
    CU(cudaFuncSetAttribute(p->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smax));
    for (int s = 0; s < 2; ++s) {
        if (p->smem[s] > (size_t)prop.sharedMemPerBlockOptin) continue;
        int nb = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, p->kernel, p->threads, p->smem[s]));
        if (nb < 1) { rg_plan_destroy(p); return fail(-6, "kernel does not fit on an SM"); }
        p->ctas_per_sm[s] = nb;
        p->grid_max[s] = nb * p->sm_count;
    }
    p->slot = alloc_slot();
    if (p->slot < 0) { rg_plan_destroy(p); return fail(-8, "too many live plans (max " + std::to_string(RG_MAX_PLANS) + " per process)"); }
    CU((cudaError_t)p->upload(&p->dp, p->slot));
    // largest cluster the device co-schedules with this kernel's footprint
    p->max_cluster = 1;
    for (int cs = 2; cs <= 4; cs *= 2) {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(cs * p->sm_count); cfg.blockDim = dim3(p->threads); cfg.dynamicSmemBytes = p->smem[0];
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = cs; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
        cfg.attrs = &attr; cfg.numAttrs = 1;
        int ncl = 0;
        if (cudaOccupancyMaxActiveClusters(&ncl, p->kernel, &cfg) == cudaSuccess && ncl * cs * 10 >= p->grid_max[0] * 8) p->max_cluster = cs;
        else { cudaGetLastError(); break; }
    }
    if (getenv("RG_B200_CLUSTER")) p->cluster_pref = atoi(getenv("RG_B200_CLUSTER"));
    p->ws_ctas = (size_t)std::max(p->grid_max[0], p->grid_max[1]);
    CU(cudaMalloc(&p->ws, p->ws_ctas * (size_t)dp.ws_stride * sizeof(double)));
    CU(cudaMemset(p->ws, 0, p->ws_ctas * (size_t)dp.ws_stride * sizeof(double)));
    CU(cudaMalloc(&p->ticket, sizeof(int)));
    CU(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
    CU(cudaEventCreate(&p->ev0));
    CU(cudaEventCreate(&p->ev1));
    // conv_kernel needs 2*vecjl doubles of dynamic shared memory
    if ((size_t)2 * g.vecjl * 8 <= (size_t)prop.sharedMemPerBlockOptin)
        CU(cudaFuncSetAttribute(conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * g.vecjl * 8));
    // the uploads and the memset above went through the legacy default stream; the kernels run on p->stream
    // (non-blocking): order them once, here
    CU(cudaDeviceSynchronize());
    *out = p;
    return 0;
}

extern "C" int rg_plan_destroy(rg_plan* p) {
    if (!p) return 0;
    cudaSetDevice(p->device);
    for (void* d : p->owned) cudaFree(d);
    for (auto& c : p->arena) cudaFree(c.first);
    if (p->ws) cudaFree(p->ws);
    if (p->ticket) cudaFree(p->ticket);
    if (p->stream) cudaStreamDestroy(p->stream);
    if (p->ev0) cudaEventDestroy(p->ev0);
    if (p->ev1) cudaEventDestroy(p->ev1);
    if (p->slot >= 0) g_slot_used[p->slot] = false;
    delete p;
    return 0;
}

extern "C" int rg_plan_info(const rg_plan* p, int32_t* threads, int32_t* smem, int32_t* ctas_per_sm, int32_t* n_support,
                            int64_t* fma_useful, int64_t* fma_executed, int32_t* resident_ctas) {
    if (!p) return fail(-1, "null plan");
    if (threads) *threads = p->threads;
    if (smem) *smem = (int32_t)p->smem[0];
    if (ctas_per_sm) *ctas_per_sm = p->ctas_per_sm[0];
    if (n_support) *n_support = p->hs[0].n_support;
    // the fused path runs the half-space schedule (operands are exactly symmetric after RG.sym)
    if (fma_useful) *fma_useful = p->hs[2].fma_useful;
    if (fma_executed) *fma_executed = p->hs[2].fma_executed;
    if (resident_ctas) *resident_ctas = p->grid_max[0];
    return 0;
}

extern "C" int rg_tables_create(rg_plan* p, int32_t n_tab, const double* nrm_tab, const double* Omega_nu_tab, rg_tables** out) {
    if (!p || !nrm_tab || !Omega_nu_tab || !out || n_tab < 1) return fail(-1, "bad argument");
    CU(cudaSetDevice(p->device));
    rg_tables* t = new rg_tables();
    t->plan = p;
    const int M = p->g.M;
    std::vector<double> onu((size_t)(n_tab + 1) * M);
    for (int k = 0; k <= n_tab; ++k)
        for (int c = 0; c < M; ++c) onu[(size_t)k * M + c] = Omega_nu_tab[(size_t)k * M + p->g.wrapped_of_centred(c)];
    CU(cudaMalloc(&t->d_nrm, sizeof(double) * n_tab));
    CU(cudaMalloc(&t->d_onu, sizeof(double) * onu.size()));
    CU(cudaMemcpy(t->d_nrm, nrm_tab, sizeof(double) * n_tab, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(t->d_onu, onu.data(), sizeof(double) * onu.size(), cudaMemcpyHostToDevice));
    t->t.n_tab = n_tab; t->t.nrm = t->d_nrm; t->t.Omega_nu = t->d_onu;
    CU(cudaDeviceSynchronize());   // legacy-stream uploads before any launch on the plan's stream
    *out = t;
    return 0;
}
extern "C" int rg_tables_destroy(rg_tables* t) {
    if (!t) return 0;
    cudaSetDevice(t->plan->device);
    cudaFree(t->d_nrm); cudaFree(t->d_onu);
    delete t;
    return 0;
}

// ---- staging: accept host or device pointers
static bool is_device_ptr(const void* p) {
    if (!p) return false;
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

struct Stage {
    rg_plan* p;
    struct Back { void* host; void* dev; size_t bytes; };
    std::vector<Back> backs;
    int rc0;
    explicit Stage(rg_plan* p_) : p(p_) { rc0 = arena_reset(p); }
    // device scratch of `bytes` from the plan's arena (lives until the next ABI call on this plan)
    int dev(size_t bytes, void** out) { if (rc0) return rc0; return arena_alloc(p, bytes, out); }
    // input: returns device pointer holding `bytes` of src
    int in(const void* src, size_t bytes, const void** out) {
        if (!src) { *out = nullptr; return 0; }
        if (is_device_ptr(src)) { *out = src; return 0; }
        void* d = nullptr;
        int rc = dev(bytes, &d);
        if (rc) return rc;
        CU(cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, p->stream));
        *out = d;
        return 0;
    }
    // output (optionally initialised from the host buffer: in/out arrays)
    int outp(void* dst, size_t bytes, void** out, bool copy_in = false) {
        if (!dst) { *out = nullptr; return 0; }
        if (is_device_ptr(dst)) { *out = dst; return 0; }
        void* d = nullptr;
        int rc = dev(bytes, &d);
        if (rc) return rc;
        if (copy_in) CU(cudaMemcpyAsync(d, dst, bytes, cudaMemcpyHostToDevice, p->stream));
        else CU(cudaMemsetAsync(d, 0, bytes, p->stream));
        backs.push_back(Back{dst, d, bytes});
        *out = d;
        return 0;
    }
    int finish() {
        for (auto& b : backs) CU(cudaMemcpyAsync(b.host, b.dev, b.bytes, cudaMemcpyDeviceToHost, p->stream));
        CU(cudaStreamSynchronize(p->stream));
        return 0;
    }
};

static int launch_job(rg_plan* p, const Tables& T, Job& job, rg_stats* stats, Stage& st) {
    const int s = job.sched;
    if (p->grid_max[s] < 1) return fail(-5, "this schedule does not fit in shared memory for the plan's (L, J)");
    const rgplan::Geometry& g = p->g;
    long long* d_cnt = nullptr;
    if (stats) {
        { void* v; int rc = st.dev(sizeof(long long) * 8 * (size_t)std::max(job.n, 1), &v); if (rc) return rc; d_cnt = (long long*)v; }
        CU(cudaMemsetAsync(d_cnt, 0, sizeof(long long) * 8 * (size_t)std::max(job.n, 1), p->stream));
        job.n_liou = d_cnt; job.n_lie = d_cnt + job.n; job.n_maps = d_cnt + 2 * (size_t)job.n;
    }
    CU(cudaMemsetAsync(p->ticket, 0, sizeof(int), p->stream));
    job.ticket = p->ticket;
    job.ws_base = p->ws;
    // Cluster size: a Hamiltonian owned by a pair (quad) of CTAs finishes 1.46x (1.9x) sooner but at 0.73 (0.48)
    // of the per-SM throughput (the epilogue is replicated, whole boxes balance to 4:3), so clusters are used
    // only when there are fewer Hamiltonians than SMs: compute_line, approach, small surfaces.
    int csize = 1;
    if ((job.mode <= 2 || job.mode == 7) && p->max_cluster > 1) {
        if (p->cluster_pref > 0) csize = std::min(p->cluster_pref, p->max_cluster);
        else if ((long long)job.n * 4 <= p->grid_max[s]) csize = std::min(4, p->max_cluster);
        else if ((long long)job.n * 2 <= p->grid_max[s]) csize = 2;
    }
    job.csize = csize;
    int grid = std::max(1, std::min(job.n, p->grid_max[s] / csize)) * csize;
    CU(cudaEventRecord(p->ev0, p->stream));
    {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(p->threads); cfg.dynamicSmemBytes = p->smem[s]; cfg.stream = p->stream;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = csize; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
        cfg.attrs = &attr; cfg.numAttrs = csize > 1 ? 1 : 0;
        int slot = p->slot;
        Tables Tc = T;
        void* args[3] = {&slot, &Tc, &job};
        CU(cudaLaunchKernelExC(&cfg, p->kernel, args));
    }
    CU(cudaGetLastError());
    CU(cudaEventRecord(p->ev1, p->stream));
    if (stats) {
        std::vector<long long> h(8 * (size_t)std::max(job.n, 1));
        CU(cudaMemcpyAsync(h.data(), d_cnt, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost, p->stream));
        CU(cudaStreamSynchronize(p->stream));
        memset(stats, 0, sizeof(*stats));
        for (int i = 0; i < job.n; ++i) {
            stats->liouville += h[i]; stats->lie_steps += h[job.n + i]; stats->rg_maps += h[2 * (size_t)job.n + i];
            stats->clk_conv += h[3 * (size_t)job.n + i]; stats->clk_epilogue += h[4 * (size_t)job.n + i];
            stats->clk_total += h[5 * (size_t)job.n + i];
            stats->clk_conv_busy += h[6 * (size_t)job.n + i];
            stats->liouville_full += h[7 * (size_t)job.n + i];
        }
        // fused modes run the half-space schedule for the series terms (exactly symmetric operands) and the full one
        // for everything else (the 1-norm estimator's probes of expm_multiply, asymmetric inputs): each bracket is
        // accounted at the cost of the schedule it ran on
        if (s == 0) {
            const long long half = stats->liouville - stats->liouville_full;
            stats->conv_fma_executed = half * p->hs[2].fma_executed + stats->liouville_full * p->hs[0].fma_executed;
            stats->conv_fma_useful = half * p->hs[2].fma_useful + stats->liouville_full * p->hs[0].fma_useful;
        } else {
            stats->conv_fma_executed = stats->liouville * p->hs[s].fma_executed;
            stats->conv_fma_useful = stats->liouville * p->hs[s].fma_useful;
        }
        CU(cudaEventElapsedTime(&stats->kernel_ms, p->ev0, p->ev1));
        stats->launches = 1;
    }
    (void)g;
    return 0;
}

static Tables null_tables() { Tables t; t.n_tab = 0; t.nrm = nullptr; t.Omega_nu = nullptr; return t; }

static int launch_job(rg_plan* p, const Tables& T, Job& job, rg_stats* stats, Stage& st);

extern "C" int rg_region(rg_plan* p, const rg_tables* tab, int32_t n, const double* amps, const int32_t* K_flat, int32_t nK,
                         int32_t* conv, int32_t* count, int32_t* err, rg_stats* stats) {
    return rg_region_work(p, tab, n, amps, K_flat, nK, conv, count, err, nullptr, stats);
}

extern "C" int rg_region_work(rg_plan* p, const rg_tables* tab, int32_t n, const double* amps, const int32_t* K_flat, int32_t nK,
                              int32_t* conv, int32_t* count, int32_t* err, int64_t* liouville_per_point, rg_stats* stats) {
    if (!p || !tab || n < 0 || !amps || !K_flat || nK < 1 || !conv || !count || !err) return fail(-1, "bad argument");
    if (p->dp.ct == RG_CT_MULTIPLY && !RG_HAVE_MULTIPLY) return fail(-7, "expm_multiply not available in this build");
    CU(cudaSetDevice(p->device));
    Stage st(p);
    Job job; memset(&job, 0, sizeof(job));
    job.mode = 0; job.n = n; job.sched = 0; job.nK = nK;
    // K_flat arrives as flat wrapped indices j*M + w; convert to centred
    std::vector<int> kf(nK);
    for (int q = 0; q < nK; ++q) { int j = K_flat[q] / p->g.M, w = K_flat[q] % p->g.M; kf[q] = j * p->g.M + p->w2c[w]; }
    int rc;
    const void* d;
    int* d_kf = nullptr;
    { void* v; if ((rc = st.dev(sizeof(int) * nK, &v))) return rc; d_kf = (int*)v; }
    CU(cudaMemcpyAsync(d_kf, kf.data(), sizeof(int) * nK, cudaMemcpyHostToDevice, p->stream));
    job.K_flat = d_kf;
    if ((rc = st.in(amps, sizeof(double) * (size_t)n * nK, &d))) return rc; job.amps = (const double*)d;
    void* o;
    if ((rc = st.outp(conv, sizeof(int) * (size_t)n, &o))) return rc; job.conv = (int*)o;
    if ((rc = st.outp(count, sizeof(int) * (size_t)n, &o))) return rc; job.count = (int*)o;
    if ((rc = st.outp(err, sizeof(int) * (size_t)n, &o))) return rc; job.err = (int*)o;
    if (liouville_per_point) { if ((rc = st.outp(liouville_per_point, sizeof(long long) * (size_t)n, &o))) return rc; job.liou_out = (long long*)o; }
    if ((rc = launch_job(p, tab->t, job, stats, st))) return rc;
    return st.finish();
}

extern "C" int rg_converge(rg_plan* p, const rg_tables* tab, int32_t n, const double* f, const int32_t* k0,
                           int32_t* conv, int32_t* count, int32_t* err_io, rg_stats* stats) {
    if (!p || !tab || n < 0 || !f || !conv || !count || !err_io) return fail(-1, "bad argument");
    if (p->dp.ct == RG_CT_MULTIPLY && !RG_HAVE_MULTIPLY) return fail(-7, "expm_multiply not available in this build");
    CU(cudaSetDevice(p->device));
    Stage st(p);
    Job job; memset(&job, 0, sizeof(job));
    job.mode = 1; job.n = n; job.sched = 0;
    int rc; const void* d; void* o;
    if ((rc = st.in(f, sizeof(double) * (size_t)n * p->g.vecjl, &d))) return rc; job.f_in = (const double*)d;
    if ((rc = st.in(k0, sizeof(int) * (size_t)n, &d))) return rc; job.k0 = (const int*)d;
    if ((rc = st.outp(conv, sizeof(int) * (size_t)n, &o))) return rc; job.conv = (int*)o;
    if ((rc = st.outp(count, sizeof(int) * (size_t)n, &o))) return rc; job.count = (int*)o;
    if ((rc = st.outp(err_io, sizeof(int) * (size_t)n, &o, true))) return rc; job.err = (int*)o;
    if ((rc = launch_job(p, tab->t, job, stats, st))) return rc;
    return st.finish();
}

extern "C" int rg_approach(rg_plan* p, const rg_tables* tab, int32_t n, double* f0, double* f1, double reldist,
                           int32_t* count0, int32_t* count1, int32_t* err0, int32_t* err1, int32_t* levels, rg_stats* stats) {
    if (!p || !tab || n < 0 || !f0 || !f1 || !count0 || !count1 || !err0 || !err1 || !levels) return fail(-1, "bad argument");
    if (p->dp.ct == RG_CT_MULTIPLY && !RG_HAVE_MULTIPLY) return fail(-7, "expm_multiply not available in this build");
    CU(cudaSetDevice(p->device));
    Stage st(p);
    Job job; memset(&job, 0, sizeof(job));
    job.mode = 7; job.n = n; job.sched = 0; job.reldist = reldist; job.max_levels = 4096;
    int rc; void* o;
    const size_t vb = sizeof(double) * (size_t)n * p->g.vecjl, ib = sizeof(int) * (size_t)n;
    if ((rc = st.outp(f0, vb, &o, true))) return rc; job.f_in = (const double*)o; job.f_out = (double*)o;
    if ((rc = st.outp(f1, vb, &o, true))) return rc; job.f1_io = (double*)o;
    if ((rc = st.outp(count0, ib, &o, true))) return rc; job.count = (int*)o;
    if ((rc = st.outp(count1, ib, &o, true))) return rc; job.count1 = (int*)o;
    if ((rc = st.outp(err0, ib, &o, true))) return rc; job.err = (int*)o;
    if ((rc = st.outp(err1, ib, &o, true))) return rc; job.err1 = (int*)o;
    if ((rc = st.outp(levels, ib, &o))) return rc; job.levels = (int*)o;
    if ((rc = launch_job(p, tab->t, job, stats, st))) return rc;
    return st.finish();
}

extern "C" int rg_map(rg_plan* p, const rg_tables* tab, int32_t n, const double* f, const int32_t* k0, int32_t n_iter,
                      double* f_out, int32_t* err, rg_stats* stats) {
    if (!p || !tab || n < 0 || !f || !f_out || !err || n_iter < 0) return fail(-1, "bad argument");
    if (p->dp.ct == RG_CT_MULTIPLY && !RG_HAVE_MULTIPLY) return fail(-7, "expm_multiply not available in this build");
    CU(cudaSetDevice(p->device));
    Stage st(p);
    Job job; memset(&job, 0, sizeof(job));
    job.mode = 2; job.n = n; job.sched = 0; job.n_iter = n_iter;
    int rc; const void* d; void* o;
    if ((rc = st.in(f, sizeof(double) * (size_t)n * p->g.vecjl, &d))) return rc; job.f_in = (const double*)d;
    if ((rc = st.in(k0, sizeof(int) * (size_t)n, &d))) return rc; job.k0 = (const int*)d;
    if ((rc = st.outp(f_out, sizeof(double) * (size_t)n * p->g.vecjl, &o))) return rc; job.f_out = (double*)o;
    if ((rc = st.outp(err, sizeof(int) * (size_t)n, &o))) return rc; job.err = (int*)o;
    if ((rc = launch_job(p, tab->t, job, stats, st))) return rc;
    return st.finish();
}

// Omega_nu arrives in wrapped order on the host or device; the kernels want it centred on the device
static int stage_Omega_nu(rg_plan* p, Stage& st, const double* Omega_nu, const double** out) {
    const int M = p->g.M;
    std::vector<double> h(M), c(M);
    if (is_device_ptr(Omega_nu)) CU(cudaMemcpy(h.data(), Omega_nu, sizeof(double) * M, cudaMemcpyDeviceToHost));
    else memcpy(h.data(), Omega_nu, sizeof(double) * M);
    for (int i = 0; i < M; ++i) c[i] = h[p->g.wrapped_of_centred(i)];
    void* d = nullptr;
    int rc = st.dev(sizeof(double) * M, &d);
    if (rc) return rc;
    CU(cudaMemcpyAsync(d, c.data(), sizeof(double) * M, cudaMemcpyHostToDevice, p->stream));
    CU(cudaStreamSynchronize(p->stream));   // `c` is a local: the copy must have left it
    *out = (const double*)d;
    return 0;
}

extern "C" int rg_expm(rg_plan* p, int32_t ct, int32_t n, const double* f, const double* y0, const double* y1,
                       const double* y2, const double* y3, const double* Omega_nu, double step, double* f_out,
                       int32_t* err_io, rg_stats* stats) {
    if (!p || n < 0 || !f || !y0 || !y1 || !y2 || !y3 || !Omega_nu || !f_out || !err_io) return fail(-1, "bad argument");
    if (ct < 0 || ct > 2) return fail(-2, "unknown canonical transformation");
    if (ct == RG_CT_MULTIPLY && !RG_HAVE_MULTIPLY) return fail(-7, "expm_multiply not available in this build");
    CU(cudaSetDevice(p->device));
    Stage st(p);
    Job job; memset(&job, 0, sizeof(job));
    job.mode = 3; job.n = n; job.sched = 1; job.ct = ct; job.step = step;
    const size_t vb = sizeof(double) * (size_t)n * p->g.vecjl;
    int rc; const void* d; void* o;
    if ((rc = st.in(f, vb, &d))) return rc; job.f_in = (const double*)d;
    if ((rc = st.in(y0, sizeof(double) * (size_t)n, &d))) return rc; job.y0 = (const double*)d;
    if ((rc = st.in(y1, vb, &d))) return rc; job.y1 = (const double*)d;
    if ((rc = st.in(y2, vb, &d))) return rc; job.y2 = (const double*)d;
    if ((rc = st.in(y3, vb, &d))) return rc; job.y3 = (const double*)d;
    if ((rc = stage_Omega_nu(p, st, Omega_nu, &job.Omega_nu))) return rc;
    if ((rc = st.outp(f_out, vb, &o))) return rc; job.f_out = (double*)o;
    if ((rc = st.outp(err_io, sizeof(int) * (size_t)n, &o, true))) return rc; job.err = (int*)o;
    if ((rc = launch_job(p, null_tables(), job, stats, st))) return rc;
    return st.finish();
}

extern "C" int rg_liouville(rg_plan* p, int32_t n, const double* y0, const double* y1, const double* y2, const double* f,
                            const double* Omega_nu, double* out) {
    if (!p || n < 0 || !f || !y0 || !y1 || !y2 || !Omega_nu || !out) return fail(-1, "bad argument");
    CU(cudaSetDevice(p->device));
    Stage st(p);
    Job job; memset(&job, 0, sizeof(job));
    job.mode = 4; job.n = n; job.sched = 1;
    const size_t vb = sizeof(double) * (size_t)n * p->g.vecjl;
    int rc; const void* d; void* o;
    if ((rc = st.in(f, vb, &d))) return rc; job.f_in = (const double*)d;
    if ((rc = st.in(y0, sizeof(double) * (size_t)n, &d))) return rc; job.y0 = (const double*)d;
    if ((rc = st.in(y1, vb, &d))) return rc; job.y1 = (const double*)d;
    if ((rc = st.in(y2, vb, &d))) return rc; job.y2 = (const double*)d;
    if ((rc = stage_Omega_nu(p, st, Omega_nu, &job.Omega_nu))) return rc;
    if ((rc = st.outp(out, vb, &o))) return rc; job.f_out = (double*)o;
    if ((rc = launch_job(p, null_tables(), job, nullptr, st))) return rc;
    return st.finish();
}

extern "C" int rg_generate_y(rg_plan* p, int32_t n, const double* f, const double* Omega_nu, double* y0, double* y1,
                             double* y2, double* y3) {
    if (!p || n < 0 || !f || !Omega_nu || !y0 || !y1 || !y2 || !y3) return fail(-1, "bad argument");
    CU(cudaSetDevice(p->device));
    Stage st(p);
    Job job; memset(&job, 0, sizeof(job));
    job.mode = 5; job.n = n; job.sched = 0;
    const size_t vb = sizeof(double) * (size_t)n * p->g.vecjl;
    int rc; const void* d; void* o;
    if ((rc = st.in(f, vb, &d))) return rc; job.f_in = (const double*)d;
    if ((rc = stage_Omega_nu(p, st, Omega_nu, &job.Omega_nu))) return rc;
    if ((rc = st.outp(y0, sizeof(double) * (size_t)n, &o))) return rc; job.y0_out = (double*)o;
    if ((rc = st.outp(y1, vb, &o))) return rc; job.y1_out = (double*)o;
    if ((rc = st.outp(y2, vb, &o))) return rc; job.y2_out = (double*)o;
    if ((rc = st.outp(y3, vb, &o))) return rc; job.y3_out = (double*)o;
    if ((rc = launch_job(p, null_tables(), job, nullptr, st))) return rc;
    return st.finish();
}

extern "C" int rg_norm(rg_plan* p, int32_t n, const double* f, int32_t which, double* out) {
    if (!p || n < 0 || !f || !out || which < 0 || which > 2) return fail(-1, "bad argument");
    CU(cudaSetDevice(p->device));
    Stage st(p);
    int rc; const void* d; void* o;
    if ((rc = st.in(f, sizeof(double) * (size_t)n * p->g.vecjl, &d))) return rc;
    if ((rc = st.outp(out, sizeof(double) * (size_t)n, &o))) return rc;
    if (n > 0) {
        norm_kernel<<<std::min(n, p->sm_count * 16), RG_SMALL_THREADS, 0, p->stream>>>(p->dp, n, (const double*)d, which, (double*)o);
        CU(cudaGetLastError());
    }
    return st.finish();
}

extern "C" int rg_sym(rg_plan* p, int32_t n, const double* f, double* out) {
    if (!p || n < 0 || !f || !out) return fail(-1, "bad argument");
    CU(cudaSetDevice(p->device));
    Stage st(p);
    int rc; const void* d; void* o;
    const size_t vb = sizeof(double) * (size_t)n * p->g.vecjl;
    if ((rc = st.in(f, vb, &d))) return rc;
    if ((rc = st.outp(out, vb, &o))) return rc;
    if (n > 0) {
        sym_kernel<<<std::min(n, p->sm_count * 16), RG_SMALL_THREADS, 0, p->stream>>>(p->dp, n, (const double*)d, (double*)o);
        CU(cudaGetLastError());
    }
    return st.finish();
}

extern "C" int rg_reindex_rescale(rg_plan* p, const rg_tables* tab, int32_t n, const double* f, const int32_t* k0, double* out) {
    if (!p || !tab || n < 0 || !f || !out) return fail(-1, "bad argument");
    CU(cudaSetDevice(p->device));
    Stage st(p);
    int rc; const void* d; const void* dk; void* o;
    const size_t vb = sizeof(double) * (size_t)n * p->g.vecjl;
    if ((rc = st.in(f, vb, &d))) return rc;
    if ((rc = st.in(k0, sizeof(int) * (size_t)n, &dk))) return rc;
    if ((rc = st.outp(out, vb, &o))) return rc;
    if (n > 0) {
        reindex_kernel<<<std::min(n, p->sm_count * 16), RG_SMALL_THREADS, 0, p->stream>>>(p->dp, tab->t, n, (const double*)d, (const int*)dk, (double*)o);
        CU(cudaGetLastError());
    }
    return st.finish();
}

extern "C" int rg_conv_product(rg_plan* p, int32_t n, const double* f1, const double* f2, double* out) {
    if (!p || n < 0 || !f1 || !f2 || !out) return fail(-1, "bad argument");
    CU(cudaSetDevice(p->device));
    cudaDeviceProp prop; CU(cudaGetDeviceProperties(&prop, p->device));
    if ((size_t)2 * p->g.vecjl * 8 > (size_t)prop.sharedMemPerBlockOptin) return fail(-5, "operands do not fit in shared memory");
    Stage st(p);
    int rc; const void* d1; const void* d2; void* o;
    const size_t vb = sizeof(double) * (size_t)n * p->g.vecjl;
    if ((rc = st.in(f1, vb, &d1))) return rc;
    if ((rc = st.in(f2, vb, &d2))) return rc;
    if ((rc = st.outp(out, vb, &o))) return rc;
    if (n > 0) {
        conv_kernel<<<std::min(n, p->sm_count * 2), 256, 2 * p->g.vecjl * 8, p->stream>>>(p->dp, n, (const double*)d1, (const double*)d2, (double*)o);
        CU(cudaGetLastError());
    }
    return st.finish();
}

static int m_s_job(rg_plan* p, int32_t n, const double* y0, const double* y1, const double* y2, const double* Omega_nu,
                   int32_t pnorm_p, double n0, double ell, int32_t* m_star, int32_t* scale, double* norm1);

extern "C" int rg_compute_m_s(rg_plan* p, int32_t n, const double* y0, const double* y1, const double* y2,
                              const double* Omega_nu, double n0, double ell, int32_t* m_star, int32_t* scale, double* norm1) {
    if (!p || n < 0 || !y0 || !y1 || !y2 || !Omega_nu || !m_star || !scale) return fail(-1, "bad argument");
    if (!(n0 > 0.0) || !(ell > 0.0)) return fail(-2, "n0 and ell must be positive");
    return m_s_job(p, n, y0, y1, y2, Omega_nu, 0, n0, ell, m_star, scale, norm1);
}

extern "C" int rg_pnorm(rg_plan* p, int32_t n, const double* y0, const double* y1, const double* y2,
                        const double* Omega_nu, int32_t power, double* out) {
    if (!p || n < 0 || !y0 || !y1 || !y2 || !Omega_nu || !out) return fail(-1, "bad argument");
    if (power < 1 || power > 64) return fail(-2, "pnorm: p must be in 1..64");
    std::vector<int32_t> m((size_t)std::max(n, 1)), s((size_t)std::max(n, 1));
    return m_s_job(p, n, y0, y1, y2, Omega_nu, power, 1.0, 2.0, m.data(), s.data(), out);
}

static int m_s_job(rg_plan* p, int32_t n, const double* y0, const double* y1, const double* y2, const double* Omega_nu,
                   int32_t pnorm_p, double n0, double ell, int32_t* m_star, int32_t* scale, double* norm1) {
    CU(cudaSetDevice(p->device));
    Stage st(p);
    Job job; memset(&job, 0, sizeof(job));
    job.mode = 6; job.n = n; job.sched = 1; job.pnorm_p = pnorm_p; job.ms_n0 = n0; job.ms_ell = ell;
    const size_t vb = sizeof(double) * (size_t)n * p->g.vecjl;
    int rc; const void* d; void* o;
    if ((rc = st.in(y0, sizeof(double) * (size_t)n, &d))) return rc; job.y0 = (const double*)d;
    if ((rc = st.in(y1, vb, &d))) return rc; job.y1 = (const double*)d;
    if ((rc = st.in(y2, vb, &d))) return rc; job.y2 = (const double*)d;
    if ((rc = stage_Omega_nu(p, st, Omega_nu, &job.Omega_nu))) return rc;
    if ((rc = st.outp(m_star, sizeof(int) * (size_t)n, &o))) return rc; job.m_out = (int*)o;
    if ((rc = st.outp(scale, sizeof(int) * (size_t)n, &o))) return rc; job.s_out = (int*)o;
    std::vector<double> dummy((size_t)std::max(n, 1));
    if ((rc = st.outp(norm1 ? norm1 : dummy.data(), sizeof(double) * (size_t)n, &o))) return rc; job.norm1_out = (double*)o;
    if ((rc = launch_job(p, null_tables(), job, nullptr, st))) return rc;
    return st.finish();
}

// ---- host-only schedule introspection (tests replay it in numpy; no device needed)
extern "C" int rg_debug_schedule(int32_t dim, int32_t L, int32_t J, const uint8_t* iminus_wrapped, int32_t mode,
                                 int32_t* n_warps, int32_t* n_boxes, int32_t** lane_out, int32_t** pieces, int32_t* n_entries,
                                 uint32_t** entries, int32_t* n_support, int32_t** supp_rec, int64_t* cost, int32_t* n_stages,
                                 int32_t* max_pieces) {
    if (dim < 1 || dim > 3 || J < 1 || J > 15 || L < 1) return fail(-2, "bad geometry");
    rgplan::Geometry g; g.init(dim, L, J);
    std::vector<uint16_t> mask = masks_from_iminus(g, iminus_wrapped);
    const bool big = needs_global_records(g, mask, RG_SMEM_OPTIN);
    int nw_plan = rg_warps_for(g.M, g.J, big);
    rgplan::Schedule s = rgplan::make_schedule(g, mask, mode == 1, mode == 2, nw_plan, 1, big);
    if (s.n_warps < 0) return fail(-5, "too many output modes for one CTA's convolution schedule");
    *n_warps = s.n_warps; *n_boxes = s.n_boxes; *n_entries = (int)s.entries.size() / 2; *n_support = s.n_support;
    if (cost) { cost[0] = s.fma_useful; cost[1] = s.fma_executed; cost[2] = s.cost_max; cost[3] = s.cost_total; }
    *lane_out = (int32_t*)malloc(sizeof(int32_t) * s.lane_out.size());
    memcpy(*lane_out, s.lane_out.data(), sizeof(int32_t) * s.lane_out.size());
    if (n_stages) *n_stages = s.n_stages;
    if (max_pieces) *max_pieces = s.maxp;
    *pieces = (int32_t*)malloc(sizeof(int32_t) * 4 * s.pieces.size());
    memcpy(*pieces, s.pieces.data(), sizeof(int32_t) * 4 * s.pieces.size());
    *entries = (uint32_t*)malloc(sizeof(uint32_t) * std::max<size_t>(s.entries.size(), 1));
    memcpy(*entries, s.entries.data(), sizeof(uint32_t) * s.entries.size());
    *supp_rec = (int32_t*)malloc(sizeof(int32_t) * s.supp_rec.size());
    memcpy(*supp_rec, s.supp_rec.data(), sizeof(int32_t) * s.supp_rec.size());
    return 0;
}
// run ends of every piece slot of the same schedule ([n_stages * n_warps * max_pieces][2], aligned with `pieces` of
// rg_debug_schedule) and whether the plan was costed for the product-rule form (BIG kernels)
extern "C" int rg_debug_piece_runs(int32_t dim, int32_t L, int32_t J, const uint8_t* iminus_wrapped, int32_t mode,
                                   int32_t** runs, int32_t* n_slots, int32_t* product_rule) {
    if (dim < 1 || dim > 3 || J < 1 || J > 15 || L < 1 || !runs || !n_slots) return fail(-2, "bad geometry");
    rgplan::Geometry g; g.init(dim, L, J);
    std::vector<uint16_t> mask = masks_from_iminus(g, iminus_wrapped);
    const bool big = needs_global_records(g, mask, RG_SMEM_OPTIN);
    rgplan::Schedule s = rgplan::make_schedule(g, mask, mode == 1, mode == 2, rg_warps_for(g.M, g.J, big), 1, big);
    if (s.n_warps < 0) return fail(-5, "too many output modes for one CTA's convolution schedule");
    *n_slots = (int32_t)s.pieces.size();
    if (product_rule) *product_rule = s.prule ? 1 : 0;
    *runs = (int32_t*)malloc(sizeof(int32_t) * std::max<size_t>(s.piece_runs.size(), 1));
    memcpy(*runs, s.piece_runs.data(), sizeof(int32_t) * s.piece_runs.size());
    return 0;
}
extern "C" void rg_debug_free(void* p) { free(p); }
