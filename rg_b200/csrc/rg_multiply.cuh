// rg_multiply.cuh -- RG.expm_multiply (RG.py:152-171) on the device: Al-Mohy--Higham scaling with the
// (m*, s) selection of RG.compute_m_s (RG.py:130-150) driven by a block 1-norm estimator
// (scipy.sparse.linalg.onenormest = Higham--Tisseur Alg. 2.4, t = 2, itmax = 5; RG.py:119-123,
// restated from the algorithm; the reference's `rmatvec = -matvec` "adjoint" is kept, RG.py:121).
// All control flow is CTA-uniform: every decision is taken on block-reduced values.
#pragma once
#include "rg_device.cuh"

namespace rgdev {

#define RG_HAVE_MULTIPLY 1

// dst = L_y src = a0*dA src + conv(dA y, Omega src) - conv(Omega y, dA src)   (RG.py:115-117)
// Y records must already hold y (build_Y).  src, dst: global [j][nu]; dst may alias src.
template <int J, bool BIG>
__device__ __noinline__ void liouville_apply(Ctx& c, const double* src, double* dst, double a0, bool negate) {
    constexpr int JP = J + 1;
    const DevPlan& P = RG_PLAN(c);
    const int M = P.M;
    const Smem sm = smem_view<BIG>(P, RG_SCHED(c), c);
    build_G<J, BIG>(c, src);
    __syncthreads();
    conv_all<J, BIG>(c);
    for (int nu = threadIdx.x; nu < M; nu += blockDim.x) {
        const double* r = reinterpret_cast<const double*>(sm.G + nu * (BIG ? P.g_units : P.rec_units));
#pragma unroll
        for (int j = 0; j < JP; ++j) {
            const double dg = (j < J) ? (BIG ? r[j + 1] * (double)(j + 1) : r[j]) : 0.0;   // dA src  (RG.py:76)
            const double v = a0 * dg + sm.out[nu * JP + j];
            dst[j * M + nu] = negate ? -v : v;
        }
    }
    __syncthreads();
}

__device__ __forceinline__ uint64_t splitmix(uint64_t& s) {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// a column of +-1 (np.random.randint(0,2,n)*2-1 in scipy's resample_column); deterministic per (point, draw)
static __device__ __noinline__ void random_signs(const Ctx& c, double* col, uint64_t key, double scale) {
    const int n = RG_PLAN(c).vecjl;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        uint64_t s = key + (uint64_t)i * 0xD1B54A32D192ED03ull;
        col[i] = ((splitmix(s) >> 40) & 1ull) ? scale : -scale;
    }
}
// (sum|a|, sum|b|) -- plain sums (the estimator's norms are 1-norms whatever NormChoice is)
static __device__ __noinline__ double2 abs_sums(const Ctx& c, const double* a, const double* b) {
    const DevPlan& P = RG_PLAN(c);
    double pa = 0.0, pb = 0.0;
    for (int i = threadIdx.x; i < P.vecjl; i += blockDim.x) { pa += fabs(a[i]); pb += fabs(b[i]); }
    return block_reduce2(smem_red(), 0, pa, pb);
}
// (a.b, a.d)
static __device__ __noinline__ double2 dots(const Ctx& c, const double* a, const double* b, const double* d) {
    const DevPlan& P = RG_PLAN(c);
    double pa = 0.0, pb = 0.0;
    for (int i = threadIdx.x; i < P.vecjl; i += blockDim.x) { const double v = a[i]; pa += v * b[i]; pb += v * d[i]; }
    return block_reduce2(smem_red(), 0, pa, pb);
}
// arg max of h over indices not in excl[0..n_excl): larger value first, ties -> larger index
static __device__ __noinline__ int block_argmax(const Ctx& c, const double* h, const int* excl, int n_excl, double* vmax) {
    const DevPlan& P = RG_PLAN(c);
    double* red = smem_red();
    double bv = -1.0; int bi = -1;
    for (int i = threadIdx.x; i < P.vecjl; i += blockDim.x) {
        bool skip = false;
        for (int q = 0; q < n_excl; ++q) skip |= (excl[q] == i);
        const double v = h[i];
        if (!skip && (v > bv || (v == bv && i > bi))) { bv = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double v2 = __shfl_down_sync(0xffffffffu, bv, o);
        const int i2 = __shfl_down_sync(0xffffffffu, bi, o);
        if (v2 > bv || (v2 == bv && i2 > bi)) { bv = v2; bi = i2; }
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (lane == 0) { red[w] = bv; red[32 + w] = (double)bi; }
    __syncthreads();
    double rv = red[0]; int ri = (int)red[32];
    for (int i = 1; i < nw; ++i) {
        const double v2 = red[i]; const int i2 = (int)red[32 + i];
        if (v2 > rv || (v2 == rv && i2 > ri)) { rv = v2; ri = i2; }
    }
    __syncthreads();
    *vmax = rv;
    return ri;
}

// out = (L_y)^p x  or, for the "adjoint" the reference hands to scipy, (-L_y)^p x
template <int J, bool BIG>
__device__ __noinline__ void apply_pow(Ctx& c, const double* x, double* out, double a0, int p, bool adj) {
    liouville_apply<J, BIG>(c, x, out, a0, adj);
    for (int k = 1; k < p; ++k) liouville_apply<J, BIG>(c, out, out, a0, adj);
}

__device__ __forceinline__ bool isin(int v, const int* set, int n) {
    bool r = false;
    for (int i = 0; i < n; ++i) r |= (set[i] == v);
    return r;
}

// onenormest(A**p) with A = L_y (Y records must hold the unscaled y).  Returns the estimate of ||A^p||_1.
template <int J, bool BIG>
__device__ __noinline__ double onenormest_pow(Ctx& c, double a0, int p) {
    const DevPlan& P = RG_PLAN(c);
    const int n = P.vecjl, t = 2, itmax = 5;
    const double dn = (double)n;
    double* X[2] = {c.ws + (size_t)WS_T0 * n, c.ws + (size_t)WS_T1 * n};
    double* Yc[2] = {c.ws + (size_t)WS_T2 * n, c.ws + (size_t)WS_T3 * n};
    double* S[2] = {c.ws + (size_t)WS_T4 * n, c.ws + (size_t)WS_T5 * n};
    double* So[2] = {c.ws + (size_t)WS_T6 * n, c.ws + (size_t)WS_T7 * n};
    double* H = c.ws + (size_t)WS_T8 * n;
    int ind[16], ind_hist[16];
    int n_hist = 0, ind_best = 0;
    for (int q = 0; q < 16; ++q) { ind[q] = -1; ind_hist[q] = -1; }
    for (int i = threadIdx.x; i < n; i += blockDim.x) { X[0][i] = 1.0 / dn; S[0][i] = 0.0; S[1][i] = 0.0; }
    random_signs(c, X[1], splitmix(c.rng), 1.0 / dn);
    __syncthreads();
    double est_old = 0.0, est = 0.0;
    int k = 1;
    for (;;) {
        apply_pow<J, BIG>(c, X[0], Yc[0], a0, p, false);
        apply_pow<J, BIG>(c, X[1], Yc[1], a0, p, false);
        const double2 mags = abs_sums(c, Yc[0], Yc[1]);
        const int best_j = (mags.y > mags.x) ? 1 : 0;
        est = best_j ? mags.y : mags.x;
        if ((est > est_old || k == 2) && k >= 2) ind_best = ind[best_j];
        if (k >= 2 && est <= est_old) { est = est_old; break; }
        est_old = est;
        { double* t0 = So[0]; So[0] = S[0]; S[0] = t0; t0 = So[1]; So[1] = S[1]; S[1] = t0; }  // S_old = S
        if (k > itmax) break;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            S[0][i] = (Yc[0][i] < 0.0) ? -1.0 : 1.0;   // sign_round_up: sign(0) = 1
            S[1][i] = (Yc[1][i] < 0.0) ? -1.0 : 1.0;
        }
        __syncthreads();
        // (2) every column of S parallel to a column of S_old
        double2 d0 = dots(c, S[0], So[0], So[1]);
        double2 d1 = dots(c, S[1], So[0], So[1]);
        if ((d0.x == dn || d0.y == dn) && (d1.x == dn || d1.y == dn)) break;
        // no column of S parallel to another column of S or to a column of S_old
        for (int tries = 0; tries < 8 && (d0.x == dn || d0.y == dn); ++tries) {
            random_signs(c, S[0], splitmix(c.rng), 1.0);
            __syncthreads();
            d0 = dots(c, S[0], So[0], So[1]);
        }
        for (int tries = 0; tries < 8; ++tries) {
            const double2 dd = dots(c, S[1], S[0], S[0]);
            if (!(dd.x == dn || d1.x == dn || d1.y == dn)) break;
            random_signs(c, S[1], splitmix(c.rng), 1.0);
            __syncthreads();
            d1 = dots(c, S[1], So[0], So[1]);
        }
        // (3) Z = A^H S ; h = max_j |Z_ij|
        apply_pow<J, BIG>(c, S[0], Yc[0], a0, p, true);
        apply_pow<J, BIG>(c, S[1], Yc[1], a0, p, true);
        for (int i = threadIdx.x; i < n; i += blockDim.x) H[i] = fmax(fabs(Yc[0][i]), fabs(Yc[1][i]));
        __syncthreads();
        // (4)
        double hmax;
        int n_ind = t + n_hist;
        if (n_ind > 14) n_ind = 14;
        ind[0] = block_argmax(c, H, ind, 0, &hmax);
        if (k >= 2 && hmax == H[ind_best]) break;
        for (int q = 1; q < n_ind; ++q) { double dummy; ind[q] = block_argmax(c, H, ind, q, &dummy); }
        // (5)
        if (isin(ind[0], ind_hist, n_hist) && isin(ind[1], ind_hist, n_hist)) break;
        {
            int tmp[16], q = 0;
            for (int i = 0; i < n_ind; ++i) if (!isin(ind[i], ind_hist, n_hist)) tmp[q++] = ind[i];
            for (int i = 0; i < n_ind; ++i) if (isin(ind[i], ind_hist, n_hist)) tmp[q++] = ind[i];
            for (int i = 0; i < n_ind; ++i) ind[i] = tmp[i];
        }
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            X[0][i] = (i == ind[0]) ? 1.0 : 0.0;
            X[1][i] = (i == ind[1]) ? 1.0 : 0.0;
        }
        {
            const bool n0 = !isin(ind[0], ind_hist, n_hist), n1 = !isin(ind[1], ind_hist, n_hist);
            if (n0 && n_hist < 16) ind_hist[n_hist++] = ind[0];
            if (n1 && n_hist < 16) ind_hist[n_hist++] = ind[1];
        }
        __syncthreads();
        ++k;
    }
    return est;
}

__device__ __forceinline__ int compute_p_max(int m_max) {   // RG.py:125-128
    const double sq = sqrt((double)m_max);
    const int p_low = (int)floor(sq), p_high = (int)ceil(sq + 1.0);
    int best = p_low;
    for (int p = p_low; p <= p_high; ++p) if (p * (p - 1) <= m_max + 1) best = p;
    return best;
}

// RG.compute_m_s (RG.py:130-150); the fused path calls it with the reference's defaults n0 = 1, ell = 2 (RG.py:155).
// Y records must hold the unscaled y.
template <int J, bool BIG>
__device__ __noinline__ void compute_m_s(Ctx& c, double a0, int& m_out, double& s_out, double& norm1_out,
                                         const double n0 = 1.0, const double ell = 2.0) {
    const DevPlan& P = RG_PLAN(c);
    const int m_max = P.n_theta;
    bool have = false;
    double best_m = 0.0, best_s = 0.0;
    const double norm_ls = onenormest_pow<J, BIG>(c, a0, 1);
    norm1_out = norm_ls;
    const int p_max = compute_p_max(m_max);
    if (norm_ls < 2 * ell * p_max * (p_max + 3) * P.theta[m_max - 1] / (n0 * m_max)) {
        for (int m = 1; m <= m_max; ++m) {
            const double s = ceil(norm_ls / P.theta[m - 1]);
            if (!have || m * s < best_m * best_s) { best_m = m; best_s = s; have = true; }
        }
    } else {
        double norm_d[16];
        const int nd = p_max < 16 ? p_max : 16;
        for (int p = 2; p <= nd + 1; ++p) norm_d[p - 2] = pow(onenormest_pow<J, BIG>(c, a0, p), 1.0 / (double)p);
        for (int p = 2; p <= nd; ++p) {
            const double a = norm_d[p - 2], b = norm_d[(p - 1) % nd];
            const double alpha = a > b ? a : b;
            for (int m = p * (p - 1) - 1; m <= m_max; ++m) {
                if (m < 1) continue;
                const double s = ceil(alpha / P.theta[m - 1]);
                if (!have || m * s < best_m * best_s) { best_m = m; best_s = s; have = true; }
            }
        }
        if (best_s < 1.0) best_s = 1.0;
    }
    m_out = (int)best_m;
    s_out = best_s;
}

// RG.expm_multiply (RG.py:152-171) in place on f.
template <int J, bool BIG>
__device__ __noinline__ int expm_multiply(Ctx& c, double* f, double a_unscaled) {
    int m_star; double scale, n1;
    build_Y<J, BIG>(c, 1.0, 1.0, false);
    __syncthreads();
    // the estimator applies L_y to probe vectors (+-1 columns, sign matrices, unit vectors) that are NOT symmetric
    // under nu -> -nu: the mirror shortcuts of build_G are for the series terms only
    const int symm_saved = c.symm;
    c.symm = 0;
    compute_m_s<J, BIG>(c, a_unscaled, m_star, scale, n1);
    c.symm = symm_saved;
    int err = 0;
    // a non-finite or absurd scale would never terminate in the reference either (range(scale)); cap it
    const long long reps = (scale < 1e9) ? (long long)scale : 1000000000LL;
    for (long long r = 0; r < reps && c.n_liou - c.budget_base <= RG_PLAN(c).max_liouville; ++r) {
        const int e = expm_series<J, BIG>(c, f, a_unscaled, 1.0, scale, true, m_star);
        if (e) err = 1;
        __syncthreads();
    }
    return err;
}

}  // namespace rgdev
