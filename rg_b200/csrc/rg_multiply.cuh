// rg_multiply.cuh -- RG.expm_multiply (RG.py:152-171) on the device: Al-Mohy--Higham scaling with the
// (m*, s) selection of RG.compute_m_s (RG.py:130-150) driven by a block 1-norm estimator
// (scipy.sparse.linalg.onenormest = Higham--Tisseur Alg. 2.4, t = 2, itmax = 5; RG.py:119-123).
#pragma once
#include "rg_device.cuh"

namespace rgdev {

#define RG_HAVE_MULTIPLY 0

// dst = L_y src = a0*dA src + conv(dA y, Omega src) - conv(Omega y, dA src)   (RG.py:115-117)
// Y records must already hold y (build_Y).  src, dst: global [j][nu]; dst may not alias src.
template <int J>
__device__ __noinline__ void liouville_apply(Ctx& c, const double* src, double* dst, double a0, bool negate) {
    constexpr int JP = J + 1;
    const DevPlan& P = RG_PLAN(c);
    const int M = P.M;
    const Smem sm = smem_view(P, RG_SCHED(c));
    build_G(c, src);
    __syncthreads();
    conv_all<J>(c);
    for (int nu = threadIdx.x; nu < M; nu += blockDim.x) {
        const double* r = reinterpret_cast<const double*>(sm.G + nu * P.rec_units);
#pragma unroll
        for (int j = 0; j < JP; ++j) {
            const double dg = (j < J) ? r[j] : 0.0;
            const double v = a0 * dg + sm.out[nu * JP + j];
            dst[j * M + nu] = negate ? -v : v;
        }
    }
    __syncthreads();
}

template <int J>
__device__ int expm_multiply(Ctx& c, double* f, double a_unscaled) {
    (void)c; (void)f; (void)a_unscaled;
    return -77;
}

}  // namespace rgdev
