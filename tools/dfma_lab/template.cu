// dfma_lab/template.cu -- carrier kernel for patched DFMA streams (tools/dfma_lab/lab.py rewrites the register fields and
// control bits of the 256 DFMAs of the loop body in the compiled binary; this source only has to provide a straight-line
// block of 256 DFMAs, 64 initialised accumulator pairs and a timing frame).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
constexpr int NACC = 60, REP = 4;
__global__ void __launch_bounds__(512, 1) carrier(double* out, long long* clk, int iters, double x, double y) {
    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0.25 + 1e-3 * i + 1e-6 * threadIdx.x;
    __shared__ double sm[256];
    sm[threadIdx.x & 255] = 1.0;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < REP; ++r)
#pragma unroll
            for (int i = 0; i < NACC; ++i) acc[i] = fma(acc[i], x, y);
    }
    const long long t1 = clock64();
    double s = sm[(threadIdx.x * 7) & 255];
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if ((threadIdx.x & 31) == 0) atomicMax((unsigned long long*)&clk[blockIdx.x], (unsigned long long)(t1 - t0));   // slowest warp
}
int main(int argc, char** argv) {
    const int warps = argc > 1 ? atoi(argv[1]) : 4, iters = argc > 2 ? atoi(argv[2]) : 2000;
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    double* d_out; long long* d_clk;
    CK(cudaMalloc(&d_out, (size_t)sms * 512 * 8)); CK(cudaMalloc(&d_clk, sms * 8));
    long long best = 1LL << 62;
    for (int r = 0; r < 3; ++r) {
        CK(cudaMemset(d_clk, 0, sms * 8));
        carrier<<<sms, warps * 32>>>(d_out, d_clk, iters, 0.5, 1e-3);
        CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
        std::vector<long long> c(sms); CK(cudaMemcpy(c.data(), d_clk, sms * 8, cudaMemcpyDeviceToHost));
        long long mx = 0; for (auto v : c) mx = v > mx ? v : mx;
        if (mx < best) best = mx;
    }
    // clocks per DFMA per sub-partition: warps/4 warps share a sub-partition
    const double per_smsp = (double)iters * NACC * REP * (warps / 4.0 < 1 ? 1 : warps / 4.0);
    printf("%d warps/SM: %.3f clk per DFMA per sub-partition (%lld clk)\n", warps, best / per_smsp, best);
    return 0;
}
