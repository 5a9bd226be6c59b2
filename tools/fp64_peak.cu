// Measures the fp64 roofline denominators on the box (not in MEASURED_PEAKS.json):
//   * DFMA vector peak (independent FMA chains, all SMs)
//   * DMMA m8n8k4 tensor peak (mma.sync f64) -- to decide whether a DMMA formulation could
//     ever beat the vector pipe (BASELINE.json north_star: tensor cores only if dense enough)
//   * shared-memory LDS.64/LDS.128 bandwidth per SM (the co-limiter of the convolution kernel)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu
// Prints one JSON object.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

template <int NCHAIN>
__global__ void __launch_bounds__(512) dfma_kernel(double* out, int iters, double a, double b) {
    double acc[NCHAIN];
#pragma unroll
    for (int i = 0; i < NCHAIN; ++i) acc[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NCHAIN; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NCHAIN; ++i) s += acc[i];
    if (s == 12345.678) out[0] = s;
}

// the convolution's instruction mix: acc[i+j] += a[i]*b[j] with three distinct register operands per DFMA
template <int NA, int NB>
__global__ void __launch_bounds__(512) dfma3_kernel(double* out, const double* in, int iters) {
    double acc[NA + NB], a[NA], b[NB];
#pragma unroll
    for (int i = 0; i < NA + NB; ++i) acc[i] = 0.0;
#pragma unroll
    for (int i = 0; i < NA; ++i) a[i] = in[(threadIdx.x + i) & 1023];
#pragma unroll
    for (int i = 0; i < NB; ++i) b[i] = in[(threadIdx.x * 3 + i) & 1023];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NA; ++i)
#pragma unroll
            for (int j = 0; j < NB; ++j) acc[i + j] = fma(a[i], b[j], acc[i + j]);
#pragma unroll
        for (int i = 0; i < NA; ++i) a[i] += 1e-9;   // keep the loop from being hoisted
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NA + NB; ++i) s += acc[i];
    if (s == 12345.678) out[0] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) dmma_kernel(double* out, int iters, double a0, double b0) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i; }
    double a = a0 + threadIdx.x * 1e-12, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;
}

// each lane reads distinct 8B (or 16B) words, conflict-free; measures LDS wavefront throughput
template <int VEC>
__global__ void __launch_bounds__(512) lds_kernel(double* out, int iters) {
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i;
    __syncthreads();
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    int base = (threadIdx.x * VEC) & 8191;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            int idx = (base + u * 512 * VEC) & 8191;
            if (VEC == 1) { s0 += sm[idx]; }
            else { double2 v = *reinterpret_cast<double2*>(&sm[idx]); s0 += v.x; s1 += v.y; }
        }
        base = (base + 64) & 8191;
    }
    if (s0 + s1 + s2 + s3 == 12345.678) out[0] = s0;
}

template <typename F>
float time_ms(F launch, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); CK(cudaGetLastError());   // a refused launch (e.g. more threads than __launch_bounds__) must not be timed
    launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaGetLastError());
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    double* d; CK(cudaMalloc(&d, 1024));
    const int iters = 20000;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d", p.name, sms, p.clockRate);
    {   // DFMA: 8 chains, 256 thr, 4 CTAs/SM
        const int NCH = 8; int blocks = sms * 4;
        float ms = time_ms([&] { dfma_kernel<NCH><<<blocks, 256>>>(d, iters, 1.0000001, 1e-9); }, 5);
        double flops = 2.0 * NCH * iters * 256.0 * blocks;
        printf(", \"dfma_tflops\": %.3f", flops / ms * 1e-9);
        blocks = sms;   // one CTA of 256 threads per SM (2 warps/SMSP): the convolution kernel's regime
        ms = time_ms([&] { dfma_kernel<NCH><<<blocks, 256>>>(d, iters, 1.0000001, 1e-9); }, 5);
        flops = 2.0 * NCH * iters * 256.0 * blocks;
        printf(", \"dfma_tflops_8warps_per_sm\": %.3f", flops / ms * 1e-9);
        ms = time_ms([&] { dfma_kernel<NCH><<<blocks, 512>>>(d, iters, 1.0000001, 1e-9); }, 5);
        flops = 2.0 * NCH * iters * 512.0 * blocks;
        printf(", \"dfma_tflops_16warps_per_sm\": %.3f", flops / ms * 1e-9);
        ms = time_ms([&] { dfma_kernel<NCH><<<blocks, 128>>>(d, iters, 1.0000001, 1e-9); }, 5);
        flops = 2.0 * NCH * iters * 128.0 * blocks;
        printf(", \"dfma_tflops_4warps_per_sm\": %.3f", flops / ms * 1e-9);
    }
    {   // three-operand DFMA (polynomial product, the convolution's inner loop), 448 and 128 threads per SM
        double* din; CK(cudaMalloc(&din, 1024 * 8)); CK(cudaMemset(din, 0, 1024 * 8));
        const int it3 = 2000;
        int blocks = sms;
        float ms = time_ms([&] { dfma3_kernel<11, 11><<<blocks, 448>>>(d, din, it3); }, 5);
        double flops = 2.0 * (121 + 11) * it3 * 448.0 * blocks;
        printf(", \"dfma_3operand_tflops_14warps_per_sm\": %.3f", flops / ms * 1e-9);
        ms = time_ms([&] { dfma3_kernel<11, 11><<<blocks, 128>>>(d, din, it3); }, 5);
        flops = 2.0 * (121 + 11) * it3 * 128.0 * blocks;
        printf(", \"dfma_3operand_tflops_4warps_per_sm\": %.3f", flops / ms * 1e-9);
    }
    {   // DMMA m8n8k4: 256 MAC per warp instruction
        const int NACC = 8; int blocks = sms * 4;
        float ms = time_ms([&] { dmma_kernel<NACC><<<blocks, 256>>>(d, iters, 1.0000001, 1e-9); }, 5);
        double flops = 2.0 * 256.0 * NACC * iters * (256.0 / 32) * blocks;
        printf(", \"dmma_m8n8k4_tflops\": %.3f", flops / ms * 1e-9);
    }
    {
        int blocks = sms;
        CK(cudaFuncSetAttribute(lds_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        CK(cudaFuncSetAttribute(lds_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        float ms = time_ms([&] { lds_kernel<1><<<blocks, 512, 65536>>>(d, iters); }, 5);
        double bytes = 8.0 * 8 * iters * 512.0;   // per SM
        printf(", \"lds64_bytes_per_clk_per_sm_at_max_clock\": %.2f", bytes / (ms * 1e-3) / (p.clockRate * 1e3));
        printf(", \"lds64_GBps_per_sm\": %.1f", bytes / ms * 1e-6);
        ms = time_ms([&] { lds_kernel<2><<<blocks, 512, 65536>>>(d, iters); }, 5);
        bytes = 16.0 * 8 * iters * 512.0;
        printf(", \"lds128_GBps_per_sm\": %.1f", bytes / ms * 1e-6);
    }
    printf("}\n");
    return 0;
}
