// fp32_peak_probe.cu — measured non-tensor FP32 peak of the device: dependent-free FFMA and packed FFMA2 loops on every SM.
// The roofline of the dominant kernel (FP32-pipe bound, no tensor cores) is reported against this number instead of the
// nominal SMs x 128 x 2 x clock ("measure it").  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a.
// Prints one JSON object.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

template <int ACC>
__global__ void __launch_bounds__(512) ffma_kernel(float* out, float b, float c, int iters) {
    float a[ACC];
#pragma unroll
    for (int k = 0; k < ACC; k++) a[k] = (float)(threadIdx.x + k) * 1e-3f;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int k = 0; k < ACC; k++) a[k] = fmaf(a[k], b, c);
        }
    }
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < ACC; k++) s += a[k];
    if (s == 12345.678f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;      // never true: keeps the loop alive
}

template <int ACC>
__global__ void __launch_bounds__(512) ffma2_kernel(float* out, float b, float c, int iters) {
    float2 a[ACC];
    const float2 bb = make_float2(b, b * 1.0001f), cc = make_float2(c, c * 0.999f);
#pragma unroll
    for (int k = 0; k < ACC; k++) a[k] = make_float2((float)(threadIdx.x + k) * 1e-3f, (float)(threadIdx.x - k) * 1e-3f);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int k = 0; k < ACC; k++) a[k] = __ffma2_rn(a[k], bb, cc);
        }
    }
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < ACC; k++) s += a[k].x + a[k].y;
    if (s == 12345.678f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ACC>
__global__ void __launch_bounds__(512) dfma_kernel(float* out, float b, float c, int iters) {
    double a[ACC];
    const double bb = (double)b, cc = (double)c;
#pragma unroll
    for (int k = 0; k < ACC; k++) a[k] = (double)(threadIdx.x + k) * 1e-3;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int k = 0; k < ACC; k++) a[k] = fma(a[k], bb, cc);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < ACC; k++) s += a[k];
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s;
}

template <class K>
double run(K kern, int sm, int ctas_per_sm, int threads, int iters, double flop_per_thread_iter, float* d_out, float* best_ms) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = sm * ctas_per_sm;
    for (int w = 0; w < 3; w++) kern<<<grid, threads>>>(d_out, 0.999f, 1e-4f, iters);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 10; rep++) {
        cudaEventRecord(e0);
        kern<<<grid, threads>>>(d_out, 0.999f, 1e-4f, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best = std::min(best, ms);
    }
    *best_ms = best;
    return flop_per_thread_iter * (double)iters * (double)grid * threads / (best * 1e-3) / 1e12;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    float* d_out; cudaMalloc(&d_out, 1 << 26);
    const int sm = p.multiProcessorCount;
    const int iters = 1 << 14;
    float ms1, ms2, ms3, ms4;
    // 8 accumulators x 8 unrolled rounds = 64 FFMA per iteration per thread (2 flop each); FFMA2 carries 4 flop
    const double t_ffma = run(ffma_kernel<8>, sm, 4, 512, iters, 64.0 * 2.0, d_out, &ms1);
    const double t_ffma16 = run(ffma_kernel<16>, sm, 2, 512, iters, 128.0 * 2.0, d_out, &ms2);
    const double t_ffma2 = run(ffma2_kernel<8>, sm, 4, 512, iters, 64.0 * 4.0, d_out, &ms3);
    const double t_ffma2_16 = run(ffma2_kernel<16>, sm, 2, 512, iters, 128.0 * 4.0, d_out, &ms4);
    float ms5;
    const double t_dfma = run(dfma_kernel<8>, sm, 4, 512, iters / 16, 64.0 * 2.0, d_out, &ms5);     // FP64 FMA (the loudness and path-finder kernels)
    // sustained: the best variant back to back for ~3 s
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const bool packed = std::max(t_ffma2, t_ffma2_16) > std::max(t_ffma, t_ffma16);
    const int reps = (int)(3000.0f / (packed ? ms3 : ms1)) + 1;
    cudaEventRecord(e0);
    for (int r = 0; r < reps; r++) {
        if (packed) ffma2_kernel<8><<<sm * 4, 512>>>(d_out, 0.999f, 1e-4f, iters);
        else ffma_kernel<8><<<sm * 4, 512>>>(d_out, 0.999f, 1e-4f, iters);
    }
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms_s; cudaEventElapsedTime(&ms_s, e0, e1);
    const double per = (packed ? 64.0 * 4.0 : 64.0 * 2.0) * iters * (double)sm * 4 * 512;
    const double t_sust = per * reps / (ms_s * 1e-3) / 1e12;
    printf("{\"device\": \"%s\", \"sm_count\": %d, \"clock_khz_prop\": %d, \"ffma_tflops\": %.2f, \"ffma_16acc_tflops\": %.2f, "
           "\"ffma2_tflops\": %.2f, \"ffma2_16acc_tflops\": %.2f, \"best_burst_tflops\": %.2f, \"sustained_3s_tflops\": %.2f, "
           "\"sustained_variant\": \"%s\", \"nominal_tflops_at_1965mhz\": %.2f, \"fp64_dfma_tflops\": %.3f, "
           "\"how\": \"dependent-free FFMA / FFMA2 loops, 8 or 16 accumulators per thread, 2048 threads per SM, best of 10 launches (CUDA events); sustained = same kernel back to back for 3 s\"}\n",
           p.name, sm, p.clockRate, t_ffma, t_ffma16, t_ffma2, t_ffma2_16,
           std::max(std::max(t_ffma, t_ffma16), std::max(t_ffma2, t_ffma2_16)), t_sust, packed ? "ffma2" : "ffma",
           sm * 128 * 2 * 1.965e9 / 1e12, t_dfma);
    return 0;
}
