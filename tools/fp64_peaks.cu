// FP64 roofline denominators measured on the GPU the bench runs on (MEASURED_PEAKS.json carries no FP64 entry):
//   dmma   mma.sync.aligned.m8n8k4.f64 (SASS DMMA), 512 flop per warp instruction -- the tensor path of the C4 kernel
//          (tcgen05 has no FP64 kind)
//   dfma   fma.rn.f64, 64 flop per warp instruction                               -- the vector FP64 pipe
// Best over a few warps-per-SM / independent-accumulator shapes; prints ONE JSON line (bench.py embeds it).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peaks tools/fp64_peaks.cu
#include <algorithm>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NACC>
__global__ void k_dmma(double* out, int iters, double a, double b) {
    double c[NACC][2];
    for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma(c[i][0], c[i][1], a, b);
    }
    double s = 0;
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void k_dfma(double* out, int iters, double a, double b) {
    double x[NACC];
    for (int i = 0; i < NACC; ++i) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0;
    for (int i = 0; i < NACC; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class K>
double tflops(K kern, int n_sm, int warps_per_sm, int iters, double flop_per_warp_iter, double* out) {
    const int threads = std::min(1024, 32 * warps_per_sm), blocks = n_sm * (32 * warps_per_sm / threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<<<blocks, threads>>>(out, 100, 1.0000001, 1e-3);
    double best = 0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        kern<<<blocks, threads>>>(out, iters, 1.0000001, 1e-3);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        best = std::max(best, (double)blocks * (threads / 32) * (double)iters * flop_per_warp_iter / ms / 1e9);
    }
    return best;
}

int main() {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) { printf("{\"error\": \"no CUDA device\"}\n"); return 1; }
    const int n_sm = prop.multiProcessorCount;
    double* out;
    cudaMalloc(&out, (size_t)n_sm * 2048 * 8);
    double dm = 0, df = 0;
    for (int w : {4, 8, 16, 32}) {
        dm = std::max(dm, tflops(k_dmma<4>, n_sm, w, 20000, 4 * 512.0, out));
        dm = std::max(dm, tflops(k_dmma<8>, n_sm, w, 20000, 8 * 512.0, out));
        df = std::max(df, tflops(k_dfma<8>, n_sm, w, 20000, 8 * 64.0, out));
        df = std::max(df, tflops(k_dfma<16>, n_sm, w, 20000, 16 * 64.0, out));
    }
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"dmma_tflops\": %.2f, \"dfma_tflops\": %.2f, \"sms\": %d, \"max_clock_mhz\": %.0f, \"gpu\": \"%s\", "
           "\"how\": \"tools/fp64_peaks.cu: best of 4..32 warps/SM x 4..16 independent accumulators, 20000 dependent steps, CUDA events\"}\n",
           dm, df, n_sm, clk / 1e3, prop.name);
    return 0;
}
