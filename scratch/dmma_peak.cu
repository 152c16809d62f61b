// scratch microbenchmark: FP64 tensor-core MMA (mma.sync.m8n8k4.f64, SASS DMMA) throughput / latency on this GPU
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template<int NACC> __global__ void k(double* out, int iters, double a, double b) {
    double c[NACC][2];
    for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma(c[i][0], c[i][1], a, b);
    }
    double s = 0; for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template<int NACC> void run(int warps_per_sm, int iters) {
    double* out; cudaMalloc(&out, 148 * 2048 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int threads = 32 * warps_per_sm > 1024 ? 1024 : 32 * warps_per_sm, blocks = 148 * (32 * warps_per_sm / threads);
    k<NACC><<<blocks, threads>>>(out, 10, 1e-3, 1e-3);
    cudaEventRecord(e0); k<NACC><<<blocks, threads>>>(out, iters, 1e-3, 1e-3); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double mmas = (double)blocks * (threads / 32) * (double)iters * NACC;
    printf("warps/SM %2d  independent acc %2d: %.3f ms  %.1f G DMMA/s  %.2f TFLOP/s  (%.1f cycles per DMMA per SM at 1.9 GHz; per warp %.1f cycles between issues)\n",
           warps_per_sm, NACC, ms, mmas / ms / 1e6, mmas * 512 / ms / 1e9, 148 * 1.9e6 * ms / mmas, 1.9e6 * ms / ((double)iters * NACC));
    cudaFree(out);
}
int main() {
    run<1>(1, 20000); run<2>(1, 20000); run<4>(1, 20000); run<8>(1, 20000);
    run<1>(4, 20000); run<2>(4, 20000); run<4>(4, 20000); run<8>(4, 20000);
    run<4>(8, 20000); run<8>(8, 20000); run<4>(16, 20000); run<8>(16, 20000); run<8>(32, 20000);
    return 0;
}
