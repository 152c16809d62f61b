// scratch microbenchmark: FP64 FMA / div / MUFU throughput on this GPU
#include <cstdio>
#include <cuda_runtime.h>
template<int MODE> __global__ void k(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0+4, x5=x0+5, x6=x0+6, x7=x0+7;
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) { x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b); x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b); }
        if (MODE == 1) { x0 = b / (x0 + a); x1 = b / (x1 + a); x2 = b / (x2 + a); x3 = b / (x3 + a); x4 = b/(x4+a); x5=b/(x5+a); x6=b/(x6+a); x7=b/(x7+a);}
        if (MODE == 2) { x0 = log(x0 + a); x1 = log(x1 + a); x2 = log(x2 + a); x3 = log(x3 + a); x4=log(x4+a); x5=log(x5+a); x6=log(x6+a); x7=log(x7+a);}
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3+x4+x5+x6+x7;
}
template<int MODE> void run(const char* name, int iters) {
    double* out; cudaMalloc(&out, 148 * 8 * 256 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(out, 10, 1.0000001, 0.5);
    cudaEventRecord(e0); k<MODE><<<148 * 8, 256>>>(out, iters, 1.0000001, 0.5); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = 148.0 * 8 * 256 * (double)iters * 8;
    printf("%s: %.3f ms, %.2f Gop/s (x2 = %.2f TFLOP/s for FMA)\n", name, ms, ops / ms / 1e6, 2 * ops / ms / 1e9);
}
int main() { run<0>("dfma", 20000); run<1>("ddiv", 2000); run<2>("dlog", 500); return 0; }
