// Dependent-chain latencies and single-warp issue rates of the FP64 / shuffle instructions the thinning kernels are
// made of, measured with clock64() on one warp (one block per SM is launched so clocks are under load).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat_probe tools/lat_probe.cu && ./lat_probe
#include <cstdio>
#include <cuda_runtime.h>

#define N 512
template <int OP, int ILP>
__global__ void probe(double* out, long long* cyc, double a0, double b0) {
    double v[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) v[k] = a0 + k + threadIdx.x;
    double b = b0;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N / 16; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
#pragma unroll
            for (int k = 0; k < ILP; ++k) {
                if (OP == 0) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(v[k]) : "d"(b));
                if (OP == 1) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(v[k]) : "d"(b));
                if (OP == 2) asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(v[k]) : "d"(b));
                if (OP == 3) asm volatile("{.reg .pred p; setp.gt.f64 p, %0, %1; selp.f64 %0, %1, %0, p;}" : "+d"(v[k]) : "d"(b));
                if (OP == 4) v[k] = __shfl_xor_sync(0xffffffffu, v[k], 1);
                if (OP == 5) asm volatile("div.rn.f64 %0, %0, %1;" : "+d"(v[k]) : "d"(b));
                if (OP == 6) asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(v[k]));
                if (OP == 7) { float f = (float)v[k]; asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(f)); v[k] = f; }
                if (OP == 8) asm volatile("{.reg .pred p; setp.gt.f64 p, %0, %1; @p add.rn.f64 %0, %0, %1;}" : "+d"(v[k]) : "d"(b));
                if (OP == 9) asm volatile("{.reg .pred p; setp.ne.s32 p, %2, 0; selp.f64 %0, %1, %0, p;}" : "+d"(v[k]) : "d"(b), "r"(i));
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += v[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP, int ILP>
void run(const char* name, int warps) {
    double* out; long long* cyc; long long h;
    cudaMalloc(&out, 148 * 1024 * sizeof(double)); cudaMalloc(&cyc, 8);
    probe<OP, ILP><<<148, 32 * warps>>>(out, cyc, 1.0, 1.0000001);
    probe<OP, ILP><<<148, 32 * warps>>>(out, cyc, 1.0, 1.0000001);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-34s ilp=%d warps/SM=%2d : %7.2f cycles per op per warp (%.2f per dependent step)\n", name, ILP, warps, (double)h / (N * ILP), (double)h / N);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0, 1>("DADD dependent", 1); run<0, 4>("DADD", 1); run<0, 8>("DADD", 1); run<0, 4>("DADD", 4); run<0, 4>("DADD", 8); run<0, 4>("DADD", 16);
    run<1, 1>("DMUL dependent", 1); run<2, 1>("DFMA dependent", 1); run<2, 4>("DFMA", 1); run<2, 8>("DFMA", 1); run<2, 4>("DFMA", 8); run<2, 4>("DFMA", 16);
    run<3, 1>("DSETP+SELP dependent", 1); run<3, 4>("DSETP+SELP", 1); run<3, 4>("DSETP+SELP", 8);
    run<8, 1>("DSETP+@p DADD dependent", 1);
    run<9, 1>("SELP.f64 (2 x SEL) dependent", 1);
    run<4, 1>("SHFL.BFLY f64 (2 x 32) dependent", 1); run<4, 4>("SHFL.BFLY f64", 1); run<4, 4>("SHFL.BFLY f64", 8);
    run<5, 1>("div.rn.f64 dependent", 1); run<5, 4>("div.rn.f64", 1); run<5, 4>("div.rn.f64", 8);
    run<6, 1>("MUFU.RCP64H dependent", 1);
    run<7, 1>("F2F+FFMA+F2F dependent", 1);
    return 0;
}
