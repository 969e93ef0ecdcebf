// Microbenchmark: issue/pipe throughput of packed fp32 (FFMA2/FADD2/FMUL2) vs scalar FFMA/FADD on sm_100a.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ffma2_probe tools/ffma2_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float ffma1(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float fadd1(float a, float b) { float r; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ int iadd(int a, int b) { int r; asm volatile("add.s32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }

template <int MODE>
__global__ void k(float *out, long long *cyc, int iters) {
    float a[8], b = 1.0001f, c = 0.5f;
    u64 A[8];
    int I[8];
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x + i; A[i] = ((u64)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] + 1.f); I[i] = i; }
    u64 Bq = ((u64)__float_as_uint(b) << 32) | __float_as_uint(b), Cq = ((u64)__float_as_uint(c) << 32) | __float_as_uint(c);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) a[i] = ffma1(a[i], b, c);
            if (MODE == 1) A[i] = ffma2(A[i], Bq, Cq);
            if (MODE == 2) a[i] = fadd1(a[i], c);
            if (MODE == 3) A[i] = fadd2(A[i], Cq);
            if (MODE == 4) { a[i] = ffma1(a[i], b, c); I[i] = iadd(I[i], it); }
            if (MODE == 5) { A[i] = ffma2(A[i], Bq, Cq); I[i] = iadd(I[i], it); }
            if (MODE == 6) { A[i] = ffma2(A[i], Bq, Cq); I[i] = iadd(I[i], it); I[i] = iadd(I[i], 3); }
        }
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float((unsigned)A[i]) + I[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char *name, int per_iter_instr) {
    float *out; long long *cyc; const int iters = 4096, threads = 1024, blocks = 148;
    cudaMalloc(&out, blocks * threads * 4); cudaMalloc(&cyc, blocks * 8);
    k<MODE><<<blocks, threads>>>(out, cyc, iters); cudaDeviceSynchronize();
    k<MODE><<<blocks, threads>>>(out, cyc, iters); cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < blocks; ++i) c += h[i]; c /= blocks;
    // warp-instructions per SMSP = iters * 8 * per_iter_instr * (32 warps / 4 SMSPs)
    double wi = (double)iters * 8 * per_iter_instr * 8;
    printf("%-28s cycles %.0f  warp-instr/SMSP %.0f  -> %.3f instr/clk/SMSP\n", name, c, wi, wi / c);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<0>("FFMA", 1); run<1>("FFMA2", 1); run<2>("FADD", 1); run<3>("FADD2", 1);
    run<4>("FFMA+IADD", 2); run<5>("FFMA2+IADD", 2); run<6>("FFMA2+2 IADD", 3);
    return 0;
}
