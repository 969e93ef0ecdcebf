// Per-SM ingest rate of 1-D bulk copies (cp.async.bulk global -> shared) from an L2-resident buffer, as a function of the
// copy size and the number of copies in flight:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_probe tools/tma_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t su32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(const uint8_t *src, size_t src_bytes, int chunk, int nslot, int iters, unsigned long long *cycles, int lsu, int same) {
    extern __shared__ __align__(1024) uint8_t sm[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm + 200 * 1024);
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int i = 0; i < nslot; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(su32(bars + i)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long t0 = clock64();
    if (!lsu) {
        if (tid == 0) {
            size_t off = same ? 0 : ((size_t)blockIdx.x * 4096) % src_bytes;
            for (int i = 0; i < iters + nslot; ++i) {
                const int s = i % nslot;
                if (i >= nslot) {      // wait for the copy issued nslot iterations ago
                    const uint32_t par = ((i / nslot) - 1) & 1;
                    asm volatile("{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D;\nbra W;\nD:\n}\n" ::"r"(su32(bars + s)), "r"(par) : "memory");
                }
                if (i < iters) {
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su32(bars + s)), "r"(chunk) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(su32(sm + (size_t)s * chunk)),
                                 "l"(src + off), "r"(chunk), "r"(su32(bars + s)) : "memory");
                    off = (off + chunk) % (src_bytes - chunk);
                    off &= ~(size_t)127;
                }
            }
        }
    } else {
        // plain vector loads by all threads into registers -> shared (the LSU path), same volume
        size_t off = same ? 0 : ((size_t)blockIdx.x * 4096) % src_bytes;
        float4 acc = make_float4(0, 0, 0, 0);
        for (int i = 0; i < iters; ++i) {
            for (int o = tid * 16; o < chunk; o += blockDim.x * 16) {
                const float4 v = *reinterpret_cast<const float4 *>(src + off + o);
                *reinterpret_cast<float4 *>(sm + ((size_t)(i % nslot) * chunk + o)) = v;
                acc.x += v.x;
            }
            off = (off + chunk) % (src_bytes - chunk);
            off &= ~(size_t)127;
        }
        if (acc.x == 1234.5f) cycles[1000] = 1;
    }
    __syncthreads();
    if (tid == 0) cycles[blockIdx.x] = clock64() - t0;
}
int main() {
    const size_t W = 1536 * 1024;      // the H = 256 weight image: L2 resident
    uint8_t *src; cudaMalloc(&src, W); cudaMemset(src, 1, W);
    unsigned long long *cyc; cudaMalloc(&cyc, 2048 * 8);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    for (int same = 0; same < 2; ++same)
    for (int lsu = 0; lsu < 2; ++lsu)
    for (int grid : {1, 64, 148})
        for (int chunk : {16384, 32768})
            for (int nslot : {1, 2, 3, 6}) {
                if ((size_t)chunk * nslot > 200 * 1024) continue;
                const int iters = 256 * 32768 / chunk;
                probe<<<grid, lsu ? 256 : 32, 201 * 1024>>>(src, W, chunk, nslot, iters, cyc, lsu, same);
                probe<<<grid, lsu ? 256 : 32, 201 * 1024>>>(src, W, chunk, nslot, iters, cyc, lsu, same);
                cudaDeviceSynchronize();
                unsigned long long h[148]; cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
                double mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
                printf("%s %s grid %3d chunk %5d slots %d: %.1f B/clk/SM  (%.0f cycles per chunk)\n", same ? "same-addr" : "spread   ", lsu ? "LDG+STS" : "bulk   ", grid, chunk, nslot,
                       (double)chunk * iters / mx, mx / iters);
            }
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
