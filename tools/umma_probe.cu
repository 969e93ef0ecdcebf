// Probe of two tcgen05 layouts that the guides do not spell out (run on the GPU box):
//  (1) TS mode: A operand (bf16) read from TMEM - which column/half holds A[row][k]?
//  (2) M = 64, cta_group::1: which TMEM lanes hold the 64 accumulator rows?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o umma_probe tools/umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include "../diffwave_sashimi_b200/csrc/umma.cuh"
using namespace dwb::umma;

__device__ __forceinline__ void mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}

// out[0]: D of the TS test (128 x 16), out[1]: D lanes dump of the M=64 test (128 lanes x 16 cols)
__global__ void probe(float *out) {
    extern __shared__ uint8_t raw[];
    uint8_t *sm = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t tptr;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // B (N=16 rows, K=64 wide block, only k<16 used): identity B[n][k] = (n == k)
    // A for the M=64 SS test: A[i][k] = i + 1 for k == 0, else 0  -> D[i][0] = i + 1
    uint8_t *Bs = sm, *As = sm + 4096;
    for (int i = tid; i < 4096 / 2; i += blockDim.x) reinterpret_cast<uint16_t *>(Bs)[i] = 0;
    for (int i = tid; i < 16384 / 2; i += blockDim.x) reinterpret_cast<uint16_t *>(As)[i] = 0;
    __syncthreads();
    if (tid < 16) {
        const int r = tid, k = tid;                       // element k of row r: chunk j = k/8, inside chunk k%8
        *reinterpret_cast<__nv_bfloat16 *>(Bs + sw128_off(r, k >> 3) + (k & 7) * 2) = __float2bfloat16(1.0f);
    }
    if (tid < 64) *reinterpret_cast<__nv_bfloat16 *>(As + sw128_off(tid, 0)) = __float2bfloat16((float)(tid + 1));
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    fence_proxy_async_smem();
    if (warp == 0) tmem_alloc(&tptr, 64);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tptr;
    const uint32_t tl = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
    // ---- (1) TS: thread (row r) writes A[r][k] = r + 100 k as packed pairs (even k low half) into columns 32..39
    {
        const int r = tid;
        uint32_t w[8];
        for (int c = 0; c < 8; ++c) {
            const __nv_bfloat16 lo = __float2bfloat16((float)(r % 64 + 100 * ((2 * c) % 2)) + (float)(2 * c)),
                                hi = __float2bfloat16((float)(r % 64) + (float)(2 * c + 1));
            w[c] = (uint32_t)__bfloat16_as_ushort(lo) | ((uint32_t)__bfloat16_as_ushort(hi) << 16);
        }
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n" ::"r"(tl + 32), "r"(w[0]), "r"(w[1]),
                     "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                     : "memory");
        tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
        tc_fence_after();
        mma_bf16_ts(tmem + 0, tmem + 32, smem_desc_sw128(smem_u32(Bs)), idesc_bf16(128, 16), 0);
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    {
        float v[16];
        tmem_ld16(tl + 0, v);
        tmem_wait_ld();
        for (int i = 0; i < 16; ++i) out[tid * 16 + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    // ---- (2) M = 64 SS: clear D columns 0..15 on all lanes, run, dump
    {
        float z[16];
        for (int i = 0; i < 16; ++i) z[i] = -7.0f;
        tmem_st16(tl + 0, z);
        tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
        tc_fence_after();
        mma_bf16_ss(tmem + 0, smem_desc_sw128(smem_u32(As)), smem_desc_sw128(smem_u32(Bs)), idesc_bf16(64, 16), 0);
        mma_commit(&bar);
    }
    mbar_wait(&bar, 1);
    tc_fence_after();
    {
        float v[16];
        tmem_ld16(tl + 0, v);
        tmem_wait_ld();
        for (int i = 0; i < 16; ++i) out[2048 + tid * 16 + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 64);
}

int main() {
    float *d, h[4096];
    cudaMalloc(&d, sizeof(h));
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    probe<<<1, 128, 32768>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("(1) TS mode: D = A (B = I16). expected with 'even k in low half': D[r][k] = r%%64 + k\n");
    for (int r : {0, 1, 5, 31, 32, 70, 127}) {
        printf("  row %3d:", r);
        for (int k = 0; k < 16; ++k) printf(" %5.0f", h[r * 16 + k]);
        printf("\n");
    }
    printf("(2) M=64: lane -> D[.][0] (expected row+1 where a row lives, -7 where untouched)\n");
    for (int l = 0; l < 128; ++l) printf("%s%4.0f", (l % 16 == 0) ? "\n  " : " ", h[2048 + l * 16]);
    printf("\n  col1 of lanes 0..15:");
    for (int l = 0; l < 16; ++l) printf(" %4.0f", h[2048 + l * 16 + 1]);
    printf("\n");
    return 0;
}
