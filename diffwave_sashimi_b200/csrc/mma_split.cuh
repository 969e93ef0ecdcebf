// Split-bf16 tensor-core inner product shared by the channel-mixing kernels (mix_mma.cu,
// wavenet_mma.cu): fp32 operands are split x = hi + lo into two bf16 halves and every product is
// evaluated as hi*hi + lo*hi + hi*lo on mma.sync.m16n8k16 with fp32 accumulation
// (relative error ~2^-17 per product).  Weights arrive pre-split in A-fragment order from global
// memory (one coalesced 16-byte load per lane per fragment); activations sit in shared memory as
// [k][t] bf16 (t contiguous) and reach the B fragments through ldmatrix.trans.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace dwb {

// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], const void *p) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(a));
}

__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint4 &a, uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void split_store(__nv_bfloat16 *hi, __nv_bfloat16 *lo, size_t i, float v) {
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[i] = h;
    lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
}
__device__ __forceinline__ void split_store2(__nv_bfloat16 *hi, __nv_bfloat16 *lo, size_t i, float v0, float v1) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
    *reinterpret_cast<__nv_bfloat162 *>(hi + i) = __halves2bfloat162(h0, h1);
    *reinterpret_cast<__nv_bfloat162 *>(lo + i) = __halves2bfloat162(
        __float2bfloat16_rn(v0 - __bfloat162float(h0)), __float2bfloat16_rn(v1 - __bfloat162float(h1)));
}

// acc[i][n][:] += sum_k A[tile_i][k] * B[k][col0 + 8n ..]     (split-bf16, 3 MMAs per product)
//   fhi/flo: A fragments of the whole weight; tiles[i]: m-tile indices this warp owns
//   Bhi/Blo: smem [K][TTP] bf16; col0: first column of this warp
template <int MT, int NT, int TTP>
__device__ __forceinline__ void gemm_split_bf16(const uint4 *__restrict__ fhi, const uint4 *__restrict__ flo, int KT,
                                                const int (&tiles)[MT], const __nv_bfloat16 *Bhi,
                                                const __nv_bfloat16 *Blo, int col0, float (&acc)[MT][NT][4], int lane) {
    static_assert(NT % 2 == 0, "n-tiles are loaded in pairs");
    // ldmatrix.x4.trans lane addressing: lanes 0-7 k0..7 @n0, 8-15 k8..15 @n0, 16-23 k0..7 @n0+8, 24-31 k8..15 @n0+8
    const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8;
    const int lcol = (lane >> 4) * 8;
    uint4 ah[MT], al[MT];
#pragma unroll
    for (int i = 0; i < MT; ++i) {
        ah[i] = __ldg(fhi + ((size_t)tiles[i] * KT) * 32 + lane);
        al[i] = __ldg(flo + ((size_t)tiles[i] * KT) * 32 + lane);
    }
    for (int kt = 0; kt < KT; ++kt) {
        uint4 nh[MT], nl[MT];
        const int kn = (kt + 1 < KT) ? kt + 1 : kt;      // prefetch the next k-step's weight fragments
#pragma unroll
        for (int i = 0; i < MT; ++i) {
            nh[i] = __ldg(fhi + ((size_t)tiles[i] * KT + kn) * 32 + lane);
            nl[i] = __ldg(flo + ((size_t)tiles[i] * KT + kn) * 32 + lane);
        }
        uint32_t bh[NT / 2][4], bl[NT / 2][4];
        const size_t boff = (size_t)(kt * 16 + lrow) * TTP + col0 + lcol;
#pragma unroll
        for (int n2 = 0; n2 < NT / 2; ++n2) {
            ldsm_x4_trans(bh[n2], Bhi + boff + n2 * 16);
            ldsm_x4_trans(bl[n2], Blo + boff + n2 * 16);
        }
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
            for (int n2 = 0; n2 < NT / 2; ++n2) {
                mma_bf16(acc[i][2 * n2], ah[i], bh[n2][0], bh[n2][1]);
                mma_bf16(acc[i][2 * n2 + 1], ah[i], bh[n2][2], bh[n2][3]);
                mma_bf16(acc[i][2 * n2], al[i], bh[n2][0], bh[n2][1]);
                mma_bf16(acc[i][2 * n2 + 1], al[i], bh[n2][2], bh[n2][3]);
                mma_bf16(acc[i][2 * n2], ah[i], bl[n2][0], bl[n2][1]);
                mma_bf16(acc[i][2 * n2 + 1], ah[i], bl[n2][2], bl[n2][3]);
            }
#pragma unroll
        for (int i = 0; i < MT; ++i) {
            ah[i] = nh[i];
            al[i] = nl[i];
        }
    }
}

template <int MT, int NT>
__device__ __forceinline__ void zero3(float (&acc)[MT][NT][4]) {
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int n = 0; n < NT; ++n)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][n][j] = 0.f;
}

}  // namespace dwb
