// sm_100a primitives used by the tcgen05 kernels (mix_umma.cu, wave_umma.cu):
// mbarriers, 1-D bulk async copies (TMA engine, no tensor map), TMEM allocation, tcgen05.mma
// with shared-memory descriptors, tcgen05.ld/st for the epilogues.  Inline PTX only.
//
// Operand layout everywhere: K-major, 128-byte swizzle ("SW128").  A [rows x 64] bf16 block is
// rows x 128 B; 8-row groups are 1024 B apart (SBO); inside a group the 16-byte chunk j of row r
// lives at chunk position j ^ (r & 7).  One tcgen05.mma consumes K = 16 (32 bytes of every row);
// the k-step inside the 64-wide block is selected by adding 32 B to the descriptor start address.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

#ifndef DWB_MBAR_HINT
#define DWB_MBAR_HINT 20000u      // mbarrier.try_wait suspend-time hint (ns)
#endif

namespace dwb {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ---------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"     // hardware sleep up to the hint (ns)
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(DWB_MBAR_HINT)
        : "memory");
}

// one non-blocking probe of a phase (polling loops over several barriers)
__device__ __forceinline__ bool mbar_try(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

// generic-proxy shared-memory writes -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- bulk async copy global -> shared, completion on an mbarrier (bytes % 16 == 0) ---------
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- 2-D tiled TMA load (tensor map built on the host): box at (c0 = innermost coordinate, c1) -> dense rows in shared memory
__device__ __forceinline__ void tma_load_2d(void *dst_smem, const void *tmap, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void *tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

// One lane of a converged warp (the same lane for the same mask).  The MMA-issuer warps run their loops with all 32 lanes and
// guard only the tcgen05.mma / tcgen05.commit instructions with this: inside an `if (lane == 0)` region the compiler has to
// move every descriptor into uniform registers through a per-instruction broadcast loop (ELECT / R2UR.BROADCAST / BRA.U.ANY,
// ~100 cycles per MMA, measured: the single issuing thread was the tensor pipe's bottleneck); in convergent code the
// descriptors are uniform values from the start.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "elect.sync _|P1, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// ---- TMEM ---------------------------------------------------------------------------------
// whole warp; ncols power of two in [32, 512]; the address is written to *dst (shared memory)
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 16 consecutive fp32 columns of this thread's lane (lane = 32*(warp%4) + laneid, encoded in taddr[31:16])
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n" ::"r"(taddr),
        "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
        : "memory");
}

// 2 columns (the cross-column-group statistics exchange: partners share a TMEM lane)
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, float &v0, float &v1) {
    uint32_t r0, r1;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];\n" : "=r"(r0), "=r"(r1) : "r"(taddr) : "memory");
    v0 = __uint_as_float(r0);
    v1 = __uint_as_float(r1);
}
__device__ __forceinline__ void tmem_st2(uint32_t taddr, float v0, float v1) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};\n" ::"r"(taddr), "r"(__float_as_uint(v0)),
                 "r"(__float_as_uint(v1))
                 : "memory");
}

// ---- descriptors --------------------------------------------------------------------------
// shared-memory matrix descriptor, K-major SW128, dense 8-row groups (SBO = 1024 B)
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);        // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                         // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                         // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                         // layout type: SWIZZLE_128B
    return d;
}
// instruction descriptor: kind::f16, A = B = bf16, D = fp32, both operands K-major, M x N
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]^T ; one thread issues
__device__ __forceinline__ void mma_bf16_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// same with the A operand in TMEM (lane = row, one 32-bit column = two consecutive k, even k in the low half;
// K = 16 per instruction = 8 columns) - layout verified on hardware by tools/umma_probe.cu
__device__ __forceinline__ void mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 8 packed columns (16 bf16 of one row) of a TMEM-resident A operand
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint4 &a, const uint4 &b) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n" ::"r"(taddr), "r"(a.x), "r"(a.y),
                 "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
                 : "memory");
}
// all tcgen05.mma issued so far by this thread -> one arrival on bar when they have completed
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// warp-level forms for the MMA-issuer warps (all 32 lanes converged; see elect_one): one elected lane issues.  elect.sync is
// itself a convergence point of the full warp, and the same lane is elected every time, so tcgen05.commit tracks the MMAs
// issued through these wrappers.
__device__ __forceinline__ void mma_bf16_ss_w(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if (elect_one()) mma_bf16_ss(d_tmem, adesc, bdesc, idesc, accumulate);
}
__device__ __forceinline__ void mma_bf16_ts_w(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if (elect_one()) mma_bf16_ts(d_tmem, a_tmem, bdesc, idesc, accumulate);
}
__device__ __forceinline__ void mma_commit_w(uint64_t *bar) {
    if (elect_one()) mma_commit(bar);
}

// ---- fp32 -> (hi, lo) bf16 split of 8 consecutive K elements, one 16-byte chunk each --------
// hi = rn(v), lo = rn(v - hi): v = hi + lo to ~2^-17 relative
__device__ __forceinline__ void split8(const float *v, uint4 &hi, uint4 &lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float a = v[2 * i], b = v[2 * i + 1];
        uint32_t hp;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hp) : "f"(b), "f"(a));     // low half = a, high half = b
        const float ha = __uint_as_float(hp << 16), hb = __uint_as_float(hp & 0xFFFF0000u);
        uint32_t lp;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lp) : "f"(b - hb), "f"(a - ha));
        h[i] = hp;
        l[i] = lp;
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}
// byte offset of (row r, 16-byte chunk j) inside a [rows x 64] SW128 block
__device__ __forceinline__ uint32_t sw128_off(int r, int j) { return (uint32_t)(r * 128 + ((j ^ (r & 7)) << 4)); }

}  // namespace umma
}  // namespace dwb
