// DiffWaveBlock channel mixing for widths whose tile does not fit the fused single-CTA kernels of mix_umma.cu
// (H = 512, the centre stage of unet d128: x1 alone would take all 512 TMEM columns and one GEMM's A operand all of
// shared memory): the same three contractions as three tcgen05 GEMM launches with fused epilogues.
//                                                                      models/sashimi.py:157-182, s4.py:1435
//   G1  x1 = x + (Wo g + bo)[:H] * sigmoid((Wo g + bo)[H:]) (+cond)        K = H,  N = 2H   -> out (holds x1)
//   --  stats(x1) over channels                                             (channel_stats_kernel)
//   G2  hid = gelu(W1 LN2(x1) + b1)                                         K = H,  N = F    -> hid (B,F,l)
//   G3  x2 = x1 + W2 hid + b2 (+skip)                                       K = F,  N = H    -> out (in place)
//   --  stats(x2): the next block's LayerNorm statistics
// One CTA = 128 time steps (MMA M, the TMEM lanes) x 128 output columns, K streamed in 64-channel chunks: eight
// loader warps read the A chunk from global memory (coalesced along time), apply LN2 where asked, split it into
// bf16 hi/lo and store it K-major / SW128 into a two-slot ring; the weights of the CTA's 128 output rows come
// pre-packed (split, swizzled, consumption order) through a three-stage bulk-copy ring; three MMAs per product
// (hi*hi + lo*hi + hi*lo).  The intermediates (x1, hid) are 16-32 MB at this stage and stay in the 126 MB L2.
// Each CTA streams 128 x K x 4 B of weights instead of all 24 H^2 bytes per 16-step tile as the mma.sync kernel does.
#include "common.cuh"
#include "kernels.h"
#include "umma.cuh"

namespace dwb {
using namespace umma;

constexpr int MG_STAGE = 32768, MG_SLAB = 32768, MG_NSW = 3, MG_NSU = 2, MG_THREADS = 320;
constexpr int MG_OFF_RING = MG_NSU * MG_SLAB;
constexpr int MG_OFF_BAR = MG_OFF_RING + MG_NSW * MG_STAGE;
constexpr int MG_NBAR = 2 * MG_NSW + 2 * MG_NSU + 1;
constexpr int MG_SMEM = MG_OFF_BAR + MG_NBAR * 8 + 16 + 1024;

enum { MG_GLU = 0, MG_GELU = 1, MG_RES = 2 };

struct MixGemmArgs {
    const float *A;                 // (B, K, l) input of the contraction
    const float *stats;             // (B, l, 2) LayerNorm statistics applied to A while loading, or null
    float ln_m, ln_s;
    const uint8_t *Wimg;            // stages (n_tile, kc), 32 KB each
    const float *bias;              // GLU: bo (2H)   GELU: b1 (F)   RES: b2 (H)
    const float *x;                 // GLU: block input x (B,H,l)    RES: x1 (B,H,l) (= out, in place)
    const float *cond;              // GLU: (cond_batch,H,l) or null
    int cond_stride_b;
    const float *skip;              // RES: (B,H,l) or null
    float *out;                     // GLU: x1 (B,H,l)   GELU: hid (B,F,l)   RES: x2 (B,H,l)
    int K, Nout, H, l;              // Nout = channels of `out`
};

__device__ __forceinline__ float mg_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

template <int EPI>
__global__ void __launch_bounds__(MG_THREADS, 1)
mix_gemm_umma_kernel(MixGemmArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *slabs = sm, *ring = sm + MG_OFF_RING;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm + MG_OFF_BAR);
    uint32_t *tptr = reinterpret_cast<uint32_t *>(bars + MG_NBAR);
    uint64_t *wfull = bars, *wempty = wfull + MG_NSW, *ufull = wempty + MG_NSW, *uempty = ufull + MG_NSU, *acc_ready = uempty + MG_NSU;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.z, nt = blockIdx.y, t0 = blockIdx.x * 128, l = a.l, KC = a.K / 64;
    if (tid == 0) {
        for (int i = 0; i < MG_NSW; ++i) {
            mbar_init(wfull + i, 1);
            mbar_init(wempty + i, 1);
        }
        for (int i = 0; i < MG_NSU; ++i) {
            mbar_init(ufull + i, 128);
            mbar_init(uempty + i, 1);
        }
        mbar_init(acc_ready, 1);
        fence_mbar_init();
    }
    if (warp == 9) tmem_alloc(tptr, 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tptr;

    if (warp == 8) {
        // ================= weight producer =====================================================
        if (lane == 0) {
            const uint8_t *img = a.Wimg + (size_t)nt * KC * MG_STAGE;
            for (int i = 0; i < KC; ++i) {
                const int s = i % MG_NSW, n = i / MG_NSW;
                mbar_wait(wempty + s, (n & 1) ^ 1);
                mbar_arrive_expect_tx(wfull + s, MG_STAGE);
                bulk_g2s(ring + (size_t)s * MG_STAGE, img + (size_t)i * MG_STAGE, MG_STAGE, wfull + s);
            }
        }
    } else if (warp == 9) {
        // ================= MMA issuer ==========================================================
        {   // all 32 lanes run the loops; the *_w forms elect the issuing lane
            const uint32_t slab0 = smem_u32(slabs), ring0 = smem_u32(ring);
            constexpr uint32_t idesc = idesc_bf16(128, 128);
#pragma unroll 1
            for (int kc = 0; kc < KC; ++kc) {
                const int us = kc % MG_NSU, s = kc % MG_NSW;
                mbar_wait(ufull + us, (kc / MG_NSU) & 1);
                mbar_wait(wfull + s, (kc / MG_NSW) & 1);
                tc_fence_after();
                const uint32_t abase = slab0 + us * MG_SLAB, bbase = ring0 + s * MG_STAGE;
                if (elect_one()) {      // one election per block of 12 MMAs
#pragma unroll
                    for (int term = 0; term < 3; ++term) {
                        const uint32_t ao = abase + (term == 1 ? MG_SLAB / 2 : 0), bo = bbase + (term == 2 ? MG_STAGE / 2 : 0);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            mma_bf16_ss(tmem, smem_desc_sw128(ao + ks * 32), smem_desc_sw128(bo + ks * 32), idesc,
                                        (kc > 0 || term > 0 || ks > 0) ? 1u : 0u);
                    }
                }
                __syncwarp();
                mma_commit_w(wempty + s);
                mma_commit_w(uempty + us);
            }
            mma_commit_w(acc_ready);
        }
    } else {
        // ================= loaders, then epilogue: one time step per thread =====================
        const int q = warp & 3, cg = warp >> 2;
        const int r = 32 * q + lane, t = t0 + r;
        const bool valid = t < l;
        const uint32_t tl = tmem + ((uint32_t)(32 * q) << 16);
        const size_t tcl = valid ? t : 0;
        float lsc = 1.f, lsh = 0.f;                     // a = lsc * (A + lsh)
        if (a.stats && valid) {
            const float2 ms = *reinterpret_cast<const float2 *>(a.stats + ((size_t)b * l + t) * 2);
            lsc = a.ln_s * ms.y;
            lsh = a.ln_m - ms.x;
        }
        const float *Ab = a.A + (size_t)b * a.K * l + tcl;
        const unsigned rowb = 4u * (unsigned)l;           // channel stride in bytes: one IMAD.WIDE per access instead of a live pointer each
#pragma unroll 1
        for (int kc = cg; kc < KC; kc += MG_NSU) {
            float v[64];
            const char *ap = reinterpret_cast<const char *>(Ab + (size_t)kc * 64 * l);
#pragma unroll
            for (int i = 0; i < 64; ++i) v[i] = valid ? __ldg(reinterpret_cast<const float *>(ap + (unsigned long long)rowb * (unsigned)i)) : 0.f;
            if (a.stats) {
#pragma unroll
                for (int i = 0; i < 64; ++i) v[i] = valid ? lsc * (v[i] + lsh) : 0.f;
            }
            mbar_wait(uempty + cg, ((kc / MG_NSU) & 1) ^ 1);
            uint8_t *slab = slabs + cg * MG_SLAB;
#pragma unroll
            for (int c8 = 0; c8 < 8; ++c8) {
                uint4 hi, lo;
                split8(v + 8 * c8, hi, lo);
                const uint32_t off = sw128_off(r, c8);
                *reinterpret_cast<uint4 *>(slab + off) = hi;
                *reinterpret_cast<uint4 *>(slab + MG_SLAB / 2 + off) = lo;
            }
            fence_proxy_async_smem();
            mbar_arrive(ufull + cg);
        }

        if (EPI == MG_GLU) {
            // columns [0,64): a-rows of channels 64 nt + .., [64,128): their gate rows; this thread: 32 channels
            const int h0 = nt * 64 + cg * 32;
            float xin[32];
            const char *xp = reinterpret_cast<const char *>(a.x + ((size_t)b * a.H + h0) * l + tcl);
#pragma unroll
            for (int i = 0; i < 32; ++i) xin[i] = valid ? __ldg(reinterpret_cast<const float *>(xp + (unsigned long long)rowb * (unsigned)i)) : 0.f;
            mbar_wait(acc_ready, 0);
            tc_fence_after();
            float *op = a.out + ((size_t)b * a.H + h0) * l + tcl;
            const float *cb = a.cond ? a.cond + ((size_t)(a.cond_stride_b ? b : 0) * a.H + h0) * l + tcl : nullptr;
#pragma unroll
            for (int sc = 0; sc < 2; ++sc) {
                float av[16], gv[16];
                tmem_ld16(tl + cg * 32 + sc * 16, av);
                tmem_ld16(tl + 64 + cg * 32 + sc * 16, gv);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int h = h0 + sc * 16 + i;
                    float y = (av[i] + __ldg(a.bias + h)) * mg_sigmoid(gv[i] + __ldg(a.bias + a.H + h));
                    if (cb && valid) y += __ldg(cb + (size_t)(sc * 16 + i) * l);
                    if (valid) op[(size_t)(sc * 16 + i) * l] = xin[sc * 16 + i] + y;
                }
            }
        } else {
            // 128 output channels 128 nt + ..; this thread: 64 of them
            const int n0 = nt * 128 + cg * 64;
            float pre[64];
            if (EPI == MG_RES) {
                const char *xp = reinterpret_cast<const char *>(a.x + ((size_t)b * a.H + n0) * l + tcl);
#pragma unroll
                for (int i = 0; i < 64; ++i) pre[i] = valid ? *reinterpret_cast<const float *>(xp + (unsigned long long)rowb * (unsigned)i) : 0.f;   // x1 (written by G1: plain loads)
                if (a.skip) {
                    const char *sp = reinterpret_cast<const char *>(a.skip + ((size_t)b * a.H + n0) * l + tcl);
#pragma unroll
                    for (int i = 0; i < 64; ++i) pre[i] += valid ? __ldg(reinterpret_cast<const float *>(sp + (unsigned long long)rowb * (unsigned)i)) : 0.f;
                }
            }
            mbar_wait(acc_ready, 0);
            tc_fence_after();
            float *op = a.out + ((size_t)b * a.Nout + n0) * l + tcl;
#pragma unroll
            for (int sc = 0; sc < 4; ++sc) {
                float v[16];
                tmem_ld16(tl + cg * 64 + sc * 16, v);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float z = v[i] + __ldg(a.bias + n0 + sc * 16 + i);
                    v[i] = EPI == MG_GELU ? gelu_fast(z) : pre[sc * 16 + i] + z;
                }
                if (valid) {
                    float *oq = op + (size_t)(sc * 16) * l;
#pragma unroll
                    for (int i = 0; i < 16; ++i, oq += l) *oq = v[i];
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        tc_fence_after();
        tmem_dealloc(tmem, 128);
    }
}

// (mean, rstd) over channels per (b, t): two passes (the tensor is L2 resident), biased variance, no epsilon
// (TransposedLN, models/sashimi.py:17-19)
__global__ void __launch_bounds__(256)
channel_stats_kernel(const float *__restrict__ x, float *__restrict__ stats, int H, int l) {
    const int b = blockIdx.y, t = blockIdx.x * 256 + threadIdx.x;
    if (t >= l) return;
    const float *xp = x + (size_t)b * H * l + t;
    float s = 0.f;
    for (int h = 0; h < H; ++h) s += xp[(size_t)h * l];
    const float mean = s / (float)H;
    float m2 = 0.f;
    for (int h = 0; h < H; ++h) {
        const float d = xp[(size_t)h * l] - mean;
        m2 = fmaf(d, d, m2);
    }
    *reinterpret_cast<float2 *>(stats + ((size_t)b * l + t) * 2) = make_float2(mean, rsqrtf(m2 / (float)H));
}

int mix_gemm_channel_stats(const float *x, float *stats, int H, int l, int B, cudaStream_t st) {
    channel_stats_kernel<<<dim3(ceil_div(l, 256), B), 256, 0, st>>>(x, stats, H, l);
    DWB_LAUNCH_CHECK();
    return DWB_OK;
}

// images: G1 (H/64 tiles x H/64 chunks) | G2 (F/128 x H/64) | G3 (H/128 x F/64), 32 KB per stage
__global__ void mix_gemm_pack_kernel(const float *__restrict__ Wo_t, const float *__restrict__ W1_t, const float *__restrict__ W2_t,
                                     int H, int F, uint8_t *__restrict__ img) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;     // one (stage, row, 16-byte chunk)
    const size_t n1 = (size_t)(H / 64) * (H / 64), n2 = (size_t)(F / 128) * (H / 64), n3 = (size_t)(H / 128) * (F / 64);
    if (idx >= (n1 + n2 + n3) * 128 * 8) return;
    const size_t stage = idx / (128 * 8);
    const int rem = idx % (128 * 8), row = rem / 8, j8 = rem % 8;
    const float *Wt;
    int M, n, kc;
    if (stage < n1) {                   // G1: tile nt = channels 64 nt.., rows [a (64) | gate (64)]
        const int ntile = stage / (H / 64);
        kc = stage % (H / 64);
        Wt = Wo_t; M = 2 * H;
        n = row < 64 ? ntile * 64 + row : H + ntile * 64 + (row - 64);
    } else if (stage < n1 + n2) {
        const size_t s2 = stage - n1;
        kc = s2 % (H / 64);
        Wt = W1_t; M = F; n = (int)(s2 / (H / 64)) * 128 + row;
    } else {
        const size_t s3 = stage - n1 - n2;
        kc = s3 % (F / 64);
        Wt = W2_t; M = H; n = (int)(s3 / (F / 64)) * 128 + row;
    }
    const int k0 = kc * 64 + j8 * 8;
    uint32_t hp[4], lp[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float w0 = Wt[(size_t)(k0 + 2 * e) * M + n], w1 = Wt[(size_t)(k0 + 2 * e + 1) * M + n];
        const __nv_bfloat16 h0 = __float2bfloat16_rn(w0), h1 = __float2bfloat16_rn(w1);
        const __nv_bfloat16 l0 = __float2bfloat16_rn(w0 - __bfloat162float(h0)), l1 = __float2bfloat16_rn(w1 - __bfloat162float(h1));
        hp[e] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
        lp[e] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
    const size_t off = stage * MG_STAGE + (size_t)row * 128 + ((j8 ^ (row & 7)) << 4);
    *reinterpret_cast<uint4 *>(img + off) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
    *reinterpret_cast<uint4 *>(img + off + MG_STAGE / 2) = make_uint4(lp[0], lp[1], lp[2], lp[3]);
}

bool mix_gemm_supported(int H, int F, int l) { return H % 128 == 0 && F == 2 * H && H >= 128 && l >= 1; }

size_t mix_gemm_image_bytes(int H, int F) {
    return ((size_t)(H / 64) * (H / 64) + (size_t)(F / 128) * (H / 64) + (size_t)(H / 128) * (F / 64)) * MG_STAGE;
}

int mix_gemm_pack(int H, int F, const float *Wo_t, const float *W1_t, const float *W2_t, uint8_t *img, cudaStream_t st) {
    const size_t total = mix_gemm_image_bytes(H, F) / MG_STAGE * 128 * 8;
    mix_gemm_pack_kernel<<<(unsigned)ceil_div64((int64_t)total, 256), 256, 0, st>>>(Wo_t, W1_t, W2_t, H, F, img);
    DWB_LAUNCH_CHECK();
    return DWB_OK;
}

template <int EPI>
static int launch_gemm(const MixGemmArgs &g, int ntiles, int B, cudaStream_t st) {
    auto k = mix_gemm_umma_kernel<EPI>;
    DWB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MG_SMEM));
    k<<<dim3(ceil_div(g.l, 128), ntiles, B), MG_THREADS, MG_SMEM, st>>>(g);
    DWB_LAUNCH_CHECK();
    return DWB_OK;
}

// a.Wimg = the three packed images, a.bo / a.b1 / a.b2 = fp32 biases; hid = (B, F, l) workspace.  5 launches.
int mix_gemm_launch(const MixArgs &a, float *hid, int B, cudaStream_t st) {
    DWB_REQUIRE(a.Wimg && hid, DWB_ERR_STATE, "mix_gemm: weights were not packed / no workspace");
    DWB_REQUIRE((int64_t)a.F * a.l < (int64_t)1 << 31 && B <= 65535, DWB_ERR_UNSUPPORTED, "mix_gemm: tensor too large");
    const int H = a.H, F = a.F, l = a.l;
    const size_t n1 = (size_t)(H / 64) * (H / 64), n2 = (size_t)(F / 128) * (H / 64);
    MixGemmArgs g{};
    g.l = l; g.H = H;
    // G1
    g.A = a.g; g.stats = nullptr; g.Wimg = a.Wimg; g.bias = a.bo; g.x = a.x; g.cond = a.cond; g.cond_stride_b = a.cond_stride_b;
    g.out = a.out; g.K = H; g.Nout = H;
    int rc = launch_gemm<MG_GLU>(g, H / 64, B, st);
    if (rc != DWB_OK) return rc;
    channel_stats_kernel<<<dim3(ceil_div(l, 256), B), 256, 0, st>>>(a.out, a.stats_out, H, l);
    DWB_LAUNCH_CHECK();
    // G2
    g.A = a.out; g.stats = a.stats_out; g.ln_m = a.ln2_m; g.ln_s = a.ln2_s; g.Wimg = a.Wimg + n1 * MG_STAGE; g.bias = a.b1;
    g.x = nullptr; g.cond = nullptr; g.out = hid; g.K = H; g.Nout = F;
    rc = launch_gemm<MG_GELU>(g, F / 128, B, st);
    if (rc != DWB_OK) return rc;
    // G3
    g.A = hid; g.stats = nullptr; g.Wimg = a.Wimg + (n1 + n2) * MG_STAGE; g.bias = a.b2; g.x = a.out; g.skip = a.skip;
    g.out = a.out; g.K = F; g.Nout = H;
    rc = launch_gemm<MG_RES>(g, H / 128, B, st);
    if (rc != DWB_OK) return rc;
    channel_stats_kernel<<<dim3(ceil_div(l, 256), B), 256, 0, st>>>(a.out, a.stats_out, H, l);
    DWB_LAUNCH_CHECK();
    return DWB_OK;
}

}  // namespace dwb
