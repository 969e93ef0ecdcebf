// Internal launch interfaces between the plan (api.cu) and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dwb {

struct MixArgs {
    const float *g, *x;            // (B,H,l) S4 activation output, block input
    const float *skip;             // (B,H,l) UNet skip added to the block output, or null
    const float *cond;             // (cond_batch,H,l) conditioning features, or null
    int cond_stride_b;             // 0: broadcast over batch
    const float *Wo_t, *bo;        // [H][2H], (2H)
    const float *W1_t, *b1;        // [H][F],  (F)
    const float *W2_t, *b2;        // [F][H],  (H)
    // split-bf16 mma A fragments of the same weights (mix_mma.cu), or null
    const uint4 *Wo_fh, *Wo_fl, *W1_fh, *W1_fl, *W2_fh, *W2_fl;
    // tcgen05 path (mix_umma.cu): packed shared-memory image of the three weights + biases, or null
    const uint8_t *Wimg;
    const float *bimg;
    long long *trace;              // debug: per-CTA phase timestamps (16 clock64 values per CTA), or null
    float ln2_m, ln2_s;
    float *out, *stats_out;        // (B,H,l), (B,l,2)
    int H, F, l;
    int rev;                       // tcgen05 kernels: walk tiles from the end of the batch (serpentine L2 reuse)
    int jitter;                    // debug (DWB_DEBUG_JITTER, ns): pseudo-random per-thread sleeps at the phase boundaries of the
                                   // tcgen05 mixing kernels - opens the windows of timing-dependent races (TMEM columns are
                                   // reused across phases and no tool tracks them); 0 = off
};

struct PoolArgs {
    const float *x;                // (B,Hi,li)
    const float *skip;             // up only: (B,Ho,li*s) or null
    const float *W_t, *bias;       // down: [Hi*s][Ho]; up: [Hi][Ho*s]
    const uint4 *W_fh, *W_fl;      // split-bf16 mma A fragments, or null
    float *out, *stats_out;
    int Hi, Ho, s, li;
};

struct HeadArgs {
    const float *x;                // (B,C,l)
    const float *stats;            // (B,l,2) or null (no LN)
    float ln_m, ln_s, prescale;
    const float *Wf_t, *bf;        // [C][C], (C)
    const uint4 *Wf_fh, *Wf_fl;    // split-bf16 A fragments of Wf (frag_pack) or null -> fp32 SIMT kernel
    const float *wz;               // (C)
    float bz;
    // optional fused DDPM update: out = (upd_x - c1 eps)/sqrt_alpha (+ sigma noise); the step's coefficients come
    // from device memory so that ONE captured step graph serves every step: ctl = {c1, sqrt_alpha, sigma, noise slot
    // (int bits, -1 = no noise)}, noise of slot i at *noise_base + i * B * l
    const float *upd_x;
    const float *ctl;
    const float *const *noise_base;
    float *out;                    // (B,l)
    int C, l;
};

struct WaveBlockArgs {
    const float *h;                // (B,C,L) layer input
    const float *part_t;           // (B,C) or (C): fc_t(emb) for this layer
    long long part_stride_b;
    const float *cond;             // (cond_batch,2C,L) or null
    int cond_stride_b;
    const float *Wd_t, *bd;        // [3C][2C] tap-major rows (t-d, t, t+d), (2C)
    const float *Wr_t, *br;        // [C][C], (C)
    const float *Ws_t, *bs;        // [C][S], (S)
    const uint4 *Wd_fh, *Wd_fl, *Wr_fh, *Wr_fl, *Ws_fh, *Ws_fl;   // split-bf16 A fragments, or null
    const uint8_t *Wimg;           // tcgen05 path (wave_umma.cu): packed shared-memory image of the three weights, or null
    long long *trace;              // debug: per-CTA phase timestamps (16 clock64 values per CTA), or null
    float *h_out;                  // (B,C,L)
    float *skip;                   // (B,S,L) accumulated in place
    int first;                     // 1: skip is written, not accumulated
    int C, S, L, dilation;
};

int fold_weight(const float *v, const float *g, int M, int Kin, int taps, float *out, cudaStream_t st);
int embed_launch(const float *t, int rows, int E_in, int E_mid, int E_out, const float *W1, const float *b1,
                 const float *W2, const float *b2, const float *Wt, const float *bt, int Mtot, float *emb, float *part,
                 cudaStream_t st);
int init_conv_launch(const float *x, const float *w, const float *bias, int B, int C, int l, float *out, float *stats,
                     cudaStream_t st);
int mix_launch(const MixArgs &a, int B, cudaStream_t st);
bool mix_mma_supported(int H, int F, int l);
int mix_mma_launch(const MixArgs &a, int B, cudaStream_t st);
bool pool_mma_supported(int Hi, int Ho, int s, bool up);
int down_pool_mma_launch(const PoolArgs &a, int B, cudaStream_t st);
int up_pool_mma_launch(const PoolArgs &a, int B, cudaStream_t st);
// tcgen05 pools (pool_umma.cu): all output columns of a 128-step tile in one CTA's TMEM
bool pool_umma_supported(int Hi, int Ho, int s, bool up, int li);
size_t pool_umma_image_bytes(int Hi, int Ho, int s, bool up);
int pool_umma_pack(int Hi, int Ho, int s, bool up, const float *W_t, uint8_t *img, cudaStream_t st);
int pool_umma_launch(const PoolArgs &a, const uint8_t *Wimg, bool up, int B, cudaStream_t st);
int mix_gemm_channel_stats(const float *x, float *stats, int H, int l, int B, cudaStream_t st);
// H = 512 block as three pool_umma-style GEMMs with 512 output columns per CTA (pool_umma.cu)
bool mix_gemm2_supported(int H, int F, int l);
size_t mix_gemm2_image_bytes(int H, int F);
int mix_gemm2_pack(int H, int F, const float *Wo_t, const float *W1_t, const float *W2_t, uint8_t *img, cudaStream_t st);
int mix_gemm2_launch(const MixArgs &a, float *hid, int B, cudaStream_t st);
bool head_umma_supported(int C);
size_t head_umma_image_bytes(int C);
int head_umma_pack(int C, const float *Wf_t, uint8_t *img, cudaStream_t st);
int head_umma_launch(const HeadArgs &h, const uint8_t *Wimg, int B, cudaStream_t st);
bool mix_umma_supported(int H, int F, int l);
size_t mix_umma_image_bytes(int H);
int mix_umma_pack(int H, const float *Wo_t, const float *W1_t, const float *W2_t, const float *bo, const float *b1,
                  const float *b2, uint8_t *img, float *bimg, cudaStream_t st);
int mix_umma_launch(const MixArgs &a, int B, cudaStream_t st);
bool mix_reverse_order();
// widths without a fused tcgen05 kernel (H = 512): three tcgen05 GEMM launches + two statistics launches (mix_gemm_umma.cu)
bool mix_gemm_supported(int H, int F, int l);
size_t mix_gemm_image_bytes(int H, int F);
int mix_gemm_pack(int H, int F, const float *Wo_t, const float *W1_t, const float *W2_t, uint8_t *img, cudaStream_t st);
int mix_gemm_launch(const MixArgs &a, float *hid /* (B,F,l) workspace */, int B, cudaStream_t st);
int frag_pack(const float *Wt, int M, int K, uint32_t *fhi, uint32_t *flo, cudaStream_t st);
int down_pool_launch(const PoolArgs &a, int B, cudaStream_t st);
int up_pool_launch(const PoolArgs &a, int B, cudaStream_t st);
int head_launch(const HeadArgs &a, int B, cudaStream_t st);
int mel_upsample_launch(const float *in, int rows, int F, int Win, int s, const float *w, const float *bias, float *out,
                        cudaStream_t st);
int mel_conv_launch(const float *u, int rows, int K, int Wu, const float *Wt, const float *bias, int Hc, int l, float *out,
                    cudaStream_t st);
int head_mma_launch(const HeadArgs &a, int B, cudaStream_t st);
int wave_block_launch(const WaveBlockArgs &a, int B, cudaStream_t st);
bool wave_mma_supported(int C, int S);
int wave_block_mma_launch(const WaveBlockArgs &a, int B, cudaStream_t st);
bool wave_umma_supported(int C, int S);
size_t wave_umma_image_bytes(int C, int S);
int wave_umma_pack(int C, int S, const float *Wd_t, const float *Wr_t, const float *Ws_t, uint8_t *img, cudaStream_t st);
int wave_block_umma_launch(const WaveBlockArgs &a, int B, cudaStream_t st);

int fftconv_launch(const float *x, const float *stats, const float *part_t, long long psb, float ln_m, float ln_s,
                   const float *kf, float *g, int B, int H, int l, cudaStream_t st, float *scratch = nullptr);
int fft_twiddles(int log2M, cudaStream_t st, const float2 **tw);
bool fftconv3_supported(int log2M, const float *x, const float *stats, const float *g, int l);
int fftconv3_launch(int log2M, const float *x, const float *stats, const float *part_t, long long psb, float ln_m,
                    float ln_s, const float *kf, const float2 *tw, const float2 *tw2 /* compact table: pair twiddles, else null */,
                    float *g, float *scratch, int B, int H, int l, cudaStream_t st);
bool fftconv5_supported(int log2M, const float *x, const float *stats, const float *g, int l);
int fftconv5_launch(int log2M, const float *x, const float *stats, const float *part_t, long long psb, float ln_m, float ln_s,
                    const float *kf, const float2 *tw, float *g, int B, int H, int l, cudaStream_t st);
int fft_pair_twiddles(int log2M, cudaStream_t st, const float2 **tw2);
int s4_generate(const float *C, const float *Bp, const float *P, const float *inv_w_real, const float *w_imag,
                const float *log_dt, const float *omega, int H, int N, int l, double *khat, double *k64, float *k32,
                cudaStream_t st, int64_t *launches);
int fftconv_prepare_f64(const double *k64, const float *D, int H, int l, float *kf, cudaStream_t st);
// table from the fp32 kernels k (2,H,ld) using their first l taps; dir 0: both directions (fftconv_launch's table for
// length l), 1: causal only with D, 2: anticausal only without D (the two v1-layout tables of fftconv_ols_launch)
int fftconv_prepare_f32(const float *k32, int ld, const float *D, int H, int l, int dir, float *kf, cudaStream_t st);
int fftconv_ols_launch(const float *x, const float *stats, const float *part_t, long long psb, float ln_m, float ln_s,
                       const float *kc_c, const float *kc_a, float *g, float *partial, int B, int H, int r, int Lk,
                       cudaStream_t st);

}  // namespace dwb
