/*
 * dwb.h — C ABI of libdwb.so, the B200 (sm_100a) DiffWave denoising engine.
 *
 * Drop-in boundary for the reverse-sampling hot path of albertfgu/diffwave-sashimi:
 *   generate.py:23-55 (sampling)  ->  net((x, t), mel)  ->  models/wavenet.py | models/sashimi.py + models/s4.py
 * and for the reference's only native FFI, the pybind11 module `cauchy_mult`
 * (extensions/cauchy/cauchy.cpp:86-95).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch types.  Every entry returns an int
 *     (DWB_OK or an error code); dwb_last_error() gives the message (thread-local).
 *   - All tensor pointers are DEVICE pointers unless the parameter name ends in _host.
 *     The caller owns every buffer it passes; the plan owns folded weights, tables,
 *     cached S4 spectra and its activation workspace.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); entries
 *     only enqueue work on it and never synchronise unless documented.
 *   - One plan per (device, model).  A plan is not re-entrant; independent plans are
 *     thread-safe.  There is no global mutable state besides the thread-local error string.
 *   - There is NO CPU fallback: without a CUDA device every compute entry fails with
 *     DWB_ERR_CUDA.
 */
#ifndef DWB_H_
#define DWB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DWB_VERSION 200

enum dwb_status {
    DWB_OK = 0,
    DWB_ERR_INVALID = 1,     /* bad argument / shape */
    DWB_ERR_CUDA = 2,        /* CUDA runtime error (message has the cudaError string) */
    DWB_ERR_STATE = 3,       /* call order violated (e.g. forward before finalize) */
    DWB_ERR_MISSING = 4,     /* a required state_dict tensor was never supplied */
    DWB_ERR_UNSUPPORTED = 5  /* configuration outside what the kernels implement */
};

enum dwb_model { DWB_MODEL_WAVENET = 0, DWB_MODEL_SASHIMI = 1 };
enum dwb_dtype { DWB_F32 = 0, DWB_I64 = 1 };

#define DWB_MAX_POOL 4

/* Mirrors the constructor kwargs of the reference plugins
 * (configs/model/{wavenet,sashimi}*.yaml; models/wavenet.py:168-176; models/sashimi.py:188-203). */
typedef struct dwb_config {
    int32_t model;              /* enum dwb_model  <- model._name_ */
    int32_t unconditional;      /* 1 = no mel conditioning */
    int32_t embed_in;           /* diffusion_step_embed_dim_in  (128) */
    int32_t embed_mid;          /* diffusion_step_embed_dim_mid (512) */
    int32_t embed_out;          /* diffusion_step_embed_dim_out (512) */
    /* wavenet */
    int32_t res_channels, skip_channels, num_res_layers, dilation_cycle;
    /* sashimi */
    int32_t d_model, n_layers, n_pool, pool[DWB_MAX_POOL], expand, ff, unet;
    int32_t L;                  /* l_max of the top stage (dataset.segment_length) */
    int32_t d_state_half;       /* N: conjugate pairs per SSM (d_state/2 = 32) */
    /* conditioning */
    int32_t mel_bands;          /* 80 */
} dwb_config;

typedef struct dwb_plan dwb_plan;

/* ---- housekeeping ------------------------------------------------------------------ */
int dwb_version(void);
const char *dwb_last_error(void);
/* number of CUDA devices visible to the library (0 => every compute entry fails) */
int dwb_device_count(int *count);

/* ---- plan life cycle --------------------------------------------------------------- */
/* replaces models.construct_model(cfg).cuda()  (models/__init__.py:4-12, generate.py:94) */
int dwb_plan_create(const dwb_config *cfg, int device, dwb_plan **out);
int dwb_plan_destroy(dwb_plan *plan);

/* replaces net.load_state_dict(...) (generate.py:103): hand over ONE state_dict entry under
 * its reference key, e.g. "d_layers.3.layer.kernel.kernel.C".
 * `data` may be host or device memory (`on_device`); it is copied.  Besides state_dict keys the
 * plan accepts "nodes.<l>" = complex64 FFT nodes omega (l/2+1, 2) for S4 kernel generation at
 * stage length l (models/s4.py:553-571); a stage without supplied nodes uses exact roots of
 * unity. */
int dwb_plan_set_tensor(dwb_plan *plan, const char *name, const void *data, int dtype,
                        const int64_t *shape, int ndim, int on_device, void *stream);

/* Folds weight norm, lays weights out for the kernels, builds the t-embedding weight stack and
 * generates + caches every S4 convolution spectrum (Cauchy -> Woodbury -> irfft -> wrapped rfft;
 * models/s4.py:674-807,1391-1403).  Input independent; once per weight load.  Synchronises. */
int dwb_plan_finalize(dwb_plan *plan, void *stream);

/* ---- the hot path ------------------------------------------------------------------- */
/* eps = net((x, t), mel_spec)                     (generate.py:51; wavenet.py:202-210; sashimi.py:277-313)
 *   x    (B,1,L) f32        t  (B) f32 diffusion steps (any real value, per batch element)
 *   cond NULL, or the concatenated per-block conditioning features produced by
 *        dwb_plan_cond_layout()/the host mirror (t-independent, cached per utterance)
 *   cond_batch 1 (broadcast over B) or B
 *   eps  (B,1,L) f32 */
int dwb_forward(dwb_plan *plan, const float *x, const float *t, const float *cond, int cond_batch,
                float *eps, int B, int L, void *stream);

/* x_0 = sampling(net, (B,1,L), diffusion_hyperparams, condition)        (generate.py:23-55)
 *   x_T        (B,1,L)        first normal draw
 *   noise      (T-1,B,1,L)    draw i is used at step t = T-1-i  (reference RNG order)
 *   coef_host  (3,T) host f32: row 0 = (1-alpha_t)/sqrt(1-alpha_bar_t), row 1 = sqrt(alpha_t),
 *                              row 2 = sigma_t      (utils.py:121-151, computed by the caller in
 *                              torch fp32 so the tables are bit-identical to the reference's)
 *   out        (B,1,L)
 *   use_graph  1: ONE diffusion step (every kernel of a network evaluation + the fused DDPM update) is captured once
 *                 into a CUDA graph and replayed T times.  The step's fc_t rows and coefficients are read from a
 *                 plan-owned device record refreshed by a stream-ordered copy before each replay, x lives in a
 *                 plan-owned buffer and the noise pointer behind a device word, so the graph is keyed on
 *                 (B, L, cond, cond_batch) only: new x_T / noise / out buffers, another T or schedule never re-capture.
 *              0: plain stream launches */
int dwb_sample(dwb_plan *plan, const float *x_T, const float *noise, const float *cond, int cond_batch,
               const float *coef_host, int T, float *out, int B, int L, int use_graph, void *stream);

/* Streaming form of dwb_sample: advance x (B,1,L), in place, by the n_steps reverse steps t_start, t_start-1, ...
 * of a T-step schedule.  noise (n_draws,B,1,L): draw i is used at step t_start - i (no draw is read at t = 0).
 * Lets the caller produce the noise of later steps (generate.py:54 draws it on the CPU, one draw per step) while
 * earlier steps run:   x = x_T;  for each chunk: dwb_sample_steps(x, chunk_noise, ..., t_start, n_steps). */
int dwb_sample_steps(dwb_plan *plan, float *x, const float *noise, const float *cond, int cond_batch,
                     const float *coef_host, int T, int t_start, int n_steps, int B, int L, int use_graph, void *stream);

/* conditioning feature layout: number of blocks and, per block, channels H_i and length l_i;
 * features for block i are (cond_batch, H_i, l_i) f32 at float offset cond_batch * offset_i.
 * Pass NULL arrays to query n_blocks only. */
int dwb_plan_cond_layout(dwb_plan *plan, int L, int *n_blocks, int *channels, int *lengths, int64_t *offsets);

/* The t-independent conditioning features of every block from a mel spectrogram, in the layout above
 * (models/sashimi.py:133-141,160-175; models/wavenet.py:62-70,98-111): per block two weight-normed
 * ConvTranspose2d(1,1,(3,2s),stride (1,s),padding (1,s/2)) + leaky-ReLU(0.4), crop to the FIRST l_i
 * samples, 1x1 mel_bands -> H_i.  Once per utterance; the result is what dwb_forward / dwb_sample take
 * as `cond`.  mel (cond_batch, mel_bands, frames) f32 device; out: cond_batch * sum_i H_i l_i floats.
 * Needs the blocks' upsample_conv2d.* / mel_conv.* tensors (DWB_ERR_MISSING otherwise).  Synchronises. */
int dwb_plan_cond_features(dwb_plan *plan, const float *mel, int cond_batch, int frames, int L, float *out,
                           void *stream);

/* ---- introspection / accounting ------------------------------------------------------ */
/* kernels launched by this plan since creation (graph replays count their node launches) */
int dwb_plan_launch_count(dwb_plan *plan, int64_t *count);
/* number of S4 blocks, and a copy of block i's generated time-domain kernel k (2,H,l) f32 */
int dwb_plan_s4_blocks(dwb_plan *plan, int *n_blocks);
int dwb_plan_s4_kernel(dwb_plan *plan, int block, float *k_out, int64_t capacity, int *H, int *l);
/* Channel-mixing half of DiffWaveBlock `block` (models/sashimi.py:157-182 after the S4 convolution)
 * on caller tensors: g, x (B,H,l) -> out (B,H,l), stats_out (B,l,2) = (mean, rstd over channels)
 * of out; skip (B,H,l) or NULL.  exact != 0 forces the fp32 SIMT kernel, 0 takes the plan's path
 * (tcgen05 where the width supports it). */
int dwb_plan_mix_block(dwb_plan *plan, int block, int exact, const float *g, const float *x, const float *skip,
                       float *out, float *stats_out, int B, void *stream);
/* algorithmic HBM bytes and flops of one forward per clip (SURVEY.md §8(d) formulas) */
int dwb_plan_work(dwb_plan *plan, int L, double *bytes_per_clip_step, double *flops_per_clip_step);

/* Device time per kernel category of `iters` eager forwards (CUDA events on `stream` after every
 * launch; synchronises).  ms[c] = total milliseconds, counts[c] = launches, c < DWB_PROF_NCAT.
 * FFTCONVs / MIXs: SaShiMi stage s (0 = top, L samples; 1, 2 = after each pool). */
enum dwb_prof_category {
    DWB_PROF_EMBED = 0, DWB_PROF_INIT = 1, DWB_PROF_HEAD = 2, DWB_PROF_POOL = 3,
    DWB_PROF_FFTCONV0 = 4, DWB_PROF_FFTCONV1 = 5, DWB_PROF_FFTCONV2 = 6, DWB_PROF_FFTCONV3 = 7,
    DWB_PROF_MIX0 = 8, DWB_PROF_MIX1 = 9, DWB_PROF_MIX2 = 10, DWB_PROF_MIX3 = 11,
    DWB_PROF_WAVEBLOCK = 12, DWB_PROF_NCAT = 13
};
int dwb_plan_profile(dwb_plan *plan, const float *x, const float *t, const float *cond, int cond_batch,
                     float *eps, int B, int L, int iters, double *ms, int64_t *counts, void *stream);

/* ---- single ops (same kernels the plan uses; exported for tests and for callers that
 *      only want to replace one reference op) ------------------------------------------ */
/* out[b,l] = sum_n v[b,n]/(z[l]-w[b,n]) + conj(v[b,n])/(z[l]-conj(w[b,n]))
 * complex64 as interleaved float pairs; v,w (batch,N); z (L); out (batch,L).
 * Replaces cauchy_mult_sym_fwd (extensions/cauchy/cauchy.cpp:55-66, cauchy_cuda.cu:242-375);
 * N is the HALF state size as passed by models/s4.py:758; any N >= 1 (the reference requires a
 * power of two in 2..1024). */
int dwb_cauchy_sym_fwd(const float *v, const float *z, const float *w, float *out,
                       int batch, int N, int L, void *stream);

/* The other three entries of the reference module `cauchy_mult` (extensions/cauchy/cauchy.cpp:86-95), same layouts:
 *   dwb_cauchy_fwd      out[b,l] = sum_n v[b,n] / (z[l] - w[b,n])       (cauchy_mult_fwd; N = FULL state size, any N >= 1,
 *                                                                         any L: the reference needs N = 64, L % 32 == 0)
 *   dwb_cauchy_bwd      dv (batch,N), dw (batch,N) from dout (batch,L)    (cauchy_mult_bwd,     cauchy_cuda.cu:139-239)
 *   dwb_cauchy_sym_bwd  the same for the symmetric op, N = half size      (cauchy_mult_sym_bwd, cauchy_cuda.cu:377-487)
 * Gradients follow PyTorch's complex convention (conjugate Wirtinger), i.e. what the reference's autograd.Functions
 * return from backward (extensions/cauchy/cauchy.py:82-86,107-111). */
int dwb_cauchy_fwd(const float *v, const float *z, const float *w, float *out, int batch, int N, int L, void *stream);
int dwb_cauchy_bwd(const float *v, const float *z, const float *w, const float *dout, float *dv, float *dw,
                   int batch, int N, int L, void *stream);
int dwb_cauchy_sym_bwd(const float *v, const float *z, const float *w, const float *dout, float *dv, float *dw,
                       int batch, int N, int L, void *stream);

/* S4 NPLR kernel generation for one layer (models/s4.py:674-807, rank 1, bidirectional):
 * parameters exactly as stored in the state_dict (f32): C (2,H,N,2) B (1,H,N,2) P (1,H,N,2)
 * inv_w_real (H,N) w_imag (H,N) log_dt (H); omega (l/2+1,2) complex64 nodes or NULL for exact
 * roots of unity (+ analytic Nyquist limit).  Evaluated in fp64; k_out (2,H,l) f32. */
int dwb_s4_kernel_gen(const float *C, const float *Bp, const float *P, const float *inv_w_real,
                      const float *w_imag, const float *log_dt, const float *omega,
                      int H, int N, int l, float *k_out, void *stream);

/* Cache the spectrum for the long convolution: kf (H, nfft/4+1, 8) f32, the kernel's internal
 * pointwise table (per conjugate pair of the packed real FFT: untangle, multiply by the spectrum of
 * the wrapped two-sided kernel + D, re-tangle, as one 2x2 complex map) from k (2,H,l) and D (H).
 * nfft = dwb_fftconv_size(l).  Synchronises the stream. */
int dwb_fftconv_size(int l, int *nfft);
int dwb_fftconv_prepare(const float *k, const float *D, int H, int l, float *kf, void *stream);

/* g = GELU( conv(y, k0 | k1) + D*y ),  y = (ln_s*rstd)*(x - mean + ln_m) + part_t[h]
 *   x (B,H,l); stats (B,l,2) = (mean, rstd) over channels; part_t (B,H) or (H) with
 *   part_stride_b = 0; kf from dwb_fftconv_prepare; g (B,H,l).
 * (models/sashimi.py:148-157 + models/s4.py:1391-1430 up to the activation) */
int dwb_fftconv(const float *x, const float *stats, const float *part_t, int64_t part_stride_b,
                float ln_m, float ln_s, const float *kf, float *g, int B, int H, int l, void *stream);

/* ---- training step (WaveNet backbone, unconditional) -----------------------------------------
 * Replaces, for model._name_ = wavenet:   optimizer.zero_grad(); loss = training_loss(net, nn.MSELoss(), audio, dh);
 * loss.backward(); optimizer.step()        (train.py:137-143,198-222; torch.optim.Adam, train.py:92)
 * Parameters, gradients and the two Adam moments are four flat f32 device buffers owned by the CALLER, laid out in
 * net.parameters() order of models/wavenet.py (dwb_trainer_layout; the names are the reference's state_dict keys).  The
 * host mirror makes every nn.Parameter / .grad a view of them, so checkpoints keep the reference's format and the
 * data-parallel gradient exchange (distributed_util.py:97-149) is one all-reduce of one contiguous buffer.
 * SaShiMi and mel-conditioned models return DWB_ERR_UNSUPPORTED. */
typedef struct dwb_trainer dwb_trainer;
/* Pure host function (no device needed): number of parameters and total f32 count; with index >= 0 also entry
 * `index`: its state_dict key (copied into name[name_cap]), float offset and element count.  Any out pointer may be NULL. */
int dwb_trainer_layout(const dwb_config *cfg, int index, char *name, int name_cap, int64_t *offset, int64_t *numel,
                       int *n_params, int64_t *total);
/* activation workspace for batches of exactly (B,1,L); owned by the trainer */
int dwb_trainer_create(const dwb_config *cfg, int device, int B, int L, dwb_trainer **out);
int dwb_trainer_destroy(dwb_trainer *trainer);
int dwb_trainer_info(dwb_trainer *trainer, int64_t *workspace_bytes, int64_t *launches);
/* GEMM family of this trainer's convolutions and their gradients: 1 = split-bf16 tensor cores (three mma.sync per product,
 * fp32 accumulation; the inference kernels' precision), 0 = exact fp32 SIMT tiles.  Default: DWB_TRAIN_GEMM=mma|simt, else the
 * library's built-in choice. */
int dwb_trainer_set_gemm(dwb_trainer *trainer, int tensor_cores);
/* One loss + backward:  x_t = coef[b,0] audio + coef[b,1] z;  eps = net((x_t, steps));  loss = mean((eps - z)^2);
 * grads = d loss / d params (overwritten, not accumulated).
 *   params, grads  flat buffers (dwb_trainer_layout)      audio, z (B,1,L)      steps (B) f32 diffusion steps
 *   coef (B,2) = (sqrt(alpha_bar_t), sqrt(1 - alpha_bar_t)) computed by the caller in torch fp32 (train.py:219)
 *   eps_out (B,1,L) or NULL      loss: one device float.  Enqueues on `stream`; does not synchronise. */
int dwb_trainer_loss_backward(dwb_trainer *trainer, const float *params, float *grads, const float *audio, const float *z,
                              const float *steps, const float *coef, float *eps_out, float *loss, void *stream);
/* torch.optim.Adam.step() (amsgrad off, weight_decay 0) over flat buffers of n floats in ONE launch; `step` counts
 * from 1; gradients are multiplied by grad_scale first (1/world_size after a summing all-reduce). */
int dwb_adam_step(float *params, const float *grads, float *exp_avg, float *exp_avg_sq, int64_t n, float lr, float beta1,
                  float beta2, float eps, int64_t step, float grad_scale, void *stream);

/* ---- debug ---------------------------------------------------------------------------------------------------
 * One tcgen05 WaveNet layer on caller tensors, writing 16 clock64 phase timestamps per CTA to `trace`
 * ((B * ceil(L/128)) x 16 int64; tools/trace_wave.py prints the phase durations). */
/* Fused tcgen05 mixing kernel of DiffWaveBlock `block` with 16 clock64 phase timestamps per CTA (tools/trace_umma.py). */
int dwb_debug_mix_trace(dwb_plan *plan, int block, const float *g, const float *x, float *out, float *stats_out, int B,
                        long long *trace, void *stream);
int dwb_debug_wave_trace(dwb_plan *plan, int layer, const float *h, const float *part, float *h_out, float *skip, int B, int L,
                         long long *trace, void *stream);

/* ---- mel front end (the step before the path for conditional generation) ------------------------------------
 * mel = log(clamp(mel_basis @ |STFT(audio * in_scale)|, clip))     dataloaders/stft.py:100-161,211-244, mel2samp.py:78-84
 *   audio     (B,T) f32 device; in_scale = 1/32768 for int16-valued wav data (MAX_WAV_VALUE), 1 for [-1,1] data
 *   basis_t   (n_fft, 2*(n_fft/2+1)) f32 device: the reference's windowed Fourier basis `forward_basis`
 *             (real rows then imaginary rows of fft(eye(n_fft))[:n_fft/2+1] times the zero-centred window), transposed
 *   mel_basis (n_mels, n_fft/2+1) f32 device   (librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax))
 *   out       (B, n_mels, frames), frames = T/hop + 1 (dwb_mel_frames); reflect padding n_fft/2, stride hop */
int dwb_mel_frames(int T, int n_fft, int hop, int *frames);
int dwb_mel_spectrogram(const float *audio, int B, int T, float in_scale, const float *basis_t, int n_fft, int hop,
                        const float *mel_basis, int n_mels, float clip, float *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DWB_H_ */
