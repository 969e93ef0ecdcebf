// Mel front end of the vocoder path on the GPU (the step before the hot path for conditional generation):
// dataloaders/stft.py:100-161 (STFT.transform as a strided conv1d with the windowed Fourier basis, reflect padding),
// :211-244 (TacotronSTFT.mel_spectrogram: magnitude -> mel filterbank -> log(clamp(., 1e-5))), mel2samp.py:78-84.
//
// One CTA = FR consecutive frames of one clip x all bins.  The FR windows overlap (hop < n_fft), so the CTA stages the
// (FR-1) hop + n_fft samples they span once, with the reflect padding folded into the index.  Thread k owns frequency
// bins k, k + 256, ...: it streams its two basis rows (stored [j][bin] so a warp reads 128 contiguous bytes per j)
// against FR samples broadcast from shared memory - each basis element is used FR times.  Magnitudes stay in shared
// memory for the (dense) mel projection.  fp32 accumulate like the reference's conv1d.
#include "common.cuh"

namespace dwb {

constexpr int MEL_THREADS = 256;
constexpr int MEL_FR = 16;

__global__ void __launch_bounds__(MEL_THREADS)
mel_spectrogram_kernel(const float *__restrict__ audio, int T, float in_scale, const float *__restrict__ basis_t /* [n_fft][2 nb] */,
                       int n_fft, int hop, const float *__restrict__ mel_basis /* [n_mels][nb] */, int n_mels, float clip,
                       float *__restrict__ out /* (B, n_mels, frames) */, int frames) {
    extern __shared__ float sm[];
    const int nb = n_fft / 2 + 1;
    const int span = (MEL_FR - 1) * hop + n_fft;
    float *xs = sm;                    // [span]
    float *mag = sm + span;            // [MEL_FR][nb]
    const int b = blockIdx.y, f0 = blockIdx.x * MEL_FR, tid = threadIdx.x;
    const float *ab = audio + (size_t)b * T;
    const int pad = n_fft / 2;
    for (int i = tid; i < span; i += MEL_THREADS) {
        int t = f0 * hop + i - pad;                 // position in the unpadded clip; reflect (no edge repeat) outside
        if (t < 0) t = -t;
        if (t >= T) t = 2 * (T - 1) - t;
        xs[i] = (t >= 0 && t < T) ? ab[t] * in_scale : 0.f;
    }
    __syncthreads();
    for (int k = tid; k < nb; k += MEL_THREADS) {
        float re[MEL_FR], im[MEL_FR];
#pragma unroll
        for (int f = 0; f < MEL_FR; ++f) re[f] = im[f] = 0.f;
        const float *bp = basis_t + k;
        for (int j = 0; j < n_fft; ++j, bp += 2 * nb) {
            const float br = __ldg(bp), bi = __ldg(bp + nb);
#pragma unroll
            for (int f = 0; f < MEL_FR; ++f) {
                const float x = xs[f * hop + j];
                re[f] = fmaf(br, x, re[f]);
                im[f] = fmaf(bi, x, im[f]);
            }
        }
#pragma unroll
        for (int f = 0; f < MEL_FR; ++f) mag[f * nb + k] = sqrtf(re[f] * re[f] + im[f] * im[f]);
    }
    __syncthreads();
    for (int o = tid; o < n_mels * MEL_FR; o += MEL_THREADS) {
        const int m = o / MEL_FR, f = o - m * MEL_FR;
        if (f0 + f >= frames) continue;
        const float *mb = mel_basis + (size_t)m * nb, *mg = mag + f * nb;
        float acc = 0.f;
        for (int k = 0; k < nb; ++k) acc = fmaf(__ldg(mb + k), mg[k], acc);
        out[((size_t)b * n_mels + m) * frames + f0 + f] = logf(fmaxf(acc, clip));
    }
}

}  // namespace dwb

using namespace dwb;

extern "C" int dwb_mel_frames(int T, int n_fft, int hop, int *frames) {
    DWB_REQUIRE(frames && T >= 1 && n_fft >= 2 && hop >= 1, DWB_ERR_INVALID, "dwb_mel_frames: bad arguments");
    *frames = T / hop + 1;          // reflect padding of n_fft/2 on both sides, stride hop, no further padding (stft.py:143-153)
    return DWB_OK;
}

extern "C" int dwb_mel_spectrogram(const float *audio, int B, int T, float in_scale, const float *basis_t, int n_fft, int hop,
                                   const float *mel_basis, int n_mels, float clip, float *out, void *stream) {
    DWB_REQUIRE(audio && basis_t && mel_basis && out, DWB_ERR_INVALID, "dwb_mel_spectrogram: null pointer");
    DWB_REQUIRE(B >= 1 && B <= 65535 && n_fft >= 2 && n_fft % 2 == 0 && hop >= 1 && n_mels >= 1, DWB_ERR_INVALID,
                "dwb_mel_spectrogram: bad sizes B=%d n_fft=%d hop=%d n_mels=%d", B, n_fft, hop, n_mels);
    DWB_REQUIRE(T > n_fft / 2, DWB_ERR_INVALID, "dwb_mel_spectrogram: %d samples cannot be reflect-padded by %d", T, n_fft / 2);
    const int frames = T / hop + 1, nb = n_fft / 2 + 1;
    const size_t smem = ((size_t)(MEL_FR - 1) * hop + n_fft + (size_t)MEL_FR * nb) * sizeof(float);
    DWB_REQUIRE(smem <= 227 * 1024, DWB_ERR_UNSUPPORTED, "dwb_mel_spectrogram: n_fft=%d hop=%d need %zu B of shared memory", n_fft, hop, smem);
    if (smem > 48 * 1024)
        DWB_CUDA(cudaFuncSetAttribute(mel_spectrogram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mel_spectrogram_kernel<<<dim3(ceil_div(frames, MEL_FR), B), MEL_THREADS, smem, (cudaStream_t)stream>>>(
        audio, T, in_scale, basis_t, n_fft, hop, mel_basis, n_mels, clip, out, frames);
    DWB_LAUNCH_CHECK();
    return DWB_OK;
}
