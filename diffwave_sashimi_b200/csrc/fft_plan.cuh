// Radix plan shared by the per-step FFT convolution (fftconv.cu) and by the one-off spectrum
// builder (s4_kernelgen.cu), which must agree on the digit-reversed order of the spectrum.
//
// The length-n real convolution is done as one M = n/2 point complex FFT per (batch, channel)
// row, in place in shared memory: decimation-in-frequency forward passes (natural -> bit
// reversed), the real-FFT untangle + spectrum product + re-tangle directly in bit-reversed
// order, then decimation-in-time inverse passes (bit reversed -> natural).  No reordering
// pass ever runs; the cached pointwise table is simply stored in the order the forward passes leave.
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>

namespace dwb {

constexpr int FFT_MIN_LOG2M = 4;
constexpr int FFT_MAX_LOG2M = 14;   // M = 16384 complex = 128 KB (+pad) of shared memory
constexpr int FFT_MAX_PASSES = 4;

// log2 of the radix of pass i (pass 0 has span M); zero-terminated.  16s first, remainder last.
__host__ __device__ constexpr int fft_radix_log2(int log2M, int pass) {
    // number of radix-16 passes, then one pass of the remaining bits
    return (pass < log2M / 4) ? 4 : ((pass == log2M / 4) ? (log2M % 4) : 0);
}
__host__ __device__ constexpr int fft_num_passes(int log2M) { return log2M / 4 + ((log2M % 4) ? 1 : 0); }

// smallest supported M = 2^log2M with M >= l (so n = 2M >= 2l: no circular wrap); 0 if too long
__host__ __device__ inline int fft_log2m_for(int l) {
    for (int lg = FFT_MIN_LOG2M; lg <= FFT_MAX_LOG2M; ++lg)
        if ((1 << lg) >= l) return lg;
    return 0;
}

// bit reversal of the low `bits` bits
__host__ __device__ constexpr int fft_brev(int q, int bits) {
    int r = 0;
    for (int i = 0; i < bits; ++i) r |= ((q >> i) & 1) << (bits - 1 - i);
    return r;
}
// Slot (before padding) where the forward passes leave X[k]: every pass stores its output digit
// bit-reversed, so the overall order is the plain bit reversal of k.  The partner of slot p in the
// real-FFT untangle, the slot of X[M - k], is then p ^ ((1 << msb(p)) - 1): no table, no digit math.
__host__ __device__ constexpr int fft_pos(int k, int log2M) { return fft_brev(k, log2M); }
__host__ __device__ constexpr int fft_freq(int pos, int log2M) { return fft_brev(pos, log2M); }
// floats per channel of the cached pointwise table used by fftconv: M/2 + 1 entries of 8
__host__ __device__ constexpr long long fft_table_floats(int log2M) { return 8LL * ((1LL << (log2M - 1)) + 1); }

// Kernel variant per transform size.  The "split" variants (v2 scalar, v3 packed fp32) evaluate the M-point
// transform of the half-empty packed row as two independent M/2-point transforms (even / odd output
// frequencies), one after the other in half the shared memory, so two rows are resident per SM; their
// pointwise table is laid out per half (see kcoef_kernel).  v1 handles every other size.
// DWB_FFT=v1 forces v1 everywhere, DWB_FFT=v2 the scalar split kernel (A/B measurements); read once.
inline int fft_forced_variant() {
    static const int forced = [] {
        const char *e = getenv("DWB_FFT");
        return (e && e[0] == 'v' && (e[1] == '1' || e[1] == '2')) ? (e[1] - '0') : 0;
    }();
    return forced;
}
inline bool fft_use_v2(int log2M) {
    static const bool s12 = getenv("DWB_FFT12") != nullptr;       // experiment: split kernels at n = 8192 too
    return fft_forced_variant() != 1 && (log2M == 14 || (s12 && log2M == 12));
}

// Layout of the cached pointwise table (kcoef kernels in s4_kernelgen.cu): 0 = v1 (one transform), 1 = split,
// 32 B per conjugate pair (the four products alpha..delta), 2 = split + compact, 16 B per pair (sum and
// difference of the two spectrum values; the pair's twiddle comes from a small shared table) — only the
// packed kernel at n = 32768 reads it.  DWB_FFT_WIDE=1 keeps mode 1 there (A/B measurements).
inline int fft_table_mode(int log2M, int l) {
    static const bool wide = getenv("DWB_FFT_WIDE") != nullptr;
    if (!fft_use_v2(log2M)) return 0;
    return (log2M == 14 && (l % 4) == 0 && fft_forced_variant() != 2 && !wide) ? 2 : 1;
}

// shared-memory padding: one float2 of slack per 16 so that the stride-16 and stride-1
// passes (16 consecutive elements per thread) are bank-conflict free
__host__ __device__ constexpr int fft_pad(int i) { return i + (i >> 4); }

}  // namespace dwb
