// S4 (NPLR, rank 1, bidirectional) convolution-kernel generation — once per weight load.
//
// Reference: models/s4.py:674-807 (SSKernelNPLR.forward) with the Cauchy sum of
// extensions/cauchy/cauchy_cuda.cu:242-347, then models/s4.py:1391-1403 (two-sided kernel and
// its spectrum).  The reference redoes all of this on every diffusion step; it depends only on
// parameters, so the engine runs it at dwb_plan_finalize() and caches the spectrum.
//
// Design: everything here is input independent and off the per-step path, so it is evaluated
// in fp64 straight from the stored fp32 parameters ("exact arithmetic is the target"):
//   1. s4_khat_kernel   fused Cauchy + Woodbury + bilinear factor: one thread per (h, node),
//                       the six (B|P)x(C0|C1|Q) products share the two reciprocals per state
//                       and never touch memory -> writes 2H x (l/2+1) instead of 6H x (l/2+1).
//   2. irdft_kernel     l-point inverse real DFT by direct summation with a rotating phasor
//                       (l = 16000/4000/1000 is 2^a 5^3; exact table re-sync every 256 terms).
//   3. kf_kernel        spectrum of the wrapped two-sided kernel at the power-of-two size used
//                       by the per-step FFT convolution, with D folded in and the transform
//                       scale absorbed; stored in the digit-reversed order of fftconv.cu.
// Also here: the standalone complex64 Cauchy op that replaces cauchy_mult_sym_fwd.
#include "common.cuh"
#include "fft_plan.cuh"

namespace dwb {

struct cd {  // complex double
    double x, y;
};
__device__ __forceinline__ cd operator+(cd a, cd b) { return {a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ cd operator-(cd a, cd b) { return {a.x - b.x, a.y - b.y}; }
__device__ __forceinline__ cd operator*(cd a, cd b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
__device__ __forceinline__ cd operator*(cd a, double s) { return {a.x * s, a.y * s}; }
__device__ __forceinline__ cd conj(cd a) { return {a.x, -a.y}; }
__device__ __forceinline__ cd inv(cd a) {
    double d = 1.0 / (a.x * a.x + a.y * a.y);
    return {a.x * d, -a.y * d};
}
__device__ __forceinline__ cd operator/(cd a, cd b) { return a * inv(b); }

// ---------------------------------------------------------------------------------------
// 1. K^j[h, m] for j in {0,1}, nodes m = 0..l/2
// ---------------------------------------------------------------------------------------
constexpr int KHAT_THREADS = 128;

__global__ void __launch_bounds__(KHAT_THREADS)
s4_khat_kernel(const float *__restrict__ C, const float *__restrict__ Bp, const float *__restrict__ P,
               const float *__restrict__ inv_w_real, const float *__restrict__ w_imag,
               const float *__restrict__ log_dt, const float *__restrict__ omega /* (nk,2) or null */,
               int H, int N, int l, double *__restrict__ khat /* (2,H,nk,2) */) {
    extern __shared__ double sm[];
    // per state n: lambda (2), then six products v (12)
    double *lam = sm;            // N*2
    double *vv = sm + 2 * N;     // N*12
    const int h = blockIdx.y;
    const int nk = l / 2 + 1;
    const double dt = exp((double)log_dt[h]);
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        const int i = h * N + n;
        double wr = -exp((double)inv_w_real[i]), wi = (double)w_imag[i];
        lam[2 * n] = wr * dt;
        lam[2 * n + 1] = wi * dt;
        cd b = {(double)Bp[2 * i], (double)Bp[2 * i + 1]};
        cd p = {(double)P[2 * i], (double)P[2 * i + 1]};
        cd c0 = {(double)C[2 * i], (double)C[2 * i + 1]};
        cd c1 = {(double)C[2 * (H * N + i)], (double)C[2 * (H * N + i) + 1]};
        cd q = conj(p);
        cd pr[6] = {b * c0, b * c1, b * q, p * c0, p * c1, p * q};
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            vv[12 * n + 2 * k] = pr[k].x;
            vv[12 * n + 2 * k + 1] = pr[k].y;
        }
    }
    __syncthreads();
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= nk) return;
    cd out0, out1;
    const bool exact = (omega == nullptr);
    if (exact && (l % 2 == 0) && m == l / 2) {
        // omega = -1: z -> inf.  Limit of the whole expression is dt * Re sum_n B_n C^j_n.
        double s0 = 0, s1 = 0;
        for (int n = 0; n < N; ++n) {
            s0 += vv[12 * n + 0];
            s1 += vv[12 * n + 2];
        }
        out0 = {dt * s0, 0.0};
        out1 = {dt * s1, 0.0};
    } else {
        cd om;
        if (exact) {
            double s, c;
            sincospi(-2.0 * (double)m / (double)l, &s, &c);
            om = {c, s};
        } else {
            om = {(double)omega[2 * m], (double)omega[2 * m + 1]};
        }
        const cd one = {1.0, 0.0};
        const cd opw = one + om;                 // 1 + omega
        const cd z = ((one - om) * 2.0) / opw;   // bilinear node
        cd acc[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) acc[k] = {0.0, 0.0};
        for (int n = 0; n < N; ++n) {
            cd la = {lam[2 * n], lam[2 * n + 1]};
            cd r1 = inv(z - la);
            cd r2 = inv(z - conj(la));
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                cd v = {vv[12 * n + 2 * k], vv[12 * n + 2 * k + 1]};
                acc[k] = acc[k] + v * r1 + conj(v) * r2;
            }
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) acc[k] = acc[k] * dt;
        // Woodbury (rank 1): k = r00 - r01 r10 / (1 + r11); rows {B,P} x cols {C0,C1,Q}
        const cd den = inv(one + acc[5]);
        const cd f = (one * 2.0) / opw;
        out0 = (acc[0] - acc[2] * acc[3] * den) * f;
        out1 = (acc[1] - acc[2] * acc[4] * den) * f;
    }
    const size_t o0 = ((size_t)(0 * H + h) * nk + m) * 2, o1 = ((size_t)(1 * H + h) * nk + m) * 2;
    khat[o0] = out0.x;
    khat[o0 + 1] = out0.y;
    khat[o1] = out1.x;
    khat[o1 + 1] = out1.y;
}

// ---------------------------------------------------------------------------------------
// 2. k[r, t] = irfft(khat[r, :], n = l)[t]      (rows r = 2H)
// ---------------------------------------------------------------------------------------
constexpr int DFT_THREADS = 128;
constexpr int DFT_CHUNK = 256;   // phasor re-synchronised from sincospi every chunk

__global__ void __launch_bounds__(DFT_THREADS)
irdft_kernel(const double *__restrict__ khat /* (R,nk,2) */, int l, double *__restrict__ k64 /* (R,l) */,
             float *__restrict__ k32 /* (R,l) or null */) {
    __shared__ double sk[2 * DFT_CHUNK];
    const int r = blockIdx.y;
    const int nk = l / 2 + 1;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const double *row = khat + (size_t)r * nk * 2;
    const int last = (l % 2 == 0) ? l / 2 : nk;   // terms 1..last-1 are doubled
    double acc = 0.0;
    double ws, wc;
    sincospi(2.0 * (double)(t % l) / (double)l, &ws, &wc);   // e^{+2 pi i t / l}
    for (int m0 = 1; m0 < last; m0 += DFT_CHUNK) {
        const int cnt = min(DFT_CHUNK, last - m0);
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * cnt; i += blockDim.x) sk[i] = row[2 * m0 + i];
        __syncthreads();
        if (t < l) {
            double ps, pc;
            const long long ph = ((long long)m0 * (long long)t) % l;
            sincospi(2.0 * (double)ph / (double)l, &ps, &pc);
            for (int i = 0; i < cnt; ++i) {
                acc += sk[2 * i] * pc - sk[2 * i + 1] * ps;
                const double nc = pc * wc - ps * ws;
                ps = pc * ws + ps * wc;
                pc = nc;
            }
        }
    }
    if (t >= l) return;
    double v = row[0] + 2.0 * acc;
    if (l % 2 == 0) v += (t & 1) ? -row[2 * (l / 2)] : row[2 * (l / 2)];
    v /= (double)l;
    k64[(size_t)r * l + t] = v;
    if (k32) k32[(size_t)r * l + t] = (float)v;
}

// ---------------------------------------------------------------------------------------
// 3. spectrum of the wrapped two-sided kernel, in fftconv's layout
//    kk[s] = k0[s] (s < l), kk[n - s] = k1[s-1] (s = 1..l);  K[f] = sum_j kk[j] e^{-2 pi i f j / n}
//    Kd[f] = (K[f] + D[h]) / (4 M), M = n/2, f = 0..M in natural order (fp64), then kcoef_kernel turns
//    each conjugate pair (k, M-k) into the 2x2 complex map fftconv applies (untangle, product, re-tangle).
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(DFT_THREADS)
kf_kernel(const T *__restrict__ k /* (2,H,l) */, const float *__restrict__ D /* (H) or null */, int H, int l,
          int log2M, double2 *__restrict__ Kd /* (H, M+1) */) {
    __shared__ double s0[DFT_CHUNK], s1[DFT_CHUNK];
    const int h = blockIdx.y;
    const int M = 1 << log2M, n = 2 * M;
    const int f = blockIdx.x * blockDim.x + threadIdx.x;   // 0..M
    const T *k0 = k + (size_t)h * l;
    const T *k1 = k + (size_t)(H + h) * l;
    double ar = 0.0, ai = 0.0;
    double ws, wc;
    sincospi(-2.0 * (double)(f % n) / (double)n, &ws, &wc);   // W = e^{-2 pi i f / n}
    for (int j0 = 0; j0 < l; j0 += DFT_CHUNK) {
        const int cnt = min(DFT_CHUNK, l - j0);
        __syncthreads();
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
            s0[i] = (double)k0[j0 + i];
            s1[i] = (double)k1[j0 + i];
        }
        __syncthreads();
        if (f <= M) {
            // causal part: k0[s] W^{f s}, s = j0 + i ; anti-causal: k1[s-1] W^{-f s}, s = j0 + i + 1
            double ps, pc, qs, qc;
            const long long ph = ((long long)f * (long long)j0) % n;
            sincospi(-2.0 * (double)ph / (double)n, &ps, &pc);           // W^{f j0}
            const long long qh = ((long long)f * (long long)(j0 + 1)) % n;
            sincospi(2.0 * (double)qh / (double)n, &qs, &qc);            // W^{-f (j0+1)}
            for (int i = 0; i < cnt; ++i) {
                ar += s0[i] * pc + s1[i] * qc;
                ai += s0[i] * ps + s1[i] * qs;
                double nc = pc * wc - ps * ws;
                ps = pc * ws + ps * wc;
                pc = nc;
                nc = qc * wc + qs * ws;       // multiply by conj(W)
                qs = qs * wc - qc * ws;
                qc = nc;
            }
        }
    }
    if (f > M) return;
    const double scale = 1.0 / (4.0 * (double)M);
    const double d = D ? (double)D[h] : 0.0;
    Kd[(size_t)h * (M + 1) + f] = make_double2((ar + d) * scale, ai * scale);
}

// Pointwise table of fftconv.cu: entry e of channel h (8 floats: alpha, beta, gamma, delta) for the
// leader slot p = 2e (e = M/2: slot 1, the self-paired k = M/2; e = 0: (K'[0], K'[M]) for DC/Nyquist).
// With a = Z[k], b = Z[M-k] of the packed transform, w = e^{-2 pi i k / n}, u = 1 - i w, v = 1 + i w:
//   Y[k] = u a + v conj(b) (real-FFT untangle, factor 1/2 in the scale), P = K'[k] Y[k], ... re-tangle gives
//   Z'[k]   = alpha a + beta conj(b),   Z'[M-k] = conj(gamma a + delta conj(b))
//   alpha = K1 |u|^2 + conj(K2) |v|^2      beta  = K1 v conj(u) + conj(K2) u conj(v)
//   gamma = K1 u conj(v) + conj(K2) v conj(u)   delta = K1 |v|^2 + conj(K2) |u|^2,   K1 = K'[k], K2 = K'[M-k]
// v2 (split transform, fft_use_v2): the same maps regrouped per half transform of size Mh = M/2, slots p' =
// bit reversal of k' over log2M - 1 bits: entries [0, Mh/2] pair the even frequencies k = 2k' (e = 0: DC/Nyquist,
// e = Mh/2: slot 1 = the self-paired k = M/2, else leader slot 2e); entries Mh/2 + 1 + e pair the odd
// frequencies k = 2k' + 1 (leader slot 2e, partner slot = complement).  Same size, M/2 + 1 entries.
__global__ void kcoef_kernel(const double2 *__restrict__ Kd, int log2M, int split, float *__restrict__ kc) {
    const int M = 1 << log2M, n = 2 * M;
    const int e = blockIdx.x * blockDim.x + threadIdx.x, h = blockIdx.y;
    if (e > M / 2) return;
    const double2 *K = Kd + (size_t)h * (M + 1);
    float *o = kc + ((size_t)h * (M / 2 + 1) + e) * 8;
    if (e == 0) {
        o[0] = (float)K[0].x;
        o[1] = (float)K[M].x;
        for (int i = 2; i < 8; ++i) o[i] = 0.f;
        return;
    }
    int k;
    if (!split) {
        const int p = (e == M / 2) ? 1 : 2 * e;
        k = fft_freq(p, log2M);
    } else {
        const int Mh = M / 2;
        if (e <= Mh / 2) {
            const int p = (e == Mh / 2) ? 1 : 2 * e;
            k = 2 * fft_freq(p, log2M - 1);
        } else {
            k = 2 * fft_freq(2 * (e - Mh / 2 - 1), log2M - 1) + 1;
        }
    }
    const double2 K1 = K[k], K2 = make_double2(K[M - k].x, -K[M - k].y);      // K2 = conj(K'[M-k])
    double ws, wc;
    sincospi(-2.0 * (double)k / (double)n, &ws, &wc);
    const double2 u = make_double2(1.0 + ws, -wc), v = make_double2(1.0 - ws, wc);   // 1 -+ i w
    auto mul = [](double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); };
    auto conj = [](double2 a) { return make_double2(a.x, -a.y); };
    const double u2 = u.x * u.x + u.y * u.y, v2 = v.x * v.x + v.y * v.y;
    const double2 vu = mul(v, conj(u)), uv = mul(u, conj(v));
    const double2 t1 = mul(K1, vu), t2 = mul(K2, uv), t3 = mul(K1, uv), t4 = mul(K2, vu);
    o[0] = (float)(K1.x * u2 + K2.x * v2);
    o[1] = (float)(K1.y * u2 + K2.y * v2);
    o[2] = (float)(t1.x + t2.x);
    o[3] = (float)(t1.y + t2.y);
    o[4] = (float)(t3.x + t4.x);
    o[5] = (float)(t3.y + t4.y);
    o[6] = (float)(K1.x * v2 + K2.x * u2);
    o[7] = (float)(K1.y * v2 + K2.y * u2);
}

// Compact table (fft_table_mode == 2; split transform with a radix-2 centre, i.e. log2M - 1 = 1 mod 4).
// With K1 = K'[k], K2 = conj(K'[M-k]), w = e^{-2 pi i k/n} = (wc, ws):  alpha = S + ws D, delta = S - ws D,
// beta = i wc D, gamma = -beta, where S = 2 (K1 + K2), D = 2 (K1 - K2).  Per half transform and centre item
// (groups ga = 2 item, gb = partner(ga)) two float4 (S, D): pair A led by slot 2 ga, pair B led by slot 2 gb;
// k_B = n/4 - k_A, so w_B = -i conj(w_A) and one twiddle per item suffices (pair_twiddle_kernel).
// float4 layout per channel: [0, 2 NI) even half, [2 NI, 4 NI) odd half, then 5 float4 in the 32 B form for
// item 0 of the even half: (K'[0], K'[M]), the self-paired slot 1 (two float4), group 1 (two float4).
__device__ inline void kcoef_full(const double2 *K, int M, int k, float *o) {
    const int n = 2 * M;
    const double2 K1 = K[k], K2 = make_double2(K[M - k].x, -K[M - k].y);
    double ws, wc;
    sincospi(-2.0 * (double)k / (double)n, &ws, &wc);
    const double S[2] = {2.0 * (K1.x + K2.x), 2.0 * (K1.y + K2.y)}, D[2] = {2.0 * (K1.x - K2.x), 2.0 * (K1.y - K2.y)};
    o[0] = (float)(S[0] + ws * D[0]);
    o[1] = (float)(S[1] + ws * D[1]);
    o[2] = (float)(-wc * D[1]);
    o[3] = (float)(wc * D[0]);
    o[4] = -o[2];
    o[5] = -o[3];
    o[6] = (float)(S[0] - ws * D[0]);
    o[7] = (float)(S[1] - ws * D[1]);
}
__global__ void kcoef_compact_kernel(const double2 *__restrict__ Kd, int log2M, float *__restrict__ kc) {
    const int M = 1 << log2M, LH = log2M - 1, Mh = M / 2, G = Mh / 2, NI = Mh / 4;
    const int t = blockIdx.x * blockDim.x + threadIdx.x, h = blockIdx.y;
    if (t >= 2 * NI) return;
    const int odd = t / NI, item = t - odd * NI;
    const double2 *K = Kd + (size_t)h * (M + 1);
    float *base = kc + (size_t)h * (M / 2 + 1) * 8;
    float *o = base + ((size_t)odd * 2 * NI + 2 * item) * 4;
    const int ga = 2 * item;
    const int gb = odd ? (ga ^ (G - 1)) : (ga == 0 ? 1 : ga ^ ((1 << (31 - __clz(ga))) - 1));
    const int lead[2] = {2 * ga, 2 * gb};
    for (int q = 0; q < 2; ++q) {
        const int k = 2 * fft_freq(lead[q], LH) + odd;
        if (k == 0) {                       // even half, item 0, pair A: DC / Nyquist live in the special block
            o[4 * q] = o[4 * q + 1] = o[4 * q + 2] = o[4 * q + 3] = 0.f;
            continue;
        }
        const double2 K1 = K[k], K2 = make_double2(K[M - k].x, -K[M - k].y);
        o[4 * q] = (float)(2.0 * (K1.x + K2.x));
        o[4 * q + 1] = (float)(2.0 * (K1.y + K2.y));
        o[4 * q + 2] = (float)(2.0 * (K1.x - K2.x));
        o[4 * q + 3] = (float)(2.0 * (K1.y - K2.y));
    }
    if (t == 0) {
        float *sp = base + (size_t)4 * NI * 4;
        sp[0] = (float)K[0].x;
        sp[1] = (float)K[M].x;
        sp[2] = sp[3] = 0.f;
        kcoef_full(K, M, 2 * fft_freq(1, LH), sp + 4);          // slot 1: k = M/2, self-paired
        kcoef_full(K, M, 2 * fft_freq(2, LH), sp + 12);         // group 1: slots 2 (leader), 3
    }
}
// W_n^{k_A(item)} for the even half, k_A = 2 bitrev_{LH}(4 item)
__global__ void pair_twiddle_kernel(float2 *tw2, int log2M) {
    const int M = 1 << log2M, LH = log2M - 1, NI = M / 8;
    const int item = blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= NI) return;
    const int k = 2 * fft_freq(4 * item, LH);
    double sn, cs;
    sincospi(-2.0 * (double)k / (double)(2 * M), &sn, &cs);
    tw2[item] = make_float2((float)cs, (float)sn);
}
int fft_pair_twiddles_launch(float2 *tw2, int log2M, cudaStream_t st) {
    const int NI = (1 << log2M) / 8;
    pair_twiddle_kernel<<<ceil_div(NI, 256), 256, 0, st>>>(tw2, log2M);
    DWB_LAUNCH_CHECK();
    return DWB_OK;
}

template <typename T>
static int fftconv_prepare_any(const T *k, const float *D, int H, int l, float *kc, cudaStream_t st, int force_mode = -1) {
    const int log2M = fft_log2m_for(l);
    DWB_REQUIRE(log2M > 0, DWB_ERR_UNSUPPORTED, "fftconv: stage length %d unsupported (max %d)", l, 1 << FFT_MAX_LOG2M);
    const int M = 1 << log2M;
    double2 *Kd = nullptr;
    DWB_CUDA(cudaMalloc(&Kd, (size_t)H * (M + 1) * sizeof(double2)));
    kf_kernel<T><<<dim3(ceil_div(M + 1, DFT_THREADS), H), DFT_THREADS, 0, st>>>(k, D, H, l, log2M, Kd);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) {
        const int mode = force_mode >= 0 ? force_mode : fft_table_mode(log2M, l);
        if (mode == 2) kcoef_compact_kernel<<<dim3(ceil_div(M / 2, 256), H), 256, 0, st>>>(Kd, log2M, kc);
        else kcoef_kernel<<<dim3(ceil_div(M / 2 + 1, 256), H), 256, 0, st>>>(Kd, log2M, mode, kc);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(Kd);
    if (e != cudaSuccess) return cuda_fail(e, "fftconv_prepare", __FILE__, __LINE__);
    return DWB_OK;
}

// ---------------------------------------------------------------------------------------
// standalone complex64 Cauchy op (drop-in for cauchy_mult_sym_fwd)
// ---------------------------------------------------------------------------------------
constexpr int CAUCHY_THREADS = 256;
constexpr int CAUCHY_NCHUNK = 256;

// LPL = lanes cooperating on one output l (split of the state dimension, warp-shuffle reduced)
template <int LPL>
__global__ void __launch_bounds__(CAUCHY_THREADS)
cauchy_sym_fwd_kernel(const float2 *__restrict__ v, const float2 *__restrict__ z, const float2 *__restrict__ w,
                      float2 *__restrict__ out, int N, int L) {
    __shared__ float2 sv[CAUCHY_NCHUNK], sw[CAUCHY_NCHUNK];
    const int b = blockIdx.y;
    const int lane = threadIdx.x % LPL;
    const int li = blockIdx.x * (CAUCHY_THREADS / LPL) + threadIdx.x / LPL;
    const bool live = li < L;
    const float2 zz = live ? z[li] : make_float2(0.f, 0.f);
    float ar = 0.f, ai = 0.f;
    for (int n0 = 0; n0 < N; n0 += CAUCHY_NCHUNK) {
        const int cnt = min(CAUCHY_NCHUNK, N - n0);
        __syncthreads();
        for (int i = threadIdx.x; i < cnt; i += CAUCHY_THREADS) {
            sv[i] = v[(size_t)b * N + n0 + i];
            sw[i] = w[(size_t)b * N + n0 + i];
        }
        __syncthreads();
        if (live) {
            for (int i = lane; i < cnt; i += LPL) {
                const float2 vn = sv[i], wn = sw[i];
                // v/(z-w) + conj(v)/(z-conj(w)); both denominators share (z.x - w.x)
                const float dx = zz.x - wn.x;
                const float dy1 = zz.y - wn.y, dy2 = zz.y + wn.y;
                // MUFU.RCP (1 ulp) instead of the IEEE division sequence: the kernel is instruction-issue bound
                // (ncu: issue-active 93 %), and two exact divisions were ~40 % of its instructions
                float i1, i2;
                asm("rcp.approx.f32 %0, %1;" : "=f"(i1) : "f"(fmaf(dx, dx, dy1 * dy1)));
                asm("rcp.approx.f32 %0, %1;" : "=f"(i2) : "f"(fmaf(dx, dx, dy2 * dy2)));
                // v * conj(d1) * i1 ; conj(v) * conj(d2) * i2
                ar += (vn.x * dx + vn.y * dy1) * i1 + (vn.x * dx - vn.y * dy2) * i2;
                ai += (vn.y * dx - vn.x * dy1) * i1 + (-vn.y * dx - vn.x * dy2) * i2;
            }
        }
    }
#pragma unroll
    for (int off = LPL / 2; off > 0; off >>= 1) {
        ar += __shfl_down_sync(0xffffffffu, ar, off, LPL);
        ai += __shfl_down_sync(0xffffffffu, ai, off, LPL);
    }
    if (live && lane == 0) out[(size_t)b * L + li] = make_float2(ar, ai);
}

// Many outputs: one thread owns LT outputs l (256 apart, so a warp's loads and stores stay coalesced) and walks the whole
// state dimension; every (v, w) pair read from shared memory is used LT times and the products v.x dx, v.y dx are
// shared by the two conjugate terms: 16 FP32 + 2 MUFU.RCP per (l, n) against ~55 instructions for the lane-split
// kernel above (ncu round 2: that one is issue bound at 93 %).
template <int LT>
__global__ void __launch_bounds__(CAUCHY_THREADS)
cauchy_sym_fwd_mt_kernel(const float2 *__restrict__ v, const float2 *__restrict__ z, const float2 *__restrict__ w,
                         float2 *__restrict__ out, int N, int L) {
    __shared__ float4 svw[CAUCHY_NCHUNK];
    const int b = blockIdx.y, l0 = blockIdx.x * (CAUCHY_THREADS * LT) + threadIdx.x;
    float2 zz[LT];
    float ar[LT], ai[LT];
#pragma unroll
    for (int j = 0; j < LT; ++j) {
        const int l = l0 + j * CAUCHY_THREADS;
        zz[j] = l < L ? z[l] : make_float2(2.f, 0.f);
        ar[j] = ai[j] = 0.f;
    }
    for (int n0 = 0; n0 < N; n0 += CAUCHY_NCHUNK) {
        const int cnt = min(CAUCHY_NCHUNK, N - n0);
        __syncthreads();
        for (int i = threadIdx.x; i < cnt; i += CAUCHY_THREADS) {
            const float2 a = v[(size_t)b * N + n0 + i], c = w[(size_t)b * N + n0 + i];
            svw[i] = make_float4(a.x, a.y, c.x, c.y);
        }
        __syncthreads();
#pragma unroll 2
        for (int i = 0; i < cnt; ++i) {
            const float4 q = svw[i];
#pragma unroll
            for (int j = 0; j < LT; ++j) {
                const float dx = zz[j].x - q.z, dy1 = zz[j].y - q.w, dy2 = zz[j].y + q.w, dx2 = dx * dx;
                float i1, i2;
                asm("rcp.approx.f32 %0, %1;" : "=f"(i1) : "f"(fmaf(dy1, dy1, dx2)));
                asm("rcp.approx.f32 %0, %1;" : "=f"(i2) : "f"(fmaf(dy2, dy2, dx2)));
                const float a = q.x * dx, c = q.y * dx;
                ar[j] = fmaf(fmaf(q.y, dy1, a), i1, fmaf(fmaf(-q.y, dy2, a), i2, ar[j]));
                ai[j] = fmaf(fmaf(-q.x, dy1, c), i1, fmaf(fmaf(-q.x, dy2, -c), i2, ai[j]));
            }
        }
    }
#pragma unroll
    for (int j = 0; j < LT; ++j) {
        const int l = l0 + j * CAUCHY_THREADS;
        if (l < L) out[(size_t)b * L + l] = make_float2(ar[j], ai[j]);
    }
}

}  // namespace dwb

using namespace dwb;

extern "C" int dwb_cauchy_sym_fwd(const float *v, const float *z, const float *w, float *out, int batch, int N,
                                  int L, void *stream) {
    DWB_REQUIRE(v && z && w && out, DWB_ERR_INVALID, "dwb_cauchy_sym_fwd: null pointer");
    DWB_REQUIRE(batch >= 0 && N >= 1 && L >= 0, DWB_ERR_INVALID, "dwb_cauchy_sym_fwd: bad sizes batch=%d N=%d L=%d",
                batch, N, L);
    DWB_REQUIRE(batch <= 65535, DWB_ERR_UNSUPPORTED, "dwb_cauchy_sym_fwd: batch %d > 65535", batch);
    if (batch == 0 || L == 0) return DWB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const float2 *v2 = (const float2 *)v, *z2 = (const float2 *)z, *w2 = (const float2 *)w;
    float2 *o2 = (float2 *)out;
    // many outputs: four outputs per thread, whole state dimension per thread
    if ((int64_t)batch * L >= 262144 && L >= 2 * CAUCHY_THREADS) {
        dim3 grid(ceil_div(L, CAUCHY_THREADS * 4), batch);
        cauchy_sym_fwd_mt_kernel<4><<<grid, CAUCHY_THREADS, 0, st>>>(v2, z2, w2, o2, N, L);
    } else
    // few outputs and many states: split the state dimension over a full warp; otherwise 4 lanes
    if ((int64_t)batch * L < 4096 && N >= 64) {
        dim3 grid(ceil_div(L, CAUCHY_THREADS / 32), batch);
        cauchy_sym_fwd_kernel<32><<<grid, CAUCHY_THREADS, 0, st>>>(v2, z2, w2, o2, N, L);
    } else if (N >= 8) {
        dim3 grid(ceil_div(L, CAUCHY_THREADS / 4), batch);
        cauchy_sym_fwd_kernel<4><<<grid, CAUCHY_THREADS, 0, st>>>(v2, z2, w2, o2, N, L);
    } else {
        dim3 grid(ceil_div(L, CAUCHY_THREADS), batch);
        cauchy_sym_fwd_kernel<1><<<grid, CAUCHY_THREADS, 0, st>>>(v2, z2, w2, o2, N, L);
    }
    DWB_LAUNCH_CHECK();
    return DWB_OK;
}

namespace dwb {
// used by the plan and by the single-op entry; khat/k64 are caller-provided fp64 scratch
int s4_generate(const float *C, const float *Bp, const float *P, const float *inv_w_real, const float *w_imag,
                const float *log_dt, const float *omega, int H, int N, int l, double *khat, double *k64,
                float *k32, cudaStream_t st, int64_t *launches) {
    const int nk = l / 2 + 1;
    DWB_REQUIRE(H <= 65535 / 2, DWB_ERR_UNSUPPORTED, "s4_generate: H=%d too large", H);
    {
        dim3 grid(ceil_div(nk, KHAT_THREADS), H);
        size_t smem = (size_t)N * 14 * sizeof(double);
        DWB_REQUIRE(smem <= 48 * 1024, DWB_ERR_UNSUPPORTED, "s4_generate: N=%d too large", N);
        s4_khat_kernel<<<grid, KHAT_THREADS, smem, st>>>(C, Bp, P, inv_w_real, w_imag, log_dt, omega, H, N, l, khat);
        DWB_LAUNCH_CHECK();
    }
    {
        dim3 grid(ceil_div(l, DFT_THREADS), 2 * H);
        irdft_kernel<<<grid, DFT_THREADS, 0, st>>>(khat, l, k64, k32);
        DWB_LAUNCH_CHECK();
    }
    if (launches) *launches += 2;
    return DWB_OK;
}

int fftconv_prepare_f64(const double *k64, const float *D, int H, int l, float *kf, cudaStream_t st) {
    return fftconv_prepare_any<double>(k64, D, H, l, kf, st);
}

int fftconv_prepare_f32(const float *k32, int ld, const float *D, int H, int l, int dir, float *kf, cudaStream_t st) {
    DWB_REQUIRE(l >= 1 && l <= ld && dir >= 0 && dir <= 2, DWB_ERR_INVALID, "fftconv_prepare_f32: l=%d ld=%d dir=%d", l, ld, dir);
    float *tmp = nullptr;                                        // (2,H,l): truncated rows, one direction zeroed
    DWB_CUDA(cudaMalloc(&tmp, (size_t)2 * H * l * sizeof(float)));
    cudaError_t e = cudaMemcpy2DAsync(tmp, (size_t)l * sizeof(float), k32, (size_t)ld * sizeof(float), (size_t)l * sizeof(float),
                                      (size_t)2 * H, cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess && dir == 1) e = cudaMemsetAsync(tmp + (size_t)H * l, 0, (size_t)H * l * sizeof(float), st);
    if (e == cudaSuccess && dir == 2) e = cudaMemsetAsync(tmp, 0, (size_t)H * l * sizeof(float), st);
    int rc = e == cudaSuccess ? fftconv_prepare_any<float>(tmp, dir == 2 ? nullptr : D, H, l, kf, st, dir == 0 ? -1 : 0)
                              : cuda_fail(e, "fftconv_prepare_f32", __FILE__, __LINE__);
    cudaStreamSynchronize(st);
    cudaFree(tmp);
    return rc;
}
}  // namespace dwb

extern "C" int dwb_s4_kernel_gen(const float *C, const float *Bp, const float *P, const float *inv_w_real,
                                 const float *w_imag, const float *log_dt, const float *omega, int H, int N, int l,
                                 float *k_out, void *stream) {
    DWB_REQUIRE(C && Bp && P && inv_w_real && w_imag && log_dt && k_out, DWB_ERR_INVALID, "dwb_s4_kernel_gen: null pointer");
    DWB_REQUIRE(H >= 1 && N >= 1 && l >= 2, DWB_ERR_INVALID, "dwb_s4_kernel_gen: bad sizes H=%d N=%d l=%d", H, N, l);
    cudaStream_t st = (cudaStream_t)stream;
    const int nk = l / 2 + 1;
    double *khat = nullptr, *k64 = nullptr;
    DWB_CUDA(cudaMalloc(&khat, (size_t)2 * H * nk * 2 * sizeof(double)));
    cudaError_t e = cudaMalloc(&k64, (size_t)2 * H * l * sizeof(double));
    if (e != cudaSuccess) {
        cudaFree(khat);
        return cuda_fail(e, "cudaMalloc k64", __FILE__, __LINE__);
    }
    int rc = s4_generate(C, Bp, P, inv_w_real, w_imag, log_dt, omega, H, N, l, khat, k64, k_out, st, nullptr);
    cudaError_t es = cudaStreamSynchronize(st);
    cudaFree(khat);
    cudaFree(k64);
    if (rc != DWB_OK) return rc;
    if (es != cudaSuccess) return cuda_fail(es, "dwb_s4_kernel_gen sync", __FILE__, __LINE__);
    return DWB_OK;
}

extern "C" int dwb_fftconv_size(int l, int *nfft) {
    DWB_REQUIRE(nfft, DWB_ERR_INVALID, "dwb_fftconv_size: null");
    int log2M = fft_log2m_for(l);
    DWB_REQUIRE(log2M > 0, DWB_ERR_UNSUPPORTED, "fftconv: stage length %d unsupported (max %d)", l, 1 << FFT_MAX_LOG2M);
    *nfft = 2 << log2M;
    return DWB_OK;
}

extern "C" int dwb_fftconv_prepare(const float *k, const float *D, int H, int l, float *kf, void *stream) {
    DWB_REQUIRE(k && kf, DWB_ERR_INVALID, "dwb_fftconv_prepare: null pointer");
    DWB_REQUIRE(H >= 1 && l >= 1, DWB_ERR_INVALID, "dwb_fftconv_prepare: bad sizes H=%d l=%d", H, l);
    return dwb::fftconv_prepare_any<float>(k, D, H, l, kf, (cudaStream_t)stream);
}
