// In-register radix butterflies, twiddle application and the real-FFT pair map shared by the fftconv
// kernel variants (complex numbers as float2, scalar fp32 arithmetic).
#pragma once
#include "common.cuh"
#include "fft_plan.cuh"
#include "fft_simd2.cuh"

namespace dwb {

// ---- in-register radix-R DFT, natural order in and out --------------------------------
// ZHI: inputs x[R/2..R) are zero and are not read
template <int R, bool INV, bool ZHI = false>
struct Radix {
    static __device__ __forceinline__ void run(float2 *x) {
        // omega_16^q = exp(-2 pi i q / 16), q = 0..7
        constexpr float WR[8] = {1.0f, 0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f,
                                 0.0f, -0.38268343236508977f, -0.70710678118654752f, -0.92387953251128674f};
        constexpr float WI[8] = {0.0f, -0.38268343236508977f, -0.70710678118654752f, -0.92387953251128674f,
                                 -1.0f, -0.92387953251128674f, -0.70710678118654752f, -0.38268343236508977f};
        if constexpr (R == 2 && ZHI) {
            x[1] = x[0];
        } else {
            float2 e[R / 2], o[R / 2];
#pragma unroll
            for (int i = 0; i < (ZHI ? R / 4 : R / 2); ++i) {
                e[i] = x[2 * i];
                o[i] = x[2 * i + 1];
            }
            Radix<R / 2, INV, ZHI>::run(e);
            Radix<R / 2, INV, ZHI>::run(o);
#pragma unroll
            for (int q = 0; q < R / 2; ++q) {
                // t = o[q] * omega_R^{+-q}; the trivial rotations (1, -+i, sqrt(1/2)(+-1 -+ i)) are spelled out:
                // "x * 0.0f" cannot be folded by the compiler and would cost an FFMA each
                constexpr int step = 16 / R;
                const int k = q * step;              // omega_16 exponent, 0..7 (compile time after unrolling)
                constexpr float h = 0.70710678118654752f;
                float2 t;
                if (k == 0) {
                    t = o[q];
                } else if (k == 4) {
                    t = INV ? make_float2(-o[q].y, o[q].x) : make_float2(o[q].y, -o[q].x);
                } else if (k == 2) {
                    t = INV ? make_float2(h * (o[q].x - o[q].y), h * (o[q].x + o[q].y))
                            : make_float2(h * (o[q].x + o[q].y), h * (o[q].y - o[q].x));
                } else if (k == 6) {
                    t = INV ? make_float2(-h * (o[q].x + o[q].y), h * (o[q].x - o[q].y))
                            : make_float2(h * (o[q].y - o[q].x), -h * (o[q].x + o[q].y));
                } else {
                    const float wr = WR[k], wi = INV ? -WI[k] : WI[k];
                    t = make_float2(o[q].x * wr - o[q].y * wi, o[q].x * wi + o[q].y * wr);
                }
                if (k == 4) {                        // t is a half-negated swap: scalar adds, no register moves
                    x[q] = make_float2(e[q].x + t.x, e[q].y + t.y);
                    x[q + R / 2] = make_float2(e[q].x - t.x, e[q].y - t.y);
                } else {                             // complex add / sub = one FADD2 each (packed fp32, sm_100a)
                    x[q] = (s2::V2(e[q]) + s2::V2(t)).v;
                    x[q + R / 2] = (s2::V2(e[q]) - s2::V2(t)).v;
                }
            }
        }
    }
};
template <bool INV, bool ZHI>
struct Radix<1, INV, ZHI> {
    static __device__ __forceinline__ void run(float2 *) {}
};

// w[q] = w1^q, q = 1..R-1, by squaring/products of depth log2 R (a serial chain w *= w1 puts R-1
// dependent complex multiplies on the critical path of every butterfly)
template <int R>
__device__ __forceinline__ void twiddle_powers(float2 w1, float2 (&w)[R]) {
    w[0] = make_float2(1.f, 0.f);
    if (R > 1) w[1] = w1;
#pragma unroll
    for (int q = 2; q < R; ++q) {
        int hb = 1;
        while (hb * 2 <= q) hb *= 2;
        w[q] = (q == hb) ? cmul(w[q / 2], w[q / 2]) : cmul(w[hb], w[q - hb]);
    }
}

// x[q] *= w1^q, q < R, keeping at most R/2 powers live (w^1..w^{R/2-1}, then w^{R/2} times those)
template <int R>
__device__ __forceinline__ void apply_twiddles(float2 (&x)[R], float2 w1) {
    if constexpr (R >= 8) {
        float2 w[R / 2];
        twiddle_powers<R / 2>(w1, w);
#pragma unroll
        for (int q = 1; q < R / 2; ++q) x[q] = cmul(x[q], w[q]);
        const float2 wh = cmul(w[R / 4], w[R / 4]);
        x[R / 2] = cmul(x[R / 2], wh);
#pragma unroll
        for (int q = 1; q < R / 2; ++q) x[R / 2 + q] = cmul(x[R / 2 + q], cmul(wh, w[q]));
    } else {
        float2 w[R];
        twiddle_powers<R>(w1, w);
#pragma unroll
        for (int q = 1; q < R; ++q) x[q] = cmul(x[q], w[q]);
    }
}

// ---- real-FFT untangle + spectrum product + re-tangle of one pair ------------------------------
// a = Z[k] (slot p, k < M/2), b = Z[M-k] (slot p2);  c0 = (alpha, beta), c1 = (gamma, delta):
//   Z'[k] = alpha a + beta conj(b),  Z'[M-k] = conj(gamma a + delta conj(b))
__device__ __forceinline__ void pair_map(float2 &a, float2 &b, const float4 c0, const float4 c1) {
    const float2 oa = make_float2(c0.x * a.x - c0.y * a.y + c0.z * b.x + c0.w * b.y,
                                  c0.x * a.y + c0.y * a.x + c0.w * b.x - c0.z * b.y);
    const float2 ob = make_float2(c1.x * a.x - c1.y * a.y + c1.z * b.x + c1.w * b.y,
                                  -(c1.x * a.y + c1.y * a.x + c1.w * b.x - c1.z * b.y));
    a = oa;
    b = ob;
}
// x[q] *= u0 v^q (HASBASE) or v^q, q < 16
template <bool HASBASE>
__device__ __forceinline__ void apply_twiddles16(float2 (&x)[16], const float2 u0, const float2 v) {
    float2 u[8];
    const float2 v2 = cmul(v, v), v4 = cmul(v2, v2), v8 = cmul(v4, v4);
    u[0] = u0;
    u[1] = HASBASE ? cmul(u0, v) : v;
    u[2] = HASBASE ? cmul(u0, v2) : v2;
    u[3] = cmul(u[1], v2);
    u[4] = HASBASE ? cmul(u0, v4) : v4;
    u[5] = cmul(u[1], v4);
    u[6] = cmul(u[2], v4);
    u[7] = cmul(u[3], v4);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        if (HASBASE || q > 0) x[q] = cmul(x[q], u[q]);
        x[8 + q] = cmul(x[8 + q], (HASBASE || q > 0) ? cmul(v8, u[q]) : v8);
    }
}

// x[p] *= W_32^{+-p}, p < 16 (forward: -, inverse: +)
template <bool INV>
__device__ __forceinline__ void rotate_w32(float2 (&x)[16]) {
    constexpr float C[16] = {1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
                             0.70710678118654752f, 0.55557023301960218f, 0.38268343236508977f, 0.19509032201612825f,
                             0.0f, -0.19509032201612825f, -0.38268343236508977f, -0.55557023301960218f,
                             -0.70710678118654752f, -0.83146961230254524f, -0.92387953251128674f, -0.98078528040323043f};
    constexpr float S[16] = {0.0f, 0.19509032201612825f, 0.38268343236508977f, 0.55557023301960218f,
                             0.70710678118654752f, 0.83146961230254524f, 0.92387953251128674f, 0.98078528040323043f,
                             1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
                             0.70710678118654752f, 0.55557023301960218f, 0.38268343236508977f, 0.19509032201612825f};
#pragma unroll
    for (int p = 1; p < 16; ++p) {
        if (p == 8) {
            x[p] = INV ? make_float2(-x[p].y, x[p].x) : make_float2(x[p].y, -x[p].x);
        } else {
            const float wr = C[p], wi = INV ? S[p] : -S[p];
            x[p] = make_float2(x[p].x * wr - x[p].y * wi, x[p].x * wi + x[p].y * wr);
        }
    }
}

__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }


}  // namespace dwb
