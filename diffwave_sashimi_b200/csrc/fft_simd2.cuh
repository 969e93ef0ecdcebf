// Packed fp32 (f32x2) arithmetic for the FFT kernels: sm_100a executes add/mul/fma.f32x2 as ONE
// warp instruction (FADD2 / FMUL2 / FFMA2, two fp32 lanes per 64-bit register pair).  The FP32 pipe
// time per flop is unchanged (measured: 0.48 FFMA2/clk/SMSP vs 0.97 FFMA) but the instruction
// halves, so the freed issue slots carry the kernel's address arithmetic and shared-memory traffic
// (tools/ffma2_probe.cu: FFMA2 + IADD pairs retire in 2.37 cycles where 2 FFMA + IADD need 3).
//
// The two lanes are two INDEPENDENT butterflies (adjacent j of the same pass), never the real and
// imaginary part of one number: every scalar operation of the butterfly then maps 1:1 onto a packed
// one, rotations by -+i stay register renames, and no lane swizzle is ever needed.
#pragma once
#include <cuda_runtime.h>

#include "common.cuh"

namespace dwb {
namespace s2 {

typedef unsigned long long u64;

struct V2 {          // two fp32 lanes
    float2 v;
    __device__ __forceinline__ V2() {}
    __device__ __forceinline__ explicit V2(float a) : v(make_float2(a, a)) {}
    __device__ __forceinline__ V2(float a, float b) : v(make_float2(a, b)) {}
    __device__ __forceinline__ explicit V2(float2 a) : v(a) {}
};
__device__ __forceinline__ u64 bits(const V2 &a) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.v.x), "f"(a.v.y));
    return r;
}
__device__ __forceinline__ V2 unbits(u64 r) {
    V2 a;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a.v.x), "=f"(a.v.y) : "l"(r));
    return a;
}
__device__ __forceinline__ V2 operator+(const V2 &a, const V2 &b) {
    u64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(bits(a)), "l"(bits(b)));
    return unbits(r);
}
__device__ __forceinline__ V2 operator-(const V2 &a, const V2 &b) {
    u64 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(bits(a)), "l"(bits(b)));
    return unbits(r);
}
__device__ __forceinline__ V2 operator*(const V2 &a, const V2 &b) {
    u64 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(bits(a)), "l"(bits(b)));
    return unbits(r);
}
__device__ __forceinline__ V2 operator*(const V2 &a, float c) { return a * V2(c); }
__device__ __forceinline__ V2 operator-(const V2 &a) { return V2(-a.v.x, -a.v.y); }   // folded into the consumer's operand modifier
__device__ __forceinline__ V2 fma(const V2 &a, const V2 &b, const V2 &c) {          // a * b + c
    u64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(bits(a)), "l"(bits(b)), "l"(bits(c)));
    return unbits(r);
}

struct C2 {          // one complex number per lane
    V2 x, y;
};
__device__ __forceinline__ C2 cmul(const C2 &a, const C2 &b) {
    C2 r;
    r.x = fma(a.y, -b.y, a.x * b.x);
    r.y = fma(a.y, b.x, a.x * b.y);
    return r;
}
__device__ __forceinline__ C2 cconj(const C2 &a) {
    C2 r;
    r.x = a.x;
    r.y = -a.y;
    return r;
}

// ---- radix-R DFT in registers, natural order in and out (same recursion as Radix<> in fft_radix.cuh)
// ZHI: inputs x[R/2..R) are zero and are not read
template <int R, bool INV, bool ZHI = false>
struct RadixS {
    static __device__ __forceinline__ void run(C2 *x) {
        if constexpr (R == 2 && ZHI) {
            x[1] = x[0];
            return;
        }
        constexpr float WR[8] = {1.0f, 0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f,
                                 0.0f, -0.38268343236508977f, -0.70710678118654752f, -0.92387953251128674f};
        constexpr float WI[8] = {0.0f, -0.38268343236508977f, -0.70710678118654752f, -0.92387953251128674f,
                                 -1.0f, -0.92387953251128674f, -0.70710678118654752f, -0.38268343236508977f};
        C2 e[R / 2], o[R / 2];
#pragma unroll
        for (int i = 0; i < (ZHI ? R / 4 : R / 2); ++i) {
            e[i] = x[2 * i];
            o[i] = x[2 * i + 1];
        }
        RadixS<R / 2, INV, ZHI>::run(e);
        RadixS<R / 2, INV, ZHI>::run(o);
#pragma unroll
        for (int q = 0; q < R / 2; ++q) {
            constexpr int step = 16 / R;
            const int k = q * step;
            constexpr float h = 0.70710678118654752f;
            C2 t;
            if (k == 0) {
                t = o[q];
            } else if (k == 4) {                       // -+ i: a rename
                if (INV) { t.x = -o[q].y; t.y = o[q].x; } else { t.x = o[q].y; t.y = -o[q].x; }
            } else if (k == 2) {
                if (INV) { t.x = (o[q].x - o[q].y) * h; t.y = (o[q].x + o[q].y) * h; }
                else { t.x = (o[q].x + o[q].y) * h; t.y = (o[q].y - o[q].x) * h; }
            } else if (k == 6) {
                if (INV) { t.x = (o[q].x + o[q].y) * (-h); t.y = (o[q].x - o[q].y) * h; }
                else { t.x = (o[q].y - o[q].x) * h; t.y = (o[q].x + o[q].y) * (-h); }
            } else {
                const float wr = WR[k], wi = INV ? -WI[k] : WI[k];
                t.x = fma(o[q].y, V2(-wi), o[q].x * wr);
                t.y = fma(o[q].y, V2(wr), o[q].x * wi);
            }
            x[q].x = e[q].x + t.x;
            x[q].y = e[q].y + t.y;
            x[q + R / 2].x = e[q].x - t.x;
            x[q + R / 2].y = e[q].y - t.y;
        }
    }
};
template <bool INV, bool ZHI>
struct RadixS<1, INV, ZHI> {
    static __device__ __forceinline__ void run(C2 *) {}
};

// x[q] *= u0 v^q (HASBASE) or v^q, q < 16
template <bool HASBASE>
__device__ __forceinline__ void apply_twiddles16(C2 (&x)[16], const C2 u0, const C2 v) {
    C2 u[8];
    const C2 v2 = cmul(v, v), v4 = cmul(v2, v2), v8 = cmul(v4, v4);
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

// x[q] *= v^q, q < 4
__device__ __forceinline__ void apply_twiddles4(C2 (&x)[4], const C2 v) {
    const C2 v2 = cmul(v, v);
    x[1] = cmul(x[1], v);
    x[2] = cmul(x[2], v2);
    x[3] = cmul(x[3], cmul(v2, v));
}

// x[p] *= W_32^{+-p}, p < 16 (forward: -, inverse: +)
template <bool INV>
__device__ __forceinline__ void rotate_w32(C2 (&x)[16]) {
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
        C2 r;
        if (p == 8) {
            if (INV) { r.x = -x[p].y; r.y = x[p].x; } else { r.x = x[p].y; r.y = -x[p].x; }
        } else {
            const float wr = C[p], wi = INV ? S[p] : -S[p];
            r.x = fma(x[p].y, V2(-wi), x[p].x * wr);
            r.y = fma(x[p].y, V2(wr), x[p].x * wi);
        }
        x[p] = r;
    }
}


// ---- elementwise activations, two lanes (MUFU and the final selects stay scalar) ---------------------
// GELU (erf form): same formula as gelu_fast (common.cuh): 0.5 x + |x| (0.5 - 2^-Q(|x|)); 5 FFMA2 + FADD2 + FMUL2 +
// FFMA2 packed, the two |x| and the two MUFU.EX2 scalar
__device__ __forceinline__ V2 gelu_fast2(const V2 &x) {
    const V2 a(fabsf(x.v.x), fabsf(x.v.y));
    V2 q = fma(a, V2(DWB_GELU_Q5), V2(DWB_GELU_Q4));
    q = fma(a, q, V2(DWB_GELU_Q3));
    q = fma(a, q, V2(DWB_GELU_Q2));
    q = fma(a, q, V2(DWB_GELU_Q1));
    q = fma(a, q, V2(DWB_GELU_Q0));
    float h0, h1;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(h0) : "f"(q.v.x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(h1) : "f"(q.v.y));
    return fma(a, V2(0.5f) - V2(h0, h1), x * 0.5f);
}

// 1 / (1 + exp(-x)), same approximations as __fdividef(1, 1 + __expf(-x))
__device__ __forceinline__ V2 sigmoid_fast2(const V2 &x) {
    const V2 ea = x * (-1.4426950408889634f);
    float e0, e1, r0, r1;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(ea.v.x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(ea.v.y));
    const V2 d = V2(e0, e1) + V2(1.0f);
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(d.v.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(d.v.y));
    return V2(r0, r1);
}

// sigmoid of a gate whose bias arrives pre-multiplied by -log2(e):  1 / (1 + 2^(-log2e x + bs))
__device__ __forceinline__ V2 sigmoid_scaled2(const V2 &x, const V2 &bs) {
    const V2 ea = fma(x, V2(-1.4426950408889634f), bs);
    float e0, e1, r0, r1;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(ea.v.x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(ea.v.y));
    const V2 d = V2(e0, e1) + V2(1.0f);
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(d.v.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(d.v.y));
    return V2(r0, r1);
}

}  // namespace s2
}  // namespace dwb
