// Shared helpers for libdwb (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/dwb.h"

namespace dwb {

// thread-local error string behind dwb_last_error()
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define DWB_CUDA(expr)                                                          \
    do {                                                                        \
        cudaError_t _e = (expr);                                                \
        if (_e != cudaSuccess) return dwb::cuda_fail(_e, #expr, __FILE__, __LINE__); \
    } while (0)

#define DWB_LAUNCH_CHECK() DWB_CUDA(cudaGetLastError())

#define DWB_REQUIRE(cond, code, ...)      \
    do {                                  \
        if (!(cond)) {                    \
            dwb::set_error(__VA_ARGS__);  \
            return (code);                \
        }                                 \
    } while (0)

// launch accounting (plan->launches); kernels launched through LAUNCH bump the counter
struct LaunchCounter {
    int64_t n = 0;
};

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ float gelu_erf(float x) {
    // torch.nn.functional.gelu (erf form): 0.5 x (1 + erf(x / sqrt 2))
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
// GELU (erf form) through erfc(|x|/sqrt2) = t(a1 + t(a2 + ...)) exp(-x^2/2), t = 1/(1 + p|x|/sqrt2)
// (Abramowitz-Stegun 7.1.26): branch free, 2 MUFU + ~12 FP32 ops instead of erff's ~40 with
// divergent ranges.  |error| <= 3.4e-7 absolute over the real line (checked against float64 erf),
// far inside the 1e-3 parity budget; the exact-fp32 SIMT path keeps erff.
__device__ __forceinline__ float gelu_fast(float x) {
    const float z = fabsf(x) * 0.70710678118654752440f;
    float t;                                          // MUFU.RCP: 1 ulp, no IEEE slow path / branch
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
    float p = fmaf(t, 1.061405429f, -1.453152027f);
    p = fmaf(t, p, 1.421413741f);
    p = fmaf(t, p, -0.284496736f);
    p = fmaf(t, p, 0.254829592f);
    float e;                                          // exp(-z^2) = 2^(-z^2 log2 e)
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * z * -1.4426950408889634f));
    e *= p * t;                                       // erfc(|x|/sqrt 2)
    const float h = 0.5f * x * e;                    // x < 0: 0.5 x (1 + erf) = 0.5 x erfc
    return x >= 0.f ? x - h : h;
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cmul_conj(float2 a, float2 b) {  // a * conj(b)
    return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }

}  // namespace dwb
