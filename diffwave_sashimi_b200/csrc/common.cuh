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

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------------
// A kernel launched through launch_pdl may begin while its predecessor in the stream is still running: its CTAs are
// scheduled as soon as every CTA of the predecessor has executed pdl_trigger() (or exited) and resources are free, run
// their input-independent set-up (barrier init, TMEM allocation, weight / twiddle copies) and then block in
// pdl_wait() until the predecessor has completed and its writes are visible.  EVERY kernel launched this way must
// call pdl_wait() before it reads or writes anything a predecessor touches; both calls are no-ops in a normal launch.
// DWB_PDL=1 turns the attribute on; the default is plain stream order (measured faster, see pdl_enabled in api.cu).
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ float gelu_erf(float x) {
    // torch.nn.functional.gelu (erf form): 0.5 x (1 + erf(x / sqrt 2))
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
// GELU (erf form) as  0.5 x + |x| (0.5 - Phi(-|x|)),  Phi(-a) = 2^-Q(a)  with Q a degree-5 minimax
// polynomial in a = |x| (weighted fit of -log2 Phi(-a) on [0, 8], tools/fit_gelu.py; Q is positive and
// increasing on the whole half line, so no clamp is needed: large |x| underflows 2^-Q to 0).
// Branch free, 1 MUFU + 8 FP32 ops (the previous Abramowitz-Stegun 7.1.26 form: 2 MUFU + ~14 and a
// select).  |error| <= 1.1e-6 absolute, 3.3e-7 rms over |x| < 4 in fp32 (checked against float64
// erfc), far inside the 1e-3 parity budget; the exact-fp32 SIMT path keeps erff.
#define DWB_GELU_Q0 -1.000034731f
#define DWB_GELU_Q1 -1.150812285f
#define DWB_GELU_Q2 -4.599330676e-01f
#define DWB_GELU_Q3 -5.188627531e-02f
#define DWB_GELU_Q4 7.109689777e-03f
#define DWB_GELU_Q5 -4.770795488e-04f
__device__ __forceinline__ float gelu_fast(float x) {
    const float a = fabsf(x);
    float q = fmaf(a, DWB_GELU_Q5, DWB_GELU_Q4);      // -Q(a)
    q = fmaf(a, q, DWB_GELU_Q3);
    q = fmaf(a, q, DWB_GELU_Q2);
    q = fmaf(a, q, DWB_GELU_Q1);
    q = fmaf(a, q, DWB_GELU_Q0);
    float h;                                          // Phi(-|x|)
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(h) : "f"(q));
    return fmaf(a, 0.5f - h, 0.5f * x);
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
