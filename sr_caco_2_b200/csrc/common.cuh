// Shared helpers for libsrk (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/srk.h"

namespace srk {

// ---- error plumbing -------------------------------------------------------------------
extern thread_local char g_err[512];
extern thread_local long long g_launches;
int fail(int code, const char* fmt, ...);

#define SRK_REQUIRE(cond, ...)                                        \
    do {                                                              \
        if (!(cond)) return ::srk::fail(SRK_ERR_INVALID, __VA_ARGS__); \
    } while (0)

#define SRK_CUDA(expr)                                                               \
    do {                                                                             \
        cudaError_t e__ = (expr);                                                    \
        if (e__ != cudaSuccess)                                                      \
            return ::srk::fail(SRK_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,        \
                               cudaGetErrorString(e__), __FILE__, __LINE__);         \
    } while (0)

#define SRK_LAUNCH_CHECK(name)                                                       \
    do {                                                                             \
        ::srk::g_launches++;                                                         \
        cudaError_t e__ = cudaGetLastError();                                        \
        if (e__ != cudaSuccess)                                                      \
            return ::srk::fail(SRK_ERR_CUDA, "launch of %s failed: %s", name,        \
                               cudaGetErrorString(e__));                             \
    } while (0)

// ---- optional event timing per kernel family (see srk_profile) ---------------------------
extern thread_local bool g_prof_on;
void prof_begin(int family, cudaStream_t st);
void prof_end(cudaStream_t st);
struct ProfScope {
    cudaStream_t st; bool on;
    ProfScope(int family, void* stream) : st((cudaStream_t)stream), on(g_prof_on) { if (on) prof_begin(family, st); }
    ~ProfScope() { if (on) prof_end(st); }
};

// per-device one-shot flag for cudaFuncSetAttribute (function attributes are per device): returns true the
// first time `flags` is asked about the current device
static inline bool first_use_on_device(bool (&flags)[64]) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return true;
    if (flags[dev]) return false;
    flags[dev] = true;
    return true;
}
static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

// ---- 16-bit conversions ----------------------------------------------------------------
__device__ __forceinline__ uint16_t f2h16(float v, int dtype) {
    if (dtype == SRK_BF16) {
        __nv_bfloat16 b = __float2bfloat16_rn(v);
        return *reinterpret_cast<uint16_t*>(&b);
    }
    // saturating fp16 (finite): conv operands are the unnormalised residual stream
    v = fminf(fmaxf(v, -65504.f), 65504.f);
    __half h = __float2half_rn(v);
    return *reinterpret_cast<uint16_t*>(&h);
}
__device__ __forceinline__ uint32_t pack2(float a, float b, int dtype) {
    return (uint32_t)f2h16(a, dtype) | ((uint32_t)f2h16(b, dtype) << 16);
}
__device__ __forceinline__ float h162f(uint16_t u, int dtype) {
    if (dtype == SRK_BF16) return __uint_as_float(((uint32_t)u) << 16);
    __half h = *reinterpret_cast<__half*>(&u);
    return __half2float(h);
}

// ---- index maps (the single definition every kernel uses) -----------------------------------
// window-major position m (within one image of H x W tokens, 8x8 windows, cyclic shift s)
// -> token id h*W + w of the un-shifted frame.   roll(-s) + window_partition.
__device__ __host__ __forceinline__ int win_pos_to_token(int m, int H, int W, int shift) {
    const int wpr = W >> 3;
    const int win = m >> 6, pos = m & 63;
    const int wi = win / wpr, wj = win - wi * wpr;
    int h = (wi << 3) + (pos >> 3) + shift;
    int w = (wj << 3) + (pos & 7) + shift;
    if (h >= H) h -= H;
    if (w >= W) w -= W;
    return h * W + w;
}
// inverse: token id -> window-major position under cyclic shift s (used when an epilogue
// writes rows for the NEXT block's shifted windows)
__device__ __host__ __forceinline__ int token_to_win_pos(int tok, int H, int W, int shift) {
    int h = tok / W, w = tok - h * W;
    h -= shift; w -= shift;
    if (h < 0) h += H;
    if (w < 0) w += W;
    return (((h >> 3) * (W >> 3) + (w >> 3)) << 6) + ((h & 7) << 3) + (w & 7);
}
// region label of a coordinate of the shifted frame along one axis (calculate_mask)
__device__ __host__ __forceinline__ int region_label(int c, int n, int shift) {
    return c < n - 8 ? 0 : (c < n - shift ? 1 : 2);
}
// label of window-major position (win, pos) in the shifted frame
__device__ __host__ __forceinline__ int win_pos_label(int win, int pos, int H, int W, int shift) {
    const int wpr = W >> 3;
    const int wi = win / wpr, wj = win - wi * wpr;
    return region_label((wi << 3) + (pos >> 3), H, shift) * 3 +
           region_label((wj << 3) + (pos & 7), W, shift);
}
// relative position index for window-local positions i (query), j (key), ws = 8
__device__ __host__ __forceinline__ int rel_pos_index(int i, int j) {
    return ((i >> 3) - (j >> 3) + 7) * 15 + ((i & 7) - (j & 7) + 7);
}

// erf by Abramowitz & Stegun 7.1.26 (|abs error| <= 1.5e-7 on the whole real line): one
// MUFU.RCP + one MUFU.EX2 + 7 FMAs instead of the ~40-instruction branchy libm erff.  The GELU
// output is rounded to bf16 (relative 4e-3) right after, so this is the exact-erf GELU of
// nn.GELU() (network_swinir.py:30) to far below one output ulp.
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// ---- packed fp32x2 arithmetic (sm_100 FFMA2 / FMUL2: two lanes of fp32 per instruction) ----------
__device__ __forceinline__ uint64_t pack64(uint32_t lo, uint32_t hi) {
    uint64_t d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
    return d;
}
__device__ __forceinline__ void unpack64(uint64_t v, uint32_t& lo, uint32_t& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t dup2(float f) { return pack64(__float_as_uint(f), __float_as_uint(f)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// Two exact-erf GELUs (nn.GELU(), network_swinir.py:30) of x = acc + bias, given the accumulator pair
// and the HALVED bias pair.  erf by Abramowitz & Stegun 7.1.28,
//     erf(z) = 1 - 1 / (1 + a1 z + ... + a6 z^6)^16,   |abs error| <= 3e-7 (z >= 0),
// which needs one MUFU (the reciprocal) per element and no exponential; everything else is packed
// FFMA2 / FMUL2.  With hx = x/2 and nh = -|hx| (z = -sqrt(2) nh, coefficients pre-multiplied):
//     gelu(x) = hx + |hx| erf(z) = nh * r + (hx - nh),   r = 1 / P(nh)^16.
// ~9 instructions per element instead of ~20 for the scalar 7.1.26 form (fast_erf below), max abs
// error 7e-7 over [-12, 12]; the result is rounded to bf16 (4e-3 relative) right after.
__device__ __forceinline__ uint64_t gelu2(uint64_t acc, uint64_t hbias) {
    const uint64_t hx = fma2(acc, dup2(0.5f), hbias);
    const uint64_t nh = hx | 0x8000000080000000ull;
    uint64_t h = fma2(nh, dup2(3.4451039391569793e-4f), dup2(-1.5645003877580166e-3f));
    h = fma2(nh, h, dup2(6.080571911297739e-4f));
    h = fma2(nh, h, dup2(-2.6221010833978653e-2f));
    h = fma2(nh, h, dup2(8.456402271986008e-2f));
    h = fma2(nh, h, dup2(-9.973469376564026e-2f));
    uint64_t q = fma2(nh, h, dup2(1.f));
    q = mul2(q, q); q = mul2(q, q); q = mul2(q, q); q = mul2(q, q);
    uint32_t q0, q1;
    unpack64(q, q0, q1);
    const uint64_t r = pack64(__float_as_uint(rcp_approx(__uint_as_float(q0))), __float_as_uint(rcp_approx(__uint_as_float(q1))));
    return fma2(nh, r, fma2(nh, dup2(-1.f), hx));
}
__device__ __forceinline__ float fast_erf(float x) {
    const float ax = fabsf(x);
    const float t = rcp_approx(fmaf(0.3275911f, ax, 1.f));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    p *= t;
    const float e = ex2_approx(-ax * ax * 1.4426950408889634f);
    return copysignf(fmaf(-p, e, 1.f), x);
}
__device__ __forceinline__ float gelu_erf(float x) {
    return 0.5f * x * (1.f + fast_erf(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == SRK_ACT_GELU) return gelu_erf(v);
    if (act == SRK_ACT_LRELU) return v > 0.f ? v : 0.01f * v;
    if (act == SRK_ACT_LRELU02) return v > 0.f ? v : 0.2f * v;
    if (act == SRK_ACT_RELU) return fmaxf(v, 0.f);
    return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// engine entry points (one per translation unit)
int gemm_mma_sync(const srk_gemm_args* g, cudaStream_t st);
int gemm_tcgen05(const srk_gemm_args* g, cudaStream_t st);
int validate_gemm(const srk_gemm_args* g);

}  // namespace srk
