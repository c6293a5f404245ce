// Epilogue / addressing definitions shared by the two GEMM engines.
#pragma once
#include "common.cuh"

namespace srk {

// device-side copy of srk_gemm_args plus derived values
struct GemmP {
    const uint16_t* A; int a_mode; int lda; int nB, H, W;
    const uint16_t* Wt; int M, N, K; int dtype;
    const float* bias; int act;
    const float* res; float res_scale; float* out32; int ld32; int win_shift;
    uint16_t* out16; int ld16; int out16_dtype; int out16_mode;
    const float* ln_g; const float* ln_b; int ln_C; int ln_win_shift;
    float* img; int img_s; float img_scale; int img_hc, img_wc;
    const float* attn_table; int attn_heads; float attn_scale; int attn_shift;
    int T;            // H*W (tokens per image) when H, W are set, else 0
    int cpb;          // channel blocks of 64 per conv tap (lda / 64)
    int kt;           // conv window (3 or 5)
    int ln_pad_one;   // fused LayerNorm writes 1.0 into pad columns ln_C, ln_C + 1
};

inline GemmP make_gemm_params(const srk_gemm_args* g) {
    GemmP p{};
    p.A = (const uint16_t*)g->A; p.a_mode = g->a_mode; p.lda = g->lda;
    p.nB = g->nB; p.H = g->H; p.W = g->W;
    p.Wt = (const uint16_t*)g->Wt; p.M = g->M; p.N = g->N; p.K = g->K; p.dtype = g->dtype;
    p.bias = g->bias; p.act = g->act;
    p.res = g->res; p.res_scale = g->res_scale; p.out32 = g->out32; p.ld32 = g->ld32;
    p.win_shift = g->win_shift;
    p.out16 = (uint16_t*)g->out16; p.ld16 = g->ld16; p.out16_dtype = g->out16_dtype;
    p.out16_mode = g->out16_mode;
    p.ln_g = g->ln_g; p.ln_b = g->ln_b; p.ln_C = g->ln_C; p.ln_win_shift = g->ln_win_shift;
    p.img = g->img; p.img_s = g->img_s; p.img_scale = g->img_scale;
    p.img_hc = g->img_hc; p.img_wc = g->img_wc;
    p.attn_table = g->attn_table; p.attn_heads = g->attn_heads; p.attn_scale = g->attn_scale; p.attn_shift = g->attn_shift;
    p.T = (g->H > 0 && g->W > 0) ? g->H * g->W : 0;
    p.cpb = g->lda / 64;
    p.kt = g->conv_k == 5 ? 5 : 3;
    p.ln_pad_one = g->ln_pad_one;
    return p;
}

// GEMM row m -> row of the fp32 residual / output (window_reverse + roll back when win_shift>=0)
__device__ __forceinline__ int row32_of(const GemmP& p, int m) {
    if (p.win_shift < 0) return m;
    const int bi = m / p.T;
    return bi * p.T + win_pos_to_token(m - bi * p.T, p.H, p.W, p.win_shift);
}

// element offset of (GEMM row m, column n) in the 16-bit output
__device__ __forceinline__ size_t off16_of(const GemmP& p, int m, int n) {
    if (p.out16_mode == SRK_O16_ROWS) return (size_t)m * p.ld16 + n;
    // PixelShuffle(2) fused: packed N order is (i, j, c) -> NHWC pixel (2y+i, 2x+j), channel c
    const int cq = p.N >> 2;
    const int grp = n / cq, c = n - grp * cq;
    const int bi = m / p.T, rem = m - bi * p.T;
    const int y = rem / p.W, x = rem - y * p.W;
    const size_t pix = ((size_t)bi * (2 * p.H) + (2 * y + (grp >> 1))) * (2 * p.W) + (2 * x + (grp & 1));
    return pix * p.ld16 + c;
}

// element offset of (row m, column n = i*s + j) in the cropped 1-channel pixelshuffle-direct
// image; returns false when the element falls outside the crop
__device__ __forceinline__ bool offimg_of(const GemmP& p, int m, int n, size_t& off) {
    const int s = p.img_s;
    const int bi = m / p.T, rem = m - bi * p.T;
    const int y = rem / p.W, x = rem - y * p.W;
    const int i = n / s, j = n - i * s;
    const int oy = y * s + i, ox = x * s + j;
    if (n >= s * s || oy >= p.img_hc || ox >= p.img_wc) return false;
    off = ((size_t)bi * p.img_hc + oy) * p.img_wc + ox;
    return true;
}

}  // namespace srk
