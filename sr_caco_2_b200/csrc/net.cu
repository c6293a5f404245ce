// Network drivers: the launch sequences of SwinIR.forward (dlib/models/network_swinir.py:930-970)
// and of the EDSR-baseline assembled from dlib/models/network_nlsn.py:72-128,325-369, over
// token-major (NHWC) activations with an fp32 residual stream.
#include "common.cuh"
#include <vector>

namespace srk {

struct Bump {
    char* base; size_t off, cap;
    void* take(size_t bytes) {
        off = align_up(off, 256);
        void* p = base ? base + off : nullptr;
        off += bytes;
        return p;
    }
};

static inline int pad8(int v) { return (v + 7) / 8 * 8; }

struct SwinBufs {
    float *F0, *XA, *XB;
    void *A16, *QKV, *AO, *HID, *Y1, *U[5];
    void *T1, *T2;                    // '3conv': the two C/4-channel (padded to 64) intermediates
    int nq_p;
};

static size_t swin_layout(const srk_swinir_plan* p, int B, int H, int W, char* base, SwinBufs* b) {
    Bump a{base, 0, 0};
    const size_t M = (size_t)B * H * W;
    int max_heads = 1;
    int nblk = 0;
    for (int l = 0; l < p->n_layers; ++l) nblk += p->depths[l];
    for (int i = 0; i < nblk; ++i) max_heads = p->stbs[i].num_heads > max_heads ? p->stbs[i].num_heads : max_heads;
    b->nq_p = (int)align_up((size_t)3 * max_heads * p->dp, 64);
    b->F0 = (float*)a.take(M * p->Cp * 4);
    b->XA = (float*)a.take(M * p->Cp * 4);
    b->XB = (float*)a.take(M * p->Cp * 4);
    b->A16 = a.take(M * p->Cp * 2);
    b->QKV = a.take(M * b->nq_p * 2);
    b->AO = a.take(M * p->ao_p * 2);
    b->HID = a.take(M * p->hid_p * 2);
    b->Y1 = a.take(M * p->Cp * 2);
    for (int k = 0; k < 5; ++k) b->U[k] = nullptr;
    if (p->upsampler == SRK_UPSAMPLER_PIXELSHUFFLE) {
        size_t m = M;
        for (int k = 0; k <= p->n_upsample; ++k) { b->U[k] = a.take(m * 64 * 2); m *= 4; }
    } else if (p->upsampler == SRK_UPSAMPLER_NEAREST_CONV) {
        b->U[0] = a.take(M * 64 * 2); b->U[1] = a.take(M * 4 * 64 * 2);
        b->U[2] = a.take(M * 16 * 64 * 2); b->U[3] = a.take(M * 16 * 64 * 2);
    }
    b->T1 = b->T2 = nullptr;
    if (p->resi_3conv) { b->T1 = a.take(M * 64 * 2); b->T2 = a.take(M * 64 * 2); }
    return align_up(a.off, 256);
}

static int check_swin_plan(const srk_swinir_plan* p) {
    SRK_REQUIRE(p && p->depths && p->stbs && p->rstb_convs, "swinir: null plan");
    if (p->in_chans != 1) return fail(SRK_ERR_UNSUPPORTED, "swinir: only in_chans == 1 is built (got %d)", p->in_chans);
    if (p->window_size != 8) return fail(SRK_ERR_UNSUPPORTED, "swinir: only window_size == 8 is built (got %d)", p->window_size);
    if (p->upsampler == SRK_UPSAMPLER_PIXELSHUFFLE) {
        SRK_REQUIRE(p->n_upsample >= 0 && p->n_upsample <= 4 && (1 << p->n_upsample) == p->upscale,
                    "swinir: pixelshuffle needs upscale = 2^n (n <= 4)");
    } else if (p->upsampler == SRK_UPSAMPLER_NEAREST_CONV) {
        SRK_REQUIRE(p->upscale == 4 && p->n_upsample == 2 && p->conv_hr.w && p->conv_last_w, "swinir: nearest+conv is built for X4");
    } else if (p->upsampler == SRK_UPSAMPLER_PIXELSHUFFLEDIRECT) {
        SRK_REQUIRE(p->n_upsample == 1 && p->upscale * p->upscale <= 64, "swinir: direct upsampler needs scale <= 8");
    } else {
        return fail(SRK_ERR_UNSUPPORTED, "swinir: upsampler %d not built", p->upsampler);
    }
    if (p->resi_3conv)
        SRK_REQUIRE(p->rstb_c0 && p->rstb_c1 && p->cab_c0.w && p->cab_c1.w && p->embed_dim / 4 <= 64,
                    "swinir: '3conv' needs the two extra convs per RSTB and embed_dim / 4 <= 64");
    SRK_REQUIRE(p->Cp % 64 == 0 && p->Cp >= p->embed_dim && p->hid_p % 64 == 0 && p->hid_p >= p->hidden_dim &&
                p->dp % 16 == 0 && p->ao_p % 64 == 0 && p->embed_dim % 4 == 0, "swinir: bad padded dims");
    return 0;
}

}  // namespace srk

using namespace srk;

namespace srk {
// the folded tail needs the tcgen05 engine's 5x5 conv; SRK_OPT_NO_FOLD_TAIL runs the upsampler convs one by one
static bool fold_tail_enabled(int options) {
    return !(options & SRK_OPT_NO_FOLD_TAIL) && srk_get_engine() == SRK_ENGINE_TCGEN05;
}
// the folded reconstruction tail: ONE 5x5 conv F -> s*s written as image pixels + the ring pass
int run_folded_tail(const srk_tail_fold& f, const void* feat, int B, int H, int W, int s, float out_scale, float* y,
                    int hc, int wc, void* stream) {
    srk_gemm_args g{};
    g.A = feat; g.a_mode = SRK_A_CONV3X3; g.conv_k = 5; g.lda = 64; g.nB = B; g.H = H; g.W = W;
    g.Wt = f.w; g.M = B * H * W; g.N = 64; g.K = 25 * 64; g.dtype = SRK_FP16;
    g.bias = f.b; g.act = SRK_ACT_NONE; g.res_scale = 1.f; g.win_shift = -1; g.ln_win_shift = -1;
    g.img = y; g.img_s = s; g.img_scale = out_scale / f.w_scale; g.img_hc = hc; g.img_wc = wc;
    if (int rc = srk_gemm(&g, stream)) return rc;
    return srk_tail_border(feat, B, H, W, s, &f, out_scale, y, hc, wc, stream);
}
}  // namespace srk


extern "C" size_t srk_swinir_workspace_bytes(const srk_swinir_plan* p, int B, int h, int w) {
    if (!p || B <= 0 || h <= 0 || w <= 0) return 0;
    SwinBufs b;
    return swin_layout(p, B, pad8(h), pad8(w), nullptr, &b);
}

#define TRY(expr) do { int rc__ = (expr); if (rc__) return rc__; } while (0)

extern "C" int srk_swinir_forward(const srk_swinir_plan* p, const float* x, float* y, int B, int h,
                                  int w, void* workspace, size_t workspace_bytes, void* stream) {
    TRY(check_swin_plan(p));
    SRK_REQUIRE(x && y && workspace, "swinir: null pointer");
    SRK_REQUIRE(B > 0 && h > 0 && w > 0, "swinir: bad input shape");
    const int H = pad8(h), W = pad8(w);
    SRK_REQUIRE(H - h < h && W - w < w, "swinir: reflect padding needs h, w > pad");
    SwinBufs b;
    const size_t need = swin_layout(p, B, H, W, (char*)workspace, &b);
    if (need > workspace_bytes) return fail(SRK_ERR_WORKSPACE, "swinir: workspace %zu < %zu bytes", workspace_bytes, need);
    const int M = B * H * W, C = p->embed_dim, Cp = p->Cp;
    const int ldt = p->linear_dtype, cdt = p->conv_dtype;
    const bool fuse_ln = srk_get_engine() == SRK_ENGINE_TCGEN05 && Cp <= 256;
    const float eps = 1e-5f;

    auto conv_gemm = [&](const void* A, int lda, int Hh, int Ww, const srk_conv_params& cv) {
        srk_gemm_args g{};
        g.A = A; g.a_mode = SRK_A_CONV3X3; g.lda = lda; g.nB = B; g.H = Hh; g.W = Ww;
        g.Wt = cv.w; g.M = B * Hh * Ww; g.N = cv.n_p; g.K = 9 * lda; g.dtype = cdt;
        g.bias = cv.b; g.act = SRK_ACT_NONE; g.res_scale = 1.f; g.win_shift = -1; g.ln_win_shift = -1;
        g.out16_dtype = cdt;
        return g;
    };
    auto lin_gemm = [&](const void* A, int lda, const void* Wt, const float* bias, int N, int K) {
        srk_gemm_args g{};
        g.A = A; g.a_mode = SRK_A_ROWS; g.lda = lda; g.nB = B; g.H = H; g.W = W;
        g.Wt = Wt; g.M = M; g.N = N; g.K = K; g.dtype = ldt; g.bias = bias; g.act = SRK_ACT_NONE;
        g.res_scale = 1.f; g.win_shift = -1; g.ln_win_shift = -1; g.out16_dtype = ldt;
        return g;
    };

    // resi_connection '3conv': Conv3x3(C -> C/4) + LeakyReLU(0.2) + Conv1x1 + LeakyReLU(0.2) in front of the
    // last conv; both intermediates are (M, 64) fp16 (C/4 zero-padded to 64)
    auto three_conv_head = [&](const void* A, const srk_conv_params& c0, const srk_conv_params& c1) -> int {
        srk_gemm_args g0 = conv_gemm(A, Cp, H, W, c0);
        g0.act = SRK_ACT_LRELU02; g0.out16 = b.T1; g0.ld16 = 64;
        if (int rc = srk_gemm(&g0, stream)) return rc;
        srk_gemm_args g1{};
        g1.A = b.T1; g1.a_mode = SRK_A_ROWS; g1.lda = 64; g1.nB = B; g1.H = H; g1.W = W;
        g1.Wt = c1.w; g1.M = M; g1.N = 64; g1.K = 64; g1.dtype = cdt; g1.bias = c1.b; g1.act = SRK_ACT_LRELU02;
        g1.res_scale = 1.f; g1.win_shift = -1; g1.ln_win_shift = -1; g1.out16 = b.T2; g1.ld16 = 64; g1.out16_dtype = cdt;
        return srk_gemm(&g1, stream);
    };

    // conv_first (+ reflect pad + input scaling), patch_embed.norm
    int nblk_total = 0;
    for (int l = 0; l < p->n_layers; ++l) nblk_total += p->depths[l];
    const bool first_fused = nblk_total > 0 && p->depths[0] > 0 && Cp <= 256;
    bool a16_ones_init = false;
    if (first_fused) {
        const srk_stb_params& s0 = p->stbs[0];
        const int ones0 = !(p->options & SRK_OPT_NO_FOLD_QKV_BIAS) && Cp - C >= 2 && C % 2 == 0 && s0.w_qkv_fb != nullptr;
        TRY(srk_conv_in_ln(x, B, h, w, H, W, p->img_range, p->conv_first_w, p->conv_first_b, C, b.F0, b.XA, Cp,
                           p->pe_norm_g, p->pe_norm_b, s0.ln1_g, s0.ln1_b, b.A16, Cp, ldt, s0.shift, ones0, stream));
        a16_ones_init = ones0 != 0;
    } else {
        TRY(srk_conv_in(x, B, h, w, H, W, p->img_range, p->conv_first_w, p->conv_first_b, C, b.F0, Cp,
                        nullptr, 0, 0, stream));
        TRY(srk_layernorm(b.F0, Cp, M, C, p->pe_norm_g, p->pe_norm_b, eps, nullptr, 0, 0, b.XA, H, W, -1, stream));
    }

    int blk = 0;
    void* a16 = b.A16;           // 16-bit operand buffer in use; `alt` is its ping-pong partner (the conv
    void* alt = b.Y1;            // epilogue may not overwrite the image it is convolving)
    bool a16_ready = first_fused; // A16 already holds LN1 of the next block (fused producer)
    // qkv bias folded into the GEMM: the LayerNorm epilogue that produces a block's A rows writes 1.0 into the pad
    // columns C, C + 1 and the block's folded weight (w_qkv_fb) carries the bias there -> no bias add in the epilogue
    const bool fold_ok = !(p->options & SRK_OPT_NO_FOLD_QKV_BIAS) && Cp - C >= 2 && C % 2 == 0;
    bool a16_ones = a16_ones_init;
    bool final_norm_done = false;
    for (int l = 0; l < p->n_layers; ++l) {
        const float* cur = b.XA;
        for (int d = 0; d < p->depths[l]; ++d, ++blk) {
            const srk_stb_params& s = p->stbs[blk];
            const int nH = s.num_heads;
            SRK_REQUIRE(nH * p->dp <= p->ao_p && 3 * nH * p->dp <= b.nq_p, "swinir: head layout overflow");
            SRK_REQUIRE(s.shift == 0 || s.shift == 4, "swinir: shift must be 0 or window_size/2");
            const int hd = C / nH;
            if (!a16_ready) {
                TRY(srk_layernorm(cur, Cp, M, C, s.ln1_g, s.ln1_b, eps, a16, Cp, ldt, nullptr, H, W, s.shift, stream));
                a16_ones = false;
            }
            a16_ready = false;
            const bool fused_attn = !(p->options & SRK_OPT_NO_FUSED_ATTN) && fuse_ln && p->dp == 32 && nH % 2 == 0 && p->ao_p == nH * 32 &&
                                    b.nq_p == 3 * nH * 32;
            // the whole attention half in ONE kernel (attn_block_tc5.cu): q, k, v, the attention output and the proj
            // accumulator stay on the SM; built for 6 heads x 32 with the qkv bias folded into the pad K columns
            const bool fused_block = fused_attn && !(p->options & SRK_OPT_NO_FUSED_BLOCK) && a16_ones && s.w_qkv_hm != nullptr &&
                                     Cp == 192 && nH == 6 && ldt == SRK_BF16;
            if (fused_block) {
                srk_attn_block_args ab{};
                ab.A = a16; ab.lda = Cp; ab.M = M; ab.C = C; ab.Cp = Cp; ab.H = H; ab.W = W; ab.shift = s.shift; ab.num_heads = nH;
                ab.Wqkv = s.w_qkv_hm; ab.Wproj = s.w_proj; ab.b_proj = s.b_proj; ab.rel_table = s.rel_table;
                ab.scale = 1.0f / sqrtf((float)hd);
                ab.res = cur; ab.out32 = b.XB; ab.ld32 = Cp;
                // the kernel reads A tile by tile while other CTAs already write LN2 rows (token order): ping-pong
                ab.out16 = alt; ab.ld16 = Cp; ab.out16_dtype = ldt;
                ab.ln_g = s.ln2_g; ab.ln_b = s.ln2_b; ab.ln_C = C;
                TRY(srk_attn_block(&ab, stream));
                { void* t = a16; a16 = alt; alt = t; }
            } else {
            if (fused_attn) {
                // qkv projection + window attention in one kernel: q, k, v stay in shared memory
                const bool fb = a16_ones && s.w_qkv_fb != nullptr;
                srk_gemm_args g = lin_gemm(a16, Cp, fb ? s.w_qkv_fb : s.w_qkv, fb ? nullptr : s.b_qkv, b.nq_p, Cp);
                g.out16 = b.AO; g.ld16 = p->ao_p;
                g.attn_table = s.rel_table; g.attn_heads = nH; g.attn_scale = 1.0f / sqrtf((float)hd); g.attn_shift = s.shift;
                TRY(srk_gemm(&g, stream));
            } else {
            // qkv
            {
                srk_gemm_args g = lin_gemm(a16, Cp, s.w_qkv, s.b_qkv, b.nq_p, Cp);
                g.out16 = b.QKV; g.ld16 = b.nq_p;
                TRY(srk_gemm(&g, stream));
            }
            TRY(srk_window_attention(b.QKV, b.nq_p, b.AO, p->ao_p, s.rel_table, B, H, W, nH, p->dp,
                                     1.0f / sqrtf((float)hd), s.shift, stream));
            }
            // proj + window_reverse + roll back + residual  [+ LN2 fused]
            {
                srk_gemm_args g = lin_gemm(b.AO, p->ao_p, s.w_proj, s.b_proj, Cp, p->ao_p);
                g.res = cur; g.out32 = b.XB; g.ld32 = Cp; g.win_shift = s.shift;
                if (fuse_ln) {
                    g.ln_g = s.ln2_g; g.ln_b = s.ln2_b; g.ln_C = C; g.ln_win_shift = -1;   // LN2 rows in token order
                    g.out16 = a16; g.ld16 = Cp;
                }
                TRY(srk_gemm(&g, stream));
            }
            }
            if (!fuse_ln)
                TRY(srk_layernorm(b.XB, Cp, M, C, s.ln2_g, s.ln2_b, eps, a16, Cp, ldt, nullptr, H, W, -1, stream));
            const bool last = d == p->depths[l] - 1;
            // fc1 + GELU + fc2 + residual [+ next norm1 | + fp16 cast] in ONE tcgen05 kernel (mlp_tc5.cu): the hidden
            // activation never reaches HBM.  The two-GEMM sequence below remains for the legacy engine / odd shapes.
            const bool fused_mlp = !(p->options & SRK_OPT_NO_FUSED_MLP) && fuse_ln && (Cp == 64 || Cp == 128 || Cp == 192) && p->hid_p % 64 == 0;
            if (fused_mlp) {
                srk_mlp_args m{};
                m.A = a16; m.lda = Cp; m.M = M; m.C = C; m.Cp = Cp; m.hid_p = p->hid_p;
                m.W1 = s.w_fc1; m.b1 = s.b_fc1; m.W2 = s.w_fc2; m.b2 = s.b_fc2;
                m.res = b.XB; m.out32 = b.XB; m.ld32 = Cp; m.out16 = a16; m.ld16 = Cp; m.H = H; m.W = W;
                m.ln_win_shift = -1;
                if (last) {
                    m.out16_dtype = cdt;
                } else {
                    const srk_stb_params& nx = p->stbs[blk + 1];
                    m.ln_g = nx.ln1_g; m.ln_b = nx.ln1_b; m.ln_C = C; m.ln_win_shift = nx.shift; m.out16_dtype = ldt;
                    m.ln_pad_one = fold_ok && nx.w_qkv_fb != nullptr; a16_ones = m.ln_pad_one != 0;
                    a16_ready = true;
                }
                // the kernel reads A (a16) by TMA tile-by-tile and writes out16 rows of OTHER tiles when the
                // next block is shifted, so the 16-bit output must not alias the operand: ping-pong
                m.out16 = alt;
                TRY(srk_mlp(&m, stream));
                { void* t = a16; a16 = alt; alt = t; }
            } else {
            // fc1 + GELU
            {
                srk_gemm_args g = lin_gemm(a16, Cp, s.w_fc1, s.b_fc1, p->hid_p, Cp);
                g.act = SRK_ACT_GELU; g.out16 = b.HID; g.ld16 = p->hid_p;
                TRY(srk_gemm(&g, stream));
            }
            // fc2 + residual  [+ cast for the RSTB conv | + LN1 of the next block fused]
            {
                srk_gemm_args g = lin_gemm(b.HID, p->hid_p, s.w_fc2, s.b_fc2, Cp, p->hid_p);
                g.res = b.XB; g.out32 = b.XB; g.ld32 = Cp;
                if (last) {
                    g.out16 = a16; g.ld16 = Cp; g.out16_dtype = cdt;
                } else if (fuse_ln) {
                    const srk_stb_params& nx = p->stbs[blk + 1];
                    g.ln_g = nx.ln1_g; g.ln_b = nx.ln1_b; g.ln_C = C; g.ln_win_shift = nx.shift;
                    g.out16 = a16; g.ld16 = Cp;
                    g.ln_pad_one = fold_ok && nx.w_qkv_fb != nullptr; a16_ones = g.ln_pad_one != 0;
                    a16_ready = true;
                }
                TRY(srk_gemm(&g, stream));
            }
            }
            cur = b.XB;
        }
        if (p->depths[l] == 0)  // degenerate RSTB: conv of its own input
            TRY(srk_layernorm(b.XA, Cp, M, C, nullptr, nullptr, eps, a16, Cp, cdt, nullptr, H, W, -1, stream));
        // RSTB tail conv + residual with the RSTB input  [+ fused LN: norm1 of the next RSTB's
        // first block (window order), or the final `norm` after the last RSTB (token order)]
        {
            if (p->resi_3conv) TRY(three_conv_head(a16, p->rstb_c0[l], p->rstb_c1[l]));
            srk_gemm_args g = p->resi_3conv ? conv_gemm(b.T2, 64, H, W, p->rstb_convs[l])
                                            : conv_gemm(a16, Cp, H, W, p->rstb_convs[l]);
            g.res = b.XA; g.out32 = b.XA; g.ld32 = Cp;
            bool swap_after = false;
            if (fuse_ln) {
                if (l + 1 < p->n_layers && p->depths[l + 1] > 0) {
                    const srk_stb_params& nx = p->stbs[blk];
                    g.ln_g = nx.ln1_g; g.ln_b = nx.ln1_b; g.ln_C = C; g.ln_win_shift = nx.shift;
                    g.out16 = alt; g.ld16 = Cp; g.out16_dtype = ldt;
                    g.ln_pad_one = fold_ok && nx.w_qkv_fb != nullptr; a16_ones = g.ln_pad_one != 0;
                    a16_ready = true; swap_after = true;
                } else if (l + 1 == p->n_layers) {
                    g.ln_g = p->norm_g; g.ln_b = p->norm_b; g.ln_C = C; g.ln_win_shift = -1;
                    g.out16 = alt; g.ld16 = Cp; g.out16_dtype = cdt;
                    final_norm_done = true; swap_after = true;
                }
            }
            TRY(srk_gemm(&g, stream));
            if (swap_after) { void* t = a16; a16 = alt; alt = t; }
        }
    }
    // final norm -> conv_after_body + shallow residual
    if (!final_norm_done)
        TRY(srk_layernorm(b.XA, Cp, M, C, p->norm_g, p->norm_b, eps, a16, Cp, cdt, nullptr, H, W, -1, stream));
    const void* normed = a16;            // LN(norm) of the body output, fp16 NHWC
    void* y1 = alt;                      // conv_after_body output
    const float out_scale = 1.f / p->img_range;
    const int s_up = p->upscale;
    {
        if (p->resi_3conv) TRY(three_conv_head(normed, p->cab_c0, p->cab_c1));
        srk_gemm_args g = p->resi_3conv ? conv_gemm(b.T2, 64, H, W, p->conv_after_body)
                                        : conv_gemm(normed, Cp, H, W, p->conv_after_body);
        g.res = b.F0; g.ld32 = Cp; g.out16 = y1; g.ld16 = Cp;
        TRY(srk_gemm(&g, stream));
    }
    if (p->upsampler == SRK_UPSAMPLER_PIXELSHUFFLE) {
        {
            srk_gemm_args g = conv_gemm(y1, Cp, H, W, p->conv_before_upsample);
            g.act = SRK_ACT_LRELU; g.out16 = b.U[0]; g.ld16 = 64;
            TRY(srk_gemm(&g, stream));
        }
        if (p->tail_fold.w && H >= 3 && W >= 3 && fold_tail_enabled(p->options)) {
            TRY(run_folded_tail(p->tail_fold, b.U[0], B, H, W, s_up, out_scale, y, h * s_up, w * s_up, stream));
        } else {
            int Hh = H, Ww = W;
            for (int k = 0; k < p->n_upsample; ++k) {
                srk_gemm_args g = conv_gemm(b.U[k], 64, Hh, Ww, p->upsample[k]);
                g.out16 = b.U[k + 1]; g.ld16 = 64; g.out16_mode = SRK_O16_PIXSHUF2;
                TRY(srk_gemm(&g, stream));
                Hh *= 2; Ww *= 2;
            }
            TRY(srk_conv_out(b.U[p->n_upsample], 64, B, Hh, Ww, 64, p->conv_last_w, p->conv_last_b,
                             out_scale, y, h * s_up, w * s_up, stream));
        }
    } else if (p->upsampler == SRK_UPSAMPLER_NEAREST_CONV) {
        // conv_before_upsample + LeakyReLU; 2 x [nearest x2 + conv + LeakyReLU(0.2)] as composed low-res convs
        // with the PixelShuffle epilogue; conv_hr + LeakyReLU(0.2); conv_last   (network_swinir.py:948-961)
        {
            srk_gemm_args g = conv_gemm(y1, Cp, H, W, p->conv_before_upsample);
            g.act = SRK_ACT_LRELU; g.out16 = b.U[0]; g.ld16 = 64;
            TRY(srk_gemm(&g, stream));
        }
        int Hh = H, Ww = W;
        for (int k = 0; k < 2; ++k) {
            srk_gemm_args g = conv_gemm(b.U[k], 64, Hh, Ww, p->upsample[k]);
            g.act = SRK_ACT_LRELU02; g.out16 = b.U[k + 1]; g.ld16 = 64; g.out16_mode = SRK_O16_PIXSHUF2;
            TRY(srk_gemm(&g, stream));
            Hh *= 2; Ww *= 2;
        }
        {
            srk_gemm_args g = conv_gemm(b.U[2], 64, Hh, Ww, p->conv_hr);
            g.act = SRK_ACT_LRELU02; g.out16 = b.U[3]; g.ld16 = 64;
            TRY(srk_gemm(&g, stream));
        }
        TRY(srk_conv_out(b.U[3], 64, B, Hh, Ww, 64, p->conv_last_w, p->conv_last_b, out_scale, y, h * s_up, w * s_up, stream));
    } else {
        srk_gemm_args g = conv_gemm(y1, Cp, H, W, p->upsample[0]);
        g.img = y; g.img_s = s_up; g.img_scale = out_scale; g.img_hc = h * s_up; g.img_wc = w * s_up;
        TRY(srk_gemm(&g, stream));
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
namespace srk {
struct EdsrBufs { float *HF, *R; void *A16, *T16, *U[5]; };
static size_t edsr_layout(const srk_edsr_plan* p, int B, int h, int w, char* base, EdsrBufs* b) {
    Bump a{base, 0, 0};
    const size_t M = (size_t)B * h * w;
    b->HF = (float*)a.take(M * p->Fp * 4);
    b->R = (float*)a.take(M * p->Fp * 4);
    b->A16 = a.take(M * p->Fp * 2);
    b->T16 = a.take(M * p->Fp * 2);
    size_t m = M;
    for (int k = 0; k < 5; ++k) b->U[k] = nullptr;
    for (int k = 0; k <= p->n_tail_up; ++k) { b->U[k] = a.take(m * p->Fp * 2); m *= 4; }
    return align_up(a.off, 256);
}
}  // namespace srk

extern "C" size_t srk_edsr_workspace_bytes(const srk_edsr_plan* p, int B, int h, int w) {
    if (!p || B <= 0 || h <= 0 || w <= 0) return 0;
    EdsrBufs b;
    return edsr_layout(p, B, h, w, nullptr, &b);
}

extern "C" int srk_edsr_forward(const srk_edsr_plan* p, const float* x, float* y, int B, int h, int w,
                                void* workspace, size_t workspace_bytes, void* stream) {
    SRK_REQUIRE(p && p->body && x && y && workspace, "edsr: null pointer");
    if (p->in_chans != 1) return fail(SRK_ERR_UNSUPPORTED, "edsr: only in_chans == 1 is built");
    SRK_REQUIRE(p->n_tail_up >= 0 && p->n_tail_up <= 4 && (1 << p->n_tail_up) == p->scale, "edsr: scale must be 2^n, n <= 4");
    SRK_REQUIRE(p->Fp % 64 == 0 && p->Fp >= p->n_feats && p->n_feats % 4 == 0 && p->Fp == p->n_feats,
                "edsr: n_feats must be a multiple of 64");
    SRK_REQUIRE(B > 0 && h > 0 && w > 0, "edsr: bad input shape");
    EdsrBufs b;
    const size_t need = edsr_layout(p, B, h, w, (char*)workspace, &b);
    if (need > workspace_bytes) return fail(SRK_ERR_WORKSPACE, "edsr: workspace %zu < %zu bytes", workspace_bytes, need);
    const int M = B * h * w, Fp = p->Fp, cdt = p->conv_dtype;
    auto conv_gemm = [&](const void* A, int Hh, int Ww, const srk_conv_params& cv) {
        srk_gemm_args g{};
        g.A = A; g.a_mode = SRK_A_CONV3X3; g.lda = Fp; g.nB = B; g.H = Hh; g.W = Ww;
        g.Wt = cv.w; g.M = B * Hh * Ww; g.N = cv.n_p; g.K = 9 * Fp; g.dtype = cdt;
        g.bias = cv.b; g.act = SRK_ACT_NONE; g.res_scale = 1.f; g.win_shift = -1; g.ln_win_shift = -1;
        g.out16_dtype = cdt;
        return g;
    };
    (void)M;
    TRY(srk_conv_in(x, B, h, w, h, w, 1.f, p->head_w, p->head_b, p->n_feats, b.HF, Fp, b.A16, Fp, cdt, stream));
    const float* cur = b.HF;
    for (int i = 0; i < p->n_resblocks; ++i) {
        {
            srk_gemm_args g = conv_gemm(b.A16, h, w, p->body[2 * i]);
            g.act = SRK_ACT_RELU; g.out16 = b.T16; g.ld16 = Fp;
            TRY(srk_gemm(&g, stream));
        }
        {
            srk_gemm_args g = conv_gemm(b.T16, h, w, p->body[2 * i + 1]);
            g.res = cur; g.res_scale = p->res_scale; g.out32 = b.R; g.ld32 = Fp;
            g.out16 = b.A16; g.ld16 = Fp;
            TRY(srk_gemm(&g, stream));
        }
        cur = b.R;
    }
    {
        srk_gemm_args g = conv_gemm(b.A16, h, w, p->body[2 * p->n_resblocks]);
        g.res = b.HF; g.ld32 = Fp; g.out16 = b.U[0]; g.ld16 = Fp;
        TRY(srk_gemm(&g, stream));
    }
    if (p->tail_fold.w && Fp == 64 && h >= 3 && w >= 3 && fold_tail_enabled(p->options))
        return run_folded_tail(p->tail_fold, b.U[0], B, h, w, p->scale, 1.f, y, h * p->scale, w * p->scale, stream);
    int Hh = h, Ww = w;
    for (int k = 0; k < p->n_tail_up; ++k) {
        srk_gemm_args g = conv_gemm(b.U[k], Hh, Ww, p->tail_up[k]);
        g.out16 = b.U[k + 1]; g.ld16 = Fp; g.out16_mode = SRK_O16_PIXSHUF2;
        TRY(srk_gemm(&g, stream));
        Hh *= 2; Ww *= 2;
    }
    TRY(srk_conv_out(b.U[p->n_tail_up], Fp, B, Hh, Ww, p->n_feats, p->tail_w, p->tail_b, 1.f, y,
                     h * p->scale, w * p->scale, stream));
    return 0;
}
