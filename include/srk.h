/*
 * srk.h -- C ABI of libsrk.so: the B200-native (sm_100a) kernels behind the SR-CACO-2
 * evaluation hot path (SwinIR / EDSR forward + PSNR / SSIM / NRMSE scoring).
 *
 * Boundary rules
 *   - plain C: raw DEVICE pointers, sizes, a CUDA stream handle passed as void*;
 *     no torch / C++ types cross this boundary;
 *   - no allocation and no ownership transfer: packed weights, activations, outputs and the
 *     workspace are allocated by the caller (PyTorch) and passed in;
 *   - every entry point returns 0 on success and a negative srk_status otherwise; the message
 *     is available through srk_last_error() (thread local);
 *   - one call in flight per stream; the library keeps no mutable global state besides the
 *     error string and lazily set kernel attributes.
 *
 * Each entry point cites the reference interface it replaces (paths relative to the
 * sbelharbi/sr-caco-2 tree).  The reference has no native boundary on this path (it is pure
 * PyTorch); the binding a maintainer adds is the ctypes stub shown in INTEGRATION.md.
 */
#ifndef SRK_H_
#define SRK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SRK_VERSION 100

typedef enum {
    SRK_OK = 0,
    SRK_ERR_INVALID = -1,   /* bad shape / argument                      */
    SRK_ERR_UNSUPPORTED = -2, /* valid in the reference, not built here  */
    SRK_ERR_CUDA = -3,      /* launch / runtime error (see srk_last_error) */
    SRK_ERR_WORKSPACE = -4, /* workspace too small                       */
    SRK_ERR_ARCH = -5       /* device is not sm_100                      */
} srk_status;

/* 16-bit operand formats of the tensor-core GEMMs */
enum { SRK_BF16 = 0, SRK_FP16 = 1 };
/* epilogue activations */
enum { SRK_ACT_NONE = 0, SRK_ACT_GELU = 1, SRK_ACT_LRELU = 2 /* slope 0.01 */, SRK_ACT_RELU = 3,
       SRK_ACT_LRELU02 = 4 /* slope 0.2 */ };
/* A-operand addressing of srk_gemm */
enum { SRK_A_ROWS = 0, SRK_A_CONV3X3 = 1 };
/* 16-bit output addressing of srk_gemm */
enum { SRK_O16_ROWS = 0, SRK_O16_PIXSHUF2 = 1 };
/* GEMM engines: tcgen05/TMEM/TMA (product path) and the legacy mma.sync kernel that is kept
 * as the in-library cross-check of the former (tests only) */
enum { SRK_ENGINE_TCGEN05 = 0, SRK_ENGINE_MMA_SYNC = 1 };
enum { SRK_UPSAMPLER_PIXELSHUFFLE = 0, SRK_UPSAMPLER_PIXELSHUFFLEDIRECT = 1, SRK_UPSAMPLER_NEAREST_CONV = 2 };
/* `options` of the network plans: 0 is the product launch sequence; each bit switches ONE fusion off so
 * that the parity tests can compare a fused kernel with the unfused sequence it replaces */
enum { SRK_OPT_NO_FUSED_MLP = 1,       /* fc1 + GELU and fc2 + residual + LN as two GEMM launches */
       SRK_OPT_NO_FUSED_ATTN = 2,      /* qkv GEMM, window attention and proj as separate launches */
       SRK_OPT_NO_FOLD_TAIL = 4,       /* run the upsampler convs one by one */
       SRK_OPT_NO_FOLD_QKV_BIAS = 8,   /* add the qkv bias in the epilogue */
       SRK_OPT_NO_FUSED_BLOCK = 16,    /* qkv + attention fused, proj + residual + LN as its own GEMM launch */
       SRK_OPT_NO_GRAPH = 32 };        /* (host layer) do not replay the forward from a CUDA graph */

const char* srk_last_error(void);
int srk_version(void);
/* 0 when `device` is an sm_100 part this library can run on */
int srk_check_device(int device);
/* selects the GEMM engine used by srk_gemm / the network drivers (default TCGEN05) */
int srk_set_engine(int engine);
int srk_get_engine(void);

/* ------------------------------------------------------------------------------------------
 * Index maps (integer, bit exact).  Replaces torch.roll + window_partition
 * (dlib/models/network_swinir.py:296-306, :48-62), window_reverse (:65-80),
 * calculate_mask (:260-285), the relative_position_index buffer (:116-129) and
 * nn.PixelShuffle (:675, :701).  Each fills a device int32 array with the map the kernels
 * apply arithmetically, so tests can compare it with the reference's.
 *   kind 0: out[nW*64]      token id at each window-major position (H, W, shift)
 *   kind 1: out[nW*64*64]   1 where the shifted-window mask is -100, else 0 (H, W, shift)
 *   kind 2: out[64*64]      relative position index (ws = 8)
 *   kind 3: out[C*H*r*W*r]  pixel-shuffle source index of every output element (C=a, r=shift)
 * ---------------------------------------------------------------------------------------- */
int srk_index_map(int kind, int H, int W, int shift, int a, int32_t* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Metrics: one pass over E and H per image.  Replaces, for 1-channel images,
 *   tensor2uint82float              dlib/utils/utils_image.py:369-372   (quantize != 0)
 *   mbatch_gpu_calculate_psnr       :843-891
 *   mbatch_gpu_calculate_mse        :894-934
 *   mbatch_gpu_calculate_nrmse      :937-1007
 *   mbatch_gpu_calculate_ssim       :1120-1198 (+ _ssim_per_channel :1010-1099)
 *   PSNR_Y via mb_gpu_rgb2ycbcr     :618-653, dlib/utils/utils_trainer.py:1005-1012
 *   the ROI variants / marginalize_roi_th_perf  dlib/utils/utils_trainer.py:874-930, :986
 * E, H: (B, 1, Hpx, Wpx) fp32, contiguous.  quantize=1: inputs in [0,1] are first mapped to
 * integer-valued [0,255] (clamp, *255, round-half-even); quantize=0: inputs are used as they
 * are (already in [0,255]).
 * roi_ths[n_ths]: variant v>0 uses roi = (H8 >= roi_ths[v-1]); variant 0 is the full image.
 * out: (B, 1+n_ths, SRK_MET_N) fp64.  flags: (B) int32, bit0 non-finite metric, bit1
 * negative metric, bit2 input outside [0,255] (the reference's sys.exit() guards,
 * dlib/utils/utils_trainer.py:933-958, dlib/utils/utils_image.py:1164-1172).
 * scratch: device buffer of srk_metrics_scratch_bytes(B, n_ths) bytes (zeroed by the call).
 * ---------------------------------------------------------------------------------------- */
enum { SRK_MET_PSNR = 0, SRK_MET_MSE = 1, SRK_MET_NRMSE = 2, SRK_MET_SSIM = 3,
       SRK_MET_PSNR_Y = 4, SRK_MET_N = 5 };
#define SRK_MAX_ROI_THS 8
size_t srk_metrics_scratch_bytes(int B, int n_ths);
/* on != 0: the quantised hot path runs the round-1 tile kernel instead of the row-streaming kernel (tests: the two
 * kernels are cross-checked against each other and against the reference known-answer values) */
int srk_metrics_use_tile_kernel(int on);
/* on == 0: the 64 -> 64 channel 3x3 / 5x5 convs of the tcgen05 engine fetch one TMA box per tap instead of one halo
 * box per tile (default on; tests compare the two operand paths) */
int srk_gemm_conv_halo(int on);
int srk_metrics(const float* E, const float* H, int B, int Hpx, int Wpx, int border,
                int quantize, const int* roi_ths, int n_ths, double* out, int32_t* flags,
                void* scratch, void* stream);
/* Variant with a caller-provided ROI mask (B,1,Hpx,Wpx) fp32 {0,1} (the `roi=` argument of
 * the reference's metric functions); out: (B, SRK_MET_N). */
/* Same as srk_metrics with quantize = 1, the target given as the uint8 levels the loader holds
 * (ship uint8 targets, convert on the device): H8: (B,1,Hpx,Wpx) uint8.  Identical
 * results to srk_metrics(E, H8 / 255.f, ...): tensor2uint82float (utils_image.py:369-372) maps
 * x / 255 back to x.  roi_ths must be ascending. */
int srk_metrics_h8(const float* E, const unsigned char* H8, int B, int Hpx, int Wpx, int border,
                   const int* roi_ths, int n_ths, double* out, int32_t* flags, void* scratch,
                   void* stream);

int srk_metrics_roi(const float* E, const float* H, const float* roi, int B, int Hpx, int Wpx,
                    int border, int quantize, double* out, int32_t* flags, void* scratch,
                    void* stream);

/* ------------------------------------------------------------------------------------------
 * Building blocks (exported for unit tests; the network drivers below call the same code).
 * Activations are token-major: row = (b, y, x), `ld` elements per row (channels padded).
 * ---------------------------------------------------------------------------------------- */

/* Tensor-core GEMM with fused epilogue:  v = act(A (*) Wt^T + bias);  if res: v = v*res_scale
 * + res[r32];  out32[r32] = v;  out16[r16] = cvt(v)
 *   SRK_A_ROWS    : A is (M, lda) 16-bit, K = padded channels
 *   SRK_A_CONV3X3 : A is the (nB, H, W, lda) image, implicit im2col, zero padding 1,
 *                   K = 9 * lda, k = tap * lda + c, tap = ky*3+kx   (nn.Conv2d(.,.,3,1,1));
 *                   conv_k = 5 makes it a 5x5 window (zero padding 2, K = 25 * lda) -- used by
 *                   the folded reconstruction tail (srk_tail_fold)
 *   Wt : (N, K) 16-bit, K contiguous (nn.Linear weight layout; conv weight repacked)
 *   win_shift >= 0: GEMM row m is a window-major position; res / out32 rows are the token it
 *                   came from (window_reverse + roll back), -1: identity
 * Replaces F.linear / F.conv2d calls of network_swinir.py:148,177,39-45,544,850,865,674 and
 * network_nlsn.py:38-41. */
typedef struct {
    const void* A; int a_mode; int lda; int nB, H, W;
    const void* Wt; int M, N, K; int dtype;
    const float* bias; int act;
    const float* res; float res_scale; float* out32; int ld32; int win_shift;
    void* out16; int ld16; int out16_dtype; int out16_mode;
    /* optional fused LayerNorm of the fp32 result row (tcgen05 engine, N tile == whole row):
     * out16 = cvt((v - mean)/sqrt(var+eps) * ln_g + ln_b) over the first ln_C columns; the
     * row is written at the token's row (ln_win_shift = -1) or at its window-major position
     * under cyclic shift ln_win_shift (the NEXT block's roll + window_partition) */
    const float* ln_g; const float* ln_b; int ln_C; int ln_win_shift;
    /* pixelshuffle-direct image output (1 channel), cropped to img_hc x img_wc:
     * img[b, y*s+i, x*s+j] = v[m, i*s+j] * img_scale */
    float* img; int img_s; float img_scale; int img_hc, img_wc;
    /* optional fused window attention (tcgen05 engine): the GEMM is the qkv projection of
     * window-major rows (N = 3*attn_heads*32, rows [q|k|v][head][32]); each CTA keeps a head
     * pair's q|k|v tile in shared memory, runs the 8x8-window attention on it (rel-pos bias
     * attn_table (heads,225), shift mask for cyclic shift attn_shift, softmax, P v) and writes
     * only the attention output to out16 (M, ld16 = attn_heads*32): q, k, v never reach HBM. */
    const float* attn_table; int attn_heads; float attn_scale; int attn_shift;
    int conv_k;                          /* SRK_A_CONV3X3 window: 0 or 3 -> 3x3, 5 -> 5x5 */
    /* fused LayerNorm: also write 1.0 into the two pad columns ln_C, ln_C + 1 of out16 (ln_C even, N >= ln_C + 2):
     * the consumer is a qkv GEMM whose weight holds the bias (hi, lo bf16) in those two K columns */
    int ln_pad_one;
} srk_gemm_args;
int srk_gemm(const srk_gemm_args* g, void* stream);

/* Fused Swin MLP (tcgen05 engine):  v = res + fc2(GELU(fc1(A) + b1)) + b2 ; out32 = v ;
 * out16 = LayerNorm(v; ln_g, ln_b) written at the token's row (ln_win_shift = -1) or at its
 * window-major row for the next block's cyclic shift, or a plain 16-bit cast when ln_g == NULL.
 * A: (M, lda) bf16 rows in token order; W1: (hid_p, Cp) bf16; W2: (Cp, hid_p) bf16; hid_p %% 64 == 0.
 * ln_pad_one: the fused LayerNorm also writes 1.0 into pad columns ln_C, ln_C + 1 (see srk_gemm_args).
 * Replaces Mlp.forward network_swinir.py:39-45 + the residual / norm of :335, :293. */
typedef struct {
    const void* A; int lda; int M; int C; int Cp; int hid_p;
    const void* W1; const float* b1; const void* W2; const float* b2;
    const float* res; float* out32; int ld32;
    void* out16; int ld16; int out16_dtype;
    const float* ln_g; const float* ln_b; int ln_C; int ln_win_shift; int H, W;
    int ln_pad_one;
} srk_mlp_args;
int srk_mlp(const srk_mlp_args* a, void* stream);

/* The attention half of a Swin block in one kernel (tcgen05 engine, padded embedding 192 = 6 heads x 32):
 *   x' = x + proj(window_attention(qkv(A))) + b_proj ;  out32[token] = x' ;  out16[token] = LayerNorm(x'; ln_g, ln_b)
 * A: (M, lda) bf16 = LN1(x) rows in window-major order under cyclic shift `shift`, with 1.0 in the pad columns
 *    C, C + 1 (the qkv bias is folded into those K columns of Wqkv);
 * Wqkv: (6 * 96, Cp) bf16, head-major rows [head][q | k | v][32] (packing.pack_qkv_heads);
 * Wproj: (Cp, 192) bf16 (packing.pack_proj); rel_table: (6, 225) fp32; scale = head_dim^-0.5;
 * res / out32: fp32 residual stream in token order (may alias: every row is read before it is written, by the
 * same warp); out16: LN2 rows in token order.  q, k, v, the attention output and the proj accumulator never reach HBM.
 * Replaces network_swinir.py:287-334 (norm1 output -> attention -> proj -> residual -> norm2) incl. :148-176, :260-285. */
typedef struct {
    const void* A; int lda; int M, C, Cp, H, W, shift, num_heads;
    const void* Wqkv; const void* Wproj; const float* b_proj;
    const float* rel_table; float scale;
    const float* res; float* out32; int ld32;
    void* out16; int ld16; int out16_dtype;
    const float* ln_g; const float* ln_b; int ln_C;
} srk_attn_block_args;
int srk_attn_block(const srk_attn_block_args* a, void* stream);

/* LayerNorm over the first C of ld32 columns of fp32 rows -> 16-bit rows (pad columns zeroed).
 * mode 0: out row m <- LN(x row m); mode 1 (win_shift>=0): out row m (window-major) <-
 * LN(x row token(m)); g == NULL: plain cast without normalisation.  Optionally also writes
 * the fp32 normalised row (x32_out) -- used for patch_embed.norm.
 * Replaces nn.LayerNorm at network_swinir.py:293,335,613,925 fused with roll/window_partition. */
int srk_layernorm(const float* x, int ld32, int M, int C, const float* g, const float* b,
                  float eps, void* out16, int ld16, int out16_dtype, float* x32_out,
                  int H, int W, int win_shift, void* stream);

/* 8x8 window attention on window-major rows.  qkv: (M, ldq) bf16, the first 3*nH*dp columns
 * laid out [q(nH x dp) | k | v]; out: (M, ldo) bf16, head h at columns [h*dp, h*dp+dp).
 * rel_table: (nH, 225) fp32.  scale = head_dim^-0.5 applied to the fp32 scores.
 * Replaces WindowAttention.forward network_swinir.py:150-176 and the mask of :260-285. */
int srk_window_attention(const void* qkv, int ldq, void* out, int ldo, const float* rel_table, int nB,
                         int H, int W, int nH, int dp, float scale, int shift, void* stream);

/* 1-channel 3x3 input conv in exact fp32 (conv_first / EDSR head), x: (B,1,h,w) fp32 with an
 * optional reflect pad to (H,W) (check_image_size, network_swinir.py:908-913) and input
 * scaling (x - mean) * img_range (:934-935, mean = 0 for 1 channel).
 * w: (C, 9) fp32, out32: (B*H*W, ld32) fp32 (may be NULL), out16 (may be NULL). */
int srk_conv_in(const float* x, int B, int h, int w, int H, int W, float in_scale,
                const float* wgt, const float* bias, int C, float* out32, int ld32,
                void* out16, int ld16, int out16_dtype, void* stream);

/* conv_in fused with patch_embed.norm and norm1 of the first Swin block: f0 = conv(x) (fp32),
 * xa = LN(f0; g1,b1) (fp32), a16 = LN(xa; g2,b2) (16-bit) stored at the token's window-major row
 * for cyclic shift win_shift (-1: token order).  ld32 == ld16 = channels padded to 64.
 * ln_pad_one: also write 1.0 into pad columns C, C + 1 of a16 (consumer: a qkv weight with the folded bias).
 * Replaces network_swinir.py:939 + :610-614 + :293-306 in one pass. */
int srk_conv_in_ln(const float* x, int B, int h, int w, int H, int W, float in_scale,
                   const float* wgt, const float* bias, int C, float* f0, float* xa, int ld32,
                   const float* g1, const float* b1, const float* g2, const float* b2,
                   void* a16, int ld16, int out16_dtype, int win_shift, int ln_pad_one, void* stream);

/* 1-channel 3x3 output conv (conv_last / EDSR tail.1): a: (B,H,W,lda) fp16, w: (9, Cin) fp16
 * tap major, y: (B,1,Hc,Wc) fp32 cropped to Hc x Wc, y = (conv + bias) * out_scale.
 * Built for Cin == lda == 64 (num_feat of the reference's upsamplers). */
int srk_conv_out(const void* a, int lda, int B, int H, int W, int Cin, const void* wgt,
                 float bias, float out_scale, float* y, int Hc, int Wc, void* stream);

/* Folded reconstruction tail.  After the last non-linearity the `pixelshuffle` upsampler is
 * log2(s) x [Conv3x3(F -> 4F) + PixelShuffle(2)] followed by Conv3x3(F -> 1) (network_swinir.py:
 * 661-680, 862-868, 960-963; network_nlsn.py Upsampler + tail conv): a LINEAR map from the
 * F-channel h x w feature image to the s*h x s*w image.  Composed offline (packing.fold_tail) it
 * is ONE 5x5 convolution F -> s*s, output n = i*s + j of feature pixel (y, x) being image pixel
 * (s*y + i, s*x + j): 30x fewer FLOPs at s = 8 and none of the 4F-channel intermediates.
 * The zero padding of the intermediate images makes the composed kernel differ on the outermost
 * ring of feature pixels only; variant v = 3*vy + vx (vy, vx: 0 = first row / column,
 * 1 = interior, 2 = last) holds the kernel of those pixels.  All weights and biases are
 * pre-multiplied by w_scale (a power of two that keeps the products in fp16 range). */
typedef struct {
    const void* w;            /* interior kernel (64, 25*64) fp16, row n = i*s + j, k = tap*64 + c; NULL: not folded */
    const float* b;           /* (64) fp32 */
    const void* border_w;     /* (9, 64, 25*64) fp16 variant kernels; inside each 64-channel block k = ks*16 + 2t + 8w + h
                               * is stored at [t][ks][w][h] (mma.sync B-fragment order of tail_border_kernel) */
    const float* border_b;    /* (9, 64) fp32 */
    float w_scale;
} srk_tail_fold;

/* The ring pass of the folded tail: recomputes the outermost ring of feature pixels of
 * a: (B,H,W,64) fp16 with their variant kernels and overwrites those pixels of
 * y: (B,1,Hc,Wc), y = (conv5x5 + bias) * out_scale / w_scale.  H, W >= 3. */
int srk_tail_border(const void* a, int B, int H, int W, int s, const srk_tail_fold* f,
                    float out_scale, float* y, int Hc, int Wc, void* stream);

/* Bicubic baseline: Interpolate.forward (dlib/utils/utils_trainer.py:120-147),
 * y = clamp(F.interpolate(x, scale_factor = s, mode = 'bicubic', antialias = True), 0, 1);
 * x: (B,1,h,w) fp32, y: (B,1,h*s,w*s) fp32. */
int srk_bicubic_upsample(const float* x, int B, int h, int w, int s, float* y, void* stream);

/* ------------------------------------------------------------------------------------------
 * Network drivers.  Replace SwinIR.forward (dlib/models/network_swinir.py:930-970) and the
 * EDSR-baseline forward assembled from dlib/models/network_nlsn.py:72-128,325-369.
 * x: (B, 1, h, w) fp32 in [0,1]; y: (B, 1, h*s, w*s) fp32.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    const void* w; const float* b;      /* packed (Np, 9*Cin_p) 16-bit weights, fp32 bias */
    int cin_p, n_p;
} srk_conv_params;

typedef struct {
    const float *ln1_g, *ln1_b, *ln2_g, *ln2_b;
    const void *w_qkv, *w_proj, *w_fc1, *w_fc2;
    const float *b_qkv, *b_proj, *b_fc1, *b_fc2;
    const float* rel_table;             /* (nH, 225) */
    int shift;                          /* 0 or window_size/2, decided at construction */
    int num_heads;
    /* optional: w_qkv with the bias folded into K columns embed_dim (bf16 hi part) and embed_dim + 1 (bf16 of the
     * remainder); used with a NULL bias when the producer of the A rows wrote 1.0 there (ln_pad_one). NULL: not built */
    const void* w_qkv_fb;
    /* optional: the same folded weight with head-major rows [head][q | k | v][32] for srk_attn_block. NULL: not built */
    const void* w_qkv_hm;
} srk_stb_params;

typedef struct {
    int upscale, in_chans, window_size, embed_dim, hidden_dim, n_layers, upsampler;
    float img_range;
    int Cp;        /* embed_dim padded to a multiple of 64 */
    int hid_p;     /* hidden_dim padded to a multiple of 64 */
    int dp;        /* head_dim padded to a multiple of 16 */
    int ao_p;      /* num_heads*dp padded to a multiple of 64 */
    const int* depths;                  /* host array [n_layers] */
    const srk_stb_params* stbs;         /* host array [sum(depths)] */
    const srk_conv_params* rstb_convs;  /* host array [n_layers] */
    const float *conv_first_w, *conv_first_b;        /* (C,9), (C) fp32 */
    const float *pe_norm_g, *pe_norm_b, *norm_g, *norm_b;
    srk_conv_params conv_after_body;
    srk_conv_params conv_before_upsample;            /* pixelshuffle only */
    srk_conv_params upsample[4];                     /* log2(s) convs (N order (i,j,c)) or the direct conv */
    int n_upsample;
    const void* conv_last_w; float conv_last_b;      /* (9,64) fp16 */
    int linear_dtype, conv_dtype;                    /* SRK_BF16 / SRK_FP16 */
    srk_tail_fold tail_fold;                         /* pixelshuffle only; w == NULL: run the convs one by one */
    /* resi_connection '3conv' (network_swinir.py:545-552, 874-884): the RSTB conv / conv_after_body is
     * Conv3x3(C -> C/4) + LeakyReLU(0.2) + Conv1x1(C/4 -> C/4) + LeakyReLU(0.2) + Conv3x3(C/4 -> C).
     * resi_3conv != 0: rstb_c0[l] / rstb_c1[l] (cab_c0 / cab_c1) hold the first two convs (C/4 padded to
     * 64: n_p = 64; c1 is a (64, 64) row-GEMM weight) and rstb_convs[l] (conv_after_body) the third. */
    int resi_3conv;
    const srk_conv_params* rstb_c0; const srk_conv_params* rstb_c1;
    srk_conv_params cab_c0, cab_c1;
    /* upsampler 'nearest+conv' (X4, :875-886, 948-961): upsample[0..1] hold conv_up1 / conv_up2 composed
     * with the nearest x2 interpolation in front of them (a 3x3 conv on the LOW-res grid, 64 -> 4*64 in
     * PixelShuffle order, packing.pack_conv_nearest2x), conv_hr the 64 -> 64 conv before conv_last. */
    srk_conv_params conv_hr;
    int options;                         /* SRK_OPT_* bits, 0 = default */
} srk_swinir_plan;

size_t srk_swinir_workspace_bytes(const srk_swinir_plan* p, int B, int h, int w);
int srk_swinir_forward(const srk_swinir_plan* p, const float* x, float* y, int B, int h, int w,
                       void* workspace, size_t workspace_bytes, void* stream);

typedef struct {
    int in_chans, n_resblocks, n_feats, scale;
    float res_scale, rgb_range;
    int Fp;                              /* n_feats padded to a multiple of 64 */
    const float *head_w, *head_b;        /* (F,9),(F) fp32 */
    const srk_conv_params* body;         /* host array [2*n_resblocks + 1] */
    srk_conv_params tail_up[4]; int n_tail_up;
    const void* tail_w; float tail_b;    /* (9,F) fp16 */
    int conv_dtype;
    srk_tail_fold tail_fold;             /* w == NULL: run the tail convs one by one */
    int options;                         /* SRK_OPT_* bits, 0 = default */
} srk_edsr_plan;

size_t srk_edsr_workspace_bytes(const srk_edsr_plan* p, int B, int h, int w);
int srk_edsr_forward(const srk_edsr_plan* p, const float* x, float* y, int B, int h, int w,
                     void* workspace, size_t workspace_bytes, void* stream);

/* Optional per-kernel-family CUDA-event timing (bench.py: roofline of the dominant kernel).
 * While enabled, every building-block call records an event pair on its stream.
 * srk_profile_read synchronises those events and returns, per family, the summed device time
 * in ms and the number of calls; reset != 0 clears the records. */
enum { SRK_PROF_GEMM = 0, SRK_PROF_ATTENTION = 1, SRK_PROF_LAYERNORM = 2, SRK_PROF_CONV_IN = 3,
       SRK_PROF_CONV_OUT = 4, SRK_PROF_METRICS = 5,
       SRK_PROF_GEMM_RES_LN = 6,   /* srk_gemm calls with residual + fused LayerNorm on row operands (proj / fc2) */
       SRK_PROF_ATTN_BLOCK = 7,    /* srk_attn_block */
       SRK_PROF_MLP = 8,           /* srk_mlp */
       SRK_PROF_N = 9 };
int srk_profile(int enable);
int srk_profile_read(double* ms_by_family, long long* calls_by_family, int reset);

/* number of kernel launches issued by this library on the calling thread since the last
 * reset (bench.py reports it as gpu_launches) */
long long srk_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif /* SRK_H_ */
