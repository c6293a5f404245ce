// HBM-bound kernels of the path: LayerNorm (+ cyclic shift / window partition gather on the
// store side), the 1-channel input conv (exact fp32) and the 1-channel output conv.
#include "common.cuh"

namespace srk {

// ------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, row held in registers (two-pass mean / variance like
// torch.nn.LayerNorm, eps inside the sqrt), 16 B loads, 8 B stores.
// Restates nn.LayerNorm at network_swinir.py:293 (norm1, fused with the roll(-s) +
// window_partition of :296-306), :335 (norm2), :613 (patch_embed.norm), :925 (norm).
// ------------------------------------------------------------------------------------------
constexpr int LN_MAXV = 4;   // float4 per lane: C <= 512

__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, int ld32, int M, int C, const float* __restrict__ g,
                 const float* __restrict__ b, float eps, uint16_t* __restrict__ out16, int ld16,
                 int dt16, float* __restrict__ x32_out, int T, int H, int W, int win_shift) {
    const int warps_per_block = blockDim.x >> 5;
    const int m = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (m >= M) return;
    int src = m;
    if (win_shift >= 0) {
        const int bi = m / T;
        src = bi * T + win_pos_to_token(m - bi * T, H, W, win_shift);
    }
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)src * ld32);
    const int nv = C >> 2;
    float4 v[LN_MAXV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
        const int c = lane + i * 32;
        v[i] = c < nv ? __ldg(xr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    float mean = 0.f, rstd = 1.f;
    if (g != nullptr) {
        mean = warp_sum(s) / (float)C;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < LN_MAXV; ++i) {
            if (lane + i * 32 < nv) {
                const float a = v[i].x - mean, bb = v[i].y - mean, c2 = v[i].z - mean, d = v[i].w - mean;
                q += (a * a + bb * bb) + (c2 * c2 + d * d);
            }
        }
        rstd = 1.f / sqrtf(warp_sum(q) / (float)C + eps);
    }
    const float4* g4 = reinterpret_cast<const float4*>(g);
    const float4* b4 = reinterpret_cast<const float4*>(b);
    uint2* o16 = out16 ? reinterpret_cast<uint2*>(out16 + (size_t)m * ld16) : nullptr;
    float4* o32 = x32_out ? reinterpret_cast<float4*>(x32_out + (size_t)src * ld32) : nullptr;
    const int nv16 = ld16 >> 2, nv32 = ld32 >> 2;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
        const int c = lane + i * 32;
        float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < nv) {
            y = v[i];
            if (g != nullptr) {
                const float4 gg = __ldg(g4 + c), bb = __ldg(b4 + c);
                y.x = (y.x - mean) * rstd * gg.x + bb.x;
                y.y = (y.y - mean) * rstd * gg.y + bb.y;
                y.z = (y.z - mean) * rstd * gg.z + bb.z;
                y.w = (y.w - mean) * rstd * gg.w + bb.w;
            }
        }
        if (o16 && c < nv16) o16[c] = make_uint2(pack2(y.x, y.y, dt16), pack2(y.z, y.w, dt16));
        if (o32 && c < nv32) o32[c] = y;
    }
}

// ------------------------------------------------------------------------------------------
// conv_in: 3x3, 1 -> C channels, exact fp32 (conv_first network_swinir.py:786,939; EDSR head
// network_nlsn.py:325), with the reflect pad of check_image_size (:908-913) and the
// (x - mean) * img_range of :934-935 folded into the load.  One thread = one pixel x 4 channels.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
conv_in_kernel(const float* __restrict__ x, int B, int h, int w, int H, int W, float in_scale,
               const float* __restrict__ wgt, const float* __restrict__ bias, int C,
               float* __restrict__ out32, int ld32, uint16_t* __restrict__ out16, int ld16,
               int dt16) {
    extern __shared__ float cw[];                 // [C][9] then [C] bias
    for (int i = threadIdx.x; i < C * 9; i += blockDim.x) cw[i] = wgt[i];
    for (int i = threadIdx.x; i < C; i += blockDim.x) cw[C * 9 + i] = bias[i];
    __syncthreads();
    const int ldmax = max(out32 ? ld32 : 0, out16 ? ld16 : 0);
    const int groups = ldmax >> 2;                // 4-channel groups incl. zero pad columns
    const long long total = (long long)B * H * W * groups;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int cg = (int)(idx % groups);
        const long long pix = idx / groups;
        const int px = (int)(pix % W), py = (int)((pix / W) % H), bi = (int)(pix / ((long long)W * H));
        float in[9];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                int yy = py + dy - 1, xx = px + dx - 1;
                float val = 0.f;
                if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
                    if (yy >= h) yy = 2 * (h - 1) - yy;      // F.pad(..., 'reflect') bottom / right
                    if (xx >= w) xx = 2 * (w - 1) - xx;
                    val = __ldg(x + ((size_t)bi * h + yy) * w + xx) * in_scale;
                }
                in[dy * 3 + dx] = val;
            }
        }
        float r[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c = cg * 4 + k;
            float acc = 0.f;
            if (c < C) {
                acc = cw[C * 9 + c];
#pragma unroll
                for (int tp = 0; tp < 9; ++tp) acc = fmaf(cw[c * 9 + tp], in[tp], acc);
            }
            r[k] = acc;
        }
        if (out32 && cg * 4 < ld32)
            *reinterpret_cast<float4*>(out32 + (size_t)pix * ld32 + cg * 4) =
                make_float4(r[0], r[1], r[2], r[3]);
        if (out16 && cg * 4 < ld16)
            *reinterpret_cast<uint2*>(out16 + (size_t)pix * ld16 + cg * 4) =
                make_uint2(pack2(r[0], r[1], dt16), pack2(r[2], r[3], dt16));
    }
}


// ------------------------------------------------------------------------------------------
// conv_in + patch_embed.norm + norm1 of the first Swin block in one pass (one warp per pixel,
// lane = channel pair): writes the shallow feature F0 (fp32), the residual stream X = LN_pe(F0)
// (fp32) and the first block's GEMM operand LN_1(X) (16-bit, window-major rows).
// network_swinir.py:939 (conv_first), :610-614 (patch_embed + norm), :293-306 (norm1 + partition).
// ------------------------------------------------------------------------------------------
template <int NP>
__global__ void __launch_bounds__(256)
conv_in_ln_kernel(const float* __restrict__ x, int B, int h, int w, int H, int W, float in_scale,
                  const float* __restrict__ wgt, const float* __restrict__ bias, int C,
                  float* __restrict__ f0, float* __restrict__ xa, int ld32,
                  const float* __restrict__ g1, const float* __restrict__ b1,
                  const float* __restrict__ g2, const float* __restrict__ b2,
                  uint16_t* __restrict__ a16, int ld16, int dt16, int win_shift, int pad_one) {
    extern __shared__ float cws[];                // [9][ld32] weights (tap major, zero padded), then bias[ld32]
    for (int i = threadIdx.x; i < 9 * ld32; i += blockDim.x) {
        const int t = i / ld32, c = i - t * ld32;
        cws[i] = c < C ? wgt[c * 9 + t] : 0.f;
    }
    for (int i = threadIdx.x; i < ld32; i += blockDim.x) cws[9 * ld32 + i] = i < C ? bias[i] : 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int npix = B * H * W;                                   // host checks B*H*W < 2^31: 32-bit index math
    const int wstride = gridDim.x * (blockDim.x >> 5);
    const int T = H * W;
    float2 G1[NP], B1[NP], G2[NP], B2[NP];
#pragma unroll
    for (int k = 0; k < NP; ++k) {
        const int c = 64 * k + 2 * lane;
        const bool in = c < C;
        G1[k] = in ? *reinterpret_cast<const float2*>(g1 + c) : make_float2(0.f, 0.f);
        B1[k] = in ? *reinterpret_cast<const float2*>(b1 + c) : make_float2(0.f, 0.f);
        G2[k] = in ? *reinterpret_cast<const float2*>(g2 + c) : make_float2(0.f, 0.f);
        B2[k] = in ? *reinterpret_cast<const float2*>(b2 + c) : make_float2(0.f, 0.f);
    }
    const float inv_c = 1.f / (float)C;
    for (int pix = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); pix < npix; pix += wstride) {
        const int bi = pix / T, tok = pix - bi * T;
        const int py = tok / W, px = tok - py * W;
        float in[9];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                int yy = py + dy - 1, xx = px + dx - 1;
                float val = 0.f;
                if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
                    if (yy >= h) yy = 2 * (h - 1) - yy;
                    if (xx >= w) xx = 2 * (w - 1) - xx;
                    val = __ldg(x + ((size_t)bi * h + yy) * w + xx) * in_scale;
                }
                in[dy * 3 + dx] = val;
            }
        float2 v[NP];
        float sm = 0.f;
#pragma unroll
        for (int k = 0; k < NP; ++k) {
            const int c = 64 * k + 2 * lane;
            float2 acc = *reinterpret_cast<const float2*>(cws + 9 * ld32 + c);
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const float2 ww = *reinterpret_cast<const float2*>(cws + t * ld32 + c);
                acc.x = fmaf(ww.x, in[t], acc.x); acc.y = fmaf(ww.y, in[t], acc.y);
            }
            v[k] = acc;
            sm += acc.x + acc.y;
            *reinterpret_cast<float2*>(f0 + (size_t)pix * ld32 + c) = acc;
        }
        // patch_embed.norm
        float mean = warp_sum(sm) * inv_c, q = 0.f;
#pragma unroll
        for (int k = 0; k < NP; ++k)
            if (64 * k + 2 * lane < C) { const float a0 = v[k].x - mean, a1 = v[k].y - mean; q += a0 * a0 + a1 * a1; }
        float rstd = 1.f / sqrtf(warp_sum(q) * inv_c + 1e-5f);
        sm = 0.f;
#pragma unroll
        for (int k = 0; k < NP; ++k) {
            const bool inb = 64 * k + 2 * lane < C;
            v[k].x = inb ? (v[k].x - mean) * rstd * G1[k].x + B1[k].x : 0.f;
            v[k].y = inb ? (v[k].y - mean) * rstd * G1[k].y + B1[k].y : 0.f;
            sm += v[k].x + v[k].y;
            *reinterpret_cast<float2*>(xa + (size_t)pix * ld32 + 64 * k + 2 * lane) = v[k];
        }
        // norm1 of the first block, stored at the token's window-major row
        mean = warp_sum(sm) * inv_c; q = 0.f;
#pragma unroll
        for (int k = 0; k < NP; ++k)
            if (64 * k + 2 * lane < C) { const float a0 = v[k].x - mean, a1 = v[k].y - mean; q += a0 * a0 + a1 * a1; }
        rstd = 1.f / sqrtf(warp_sum(q) * inv_c + 1e-5f);
        const size_t row16 = (size_t)bi * T + (win_shift >= 0 ? token_to_win_pos(tok, H, W, win_shift) : tok);
#pragma unroll
        for (int k = 0; k < NP; ++k) {
            const bool inb = 64 * k + 2 * lane < C;
            const float padv = (pad_one && 64 * k + 2 * lane == C) ? 1.f : 0.f;      // pad columns C, C + 1 carry the folded qkv bias
            const float y0 = inb ? (v[k].x - mean) * rstd * G2[k].x + B2[k].x : padv;
            const float y1 = inb ? (v[k].y - mean) * rstd * G2[k].y + B2[k].y : padv;
            *reinterpret_cast<uint32_t*>(a16 + row16 * ld16 + 64 * k + 2 * lane) = pack2(y0, y1, dt16);
        }
    }
}

// ------------------------------------------------------------------------------------------
// conv_out: 3x3, 64 -> 1 channel on an NHWC fp16 image (conv_last network_swinir.py:868,942;
// EDSR tail.1 network_nlsn.py:352-356).  Tensor-core direct convolution from a shared-memory
// halo tile: a CTA owns 8 x 32 output pixels, stages the 10 x 34 x 64ch halo once (cp.async,
// zero-filled outside the image = the conv padding, double buffered across tiles) and each
// warp produces one output row.  The MMA N dimension carries the three horizontal taps:
//   D[p][dx] = sum_dy sum_c halo[row+dy][p][c] * w[dy][dx][c]        (mma.sync m16n8k16, fp16)
//   out[x]   = D[x][0] + D[x+1][1] + D[x+2][2] + bias
// so every A fragment (ldmatrix) is used for 3 taps.  Weights are fp16 here (fp32 accumulate).
// ------------------------------------------------------------------------------------------
constexpr int CO_TH = 8, CO_TW = 32, CO_HH = CO_TH + 2, CO_HW = CO_TW + 2;
constexpr int CO_PIX_STRIDE = 144;                               // 128 B of channels + 16 B pad (ldmatrix conflict-free)
constexpr int CO_TILE_BYTES = CO_HH * CO_HW * CO_PIX_STRIDE;     // 48,960 B
constexpr int CO_THREADS = 256;

__device__ __forceinline__ void co_cp16(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(sz));
}

__global__ void __launch_bounds__(CO_THREADS, 2)
conv_out_mma_kernel(const __half* __restrict__ a, int B, int H, int W, const __half* __restrict__ wgt16,
                    float bias, float out_scale, float* __restrict__ y, int Hc, int Wc) {
    extern __shared__ __align__(16) unsigned char co_smem[];
    const uint32_t sm = (uint32_t)__cvta_generic_to_shared(co_smem);
    float* dsm = reinterpret_cast<float*>(co_smem + 2 * CO_TILE_BYTES);          // [8 warps][3][48]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int tx_n = (Wc + CO_TW - 1) / CO_TW, ty_n = (Hc + CO_TH - 1) / CO_TH;
    const int n_tiles = B * ty_n * tx_n;

    // B fragments: B[k = channel][n = dx] = w[dy][dx][channel] for n < 3 (wgt16 is [9][64], tap major)
    uint32_t bf[3][4][2];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            uint32_t v0 = 0, v1 = 0;
            if (g < 3) {
                const __half* wp = wgt16 + (dy * 3 + g) * 64 + ks * 16 + 2 * t;
                v0 = *reinterpret_cast<const uint32_t*>(wp);
                v1 = *reinterpret_cast<const uint32_t*>(wp + 8);
            }
            bf[dy][ks][0] = v0; bf[dy][ks][1] = v1;
        }

    auto load_tile = [&](int tile, int buf) {
        const int bi = tile / (ty_n * tx_n), rem = tile - bi * ty_n * tx_n;
        const int y0 = (rem / tx_n) * CO_TH - 1, x0 = (rem % tx_n) * CO_TW - 1;
        const uint32_t dst0 = sm + buf * CO_TILE_BYTES;
        for (int i = tid; i < CO_HH * CO_HW * 8; i += CO_THREADS) {
            const int pix = i >> 3, ch = i & 7;
            const int py = pix / CO_HW, px = pix - py * CO_HW;
            const int yy = y0 + py, xx = x0 + px;
            const bool ok = yy >= 0 && yy < H && xx >= 0 && xx < W;
            const __half* src = ok ? a + (((size_t)bi * H + yy) * W + xx) * 64 + ch * 8 : a;
            co_cp16(dst0 + pix * CO_PIX_STRIDE + ch * 16, src, ok);
        }
        asm volatile("cp.async.commit_group;\n");
    };

    int buf = 0;
    if ((int)blockIdx.x < n_tiles) load_tile(blockIdx.x, 0);
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, buf ^= 1) {
        const int nxt = tile + gridDim.x;
        if (nxt < n_tiles) { load_tile(nxt, buf ^ 1); asm volatile("cp.async.wait_group 1;\n"); }
        else asm volatile("cp.async.wait_group 0;\n");
        __syncthreads();
        const uint32_t tb = sm + buf * CO_TILE_BYTES;
        float acc[3][4];
#pragma unroll
        for (int mt = 0; mt < 3; ++mt) acc[mt][0] = acc[mt][1] = acc[mt][2] = acc[mt][3] = 0.f;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
            for (int mt = 0; mt < 3; ++mt) {
                // halo pixels mt*16 .. mt*16+15 of halo row (warp + dy); the third tile is clamped
                int prow = mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                prow = prow < CO_HW ? prow : CO_HW - 1;
                const uint32_t abase = tb + ((warp + dy) * CO_HW + prow) * CO_PIX_STRIDE + (lane >> 4) * 16;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    uint32_t af[4];
                    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                                 : "=r"(af[0]), "=r"(af[1]), "=r"(af[2]), "=r"(af[3]) : "r"(abase + ks * 32));
                    asm volatile(
                        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                        : "+f"(acc[mt][0]), "+f"(acc[mt][1]), "+f"(acc[mt][2]), "+f"(acc[mt][3])
                        : "r"(af[0]), "r"(af[1]), "r"(af[2]), "r"(af[3]), "r"(bf[dy][ks][0]), "r"(bf[dy][ks][1]));
                }
            }
        }
        // D columns 0..2 of the 48 halo pixels -> per-warp scratch, then out[x] = D[x][0] + D[x+1][1] + D[x+2][2]
        float* dw = dsm + warp * 3 * 48;
#pragma unroll
        for (int mt = 0; mt < 3; ++mt) {
            if (t == 0) {
                dw[0 * 48 + mt * 16 + g] = acc[mt][0]; dw[1 * 48 + mt * 16 + g] = acc[mt][1];
                dw[0 * 48 + mt * 16 + g + 8] = acc[mt][2]; dw[1 * 48 + mt * 16 + g + 8] = acc[mt][3];
            } else if (t == 1) {
                dw[2 * 48 + mt * 16 + g] = acc[mt][0]; dw[2 * 48 + mt * 16 + g + 8] = acc[mt][2];
            }
        }
        __syncwarp();
        {
            const int bi = tile / (ty_n * tx_n), rem = tile - bi * ty_n * tx_n;
            const int oy = (rem / tx_n) * CO_TH + warp, ox = (rem % tx_n) * CO_TW + lane;
            const float v = dw[lane] + dw[48 + lane + 1] + dw[96 + lane + 2];
            if (oy < Hc && ox < Wc) y[((size_t)bi * Hc + oy) * Wc + ox] = (v + bias) * out_scale;
        }
        __syncthreads();                                             // tile buffer `buf` is free again
    }
}

// ---- ring pass of the folded reconstruction tail (srk_tail_fold) --------------------------------
// The composed 5x5 kernel of the outermost ring of feature pixels differs from the interior one
// (zero padding of the intermediate images).  One CTA takes up to 16 ring pixels that share a
// variant: a stretch of an edge, or one corner.  Their 5x5 x 64-channel neighbourhoods are staged in
// shared memory as a 16 x 1600 fp16 matrix and multiplied with the variant kernel (64 x 1600, read
// as mma.sync B fragments straight from global / L2) on the tensor cores.
constexpr int TB_K = 25 * 64;
constexpr int TB_STRIDE = (TB_K + 8) * 2;            // bytes per staged pixel row (16 B skew: conflict-free ldmatrix)
constexpr int TB_THREADS = 256;                      // 8 warps, one 8-output n-tile each

// border_w layout (packing.fold_tail): [variant][n][tap][t][ks][w][2] fp16 -- inside every 64-channel
// block the K order is permuted so that the B fragments of the 4 k-steps of a tap,
//   b[ks][w] = W[n][tap*64 + ks*16 + 2t + 8w + {0,1}],
// are 32 contiguous bytes for lane (g = n % 8, t): two 16 B loads per tap.
__global__ void __launch_bounds__(TB_THREADS)
tail_border_kernel(const __half* __restrict__ a, int H, int W, int s, const __half* __restrict__ bw,
                   const float* __restrict__ bb, float out_scale, float* __restrict__ y, int Hc, int Wc) {
    extern __shared__ __align__(16) unsigned char tb_smem[];
    const uint32_t sm = (uint32_t)__cvta_generic_to_shared(tb_smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int bi = blockIdx.y;
    // decode the CTA's stretch: segments 0..3 = top / bottom / left / right edge without corners, 4..7 = corners
    const int cw = (W - 2 + 15) / 16, ch = (H - 2 + 15) / 16;
    int c = blockIdx.x, seg, first, count;
    if (c < 2 * cw) { seg = c / cw; first = 1 + (c % cw) * 16; count = min(16, W - 1 - first); }
    else if (c < 2 * cw + 2 * ch) { c -= 2 * cw; seg = 2 + c / ch; first = 1 + (c % ch) * 16; count = min(16, H - 1 - first); }
    else { seg = 4 + (c - 2 * cw - 2 * ch); first = 0; count = 1; }
    auto pixel = [&](int i, int& py, int& px) {
        switch (seg) {
            case 0: py = 0; px = first + i; break;
            case 1: py = H - 1; px = first + i; break;
            case 2: py = first + i; px = 0; break;
            case 3: py = first + i; px = W - 1; break;
            default: py = (seg & 2) ? H - 1 : 0; px = (seg & 1) ? W - 1 : 0; break;
        }
    };
    int vy, vx;
    switch (seg) {
        case 0: vy = 0; vx = 1; break;
        case 1: vy = 2; vx = 1; break;
        case 2: vy = 1; vx = 0; break;
        case 3: vy = 1; vx = 2; break;
        default: vy = (seg & 2) ? 2 : 0; vx = (seg & 1) ? 2 : 0; break;
    }
    const int variant = vy * 3 + vx;
    // stage the neighbourhoods: row i = pixel i of the stretch, column k = tap*64 + c
    for (int idx = tid; idx < 16 * 25 * 8; idx += TB_THREADS) {
        const int i = idx / 200, rem = idx - i * 200, tap = rem >> 3, chunk = rem & 7;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (i < count) {
            int py, px;
            pixel(i, py, px);
            const int yy = py + tap / 5 - 2, xx = px + tap % 5 - 2;
            if (yy >= 0 && yy < H && xx >= 0 && xx < W)
                v = *reinterpret_cast<const uint4*>(a + (((size_t)bi * H + yy) * W + xx) * 64 + chunk * 8);
        }
        *reinterpret_cast<uint4*>(tb_smem + i * TB_STRIDE + tap * 128 + chunk * 16) = v;
    }
    __syncthreads();
    const int n_out = s * s;
    if (warp * 8 >= n_out) return;                               // this warp's 8 outputs do not exist for this scale
    const int g = lane >> 2, t = lane & 3;
    const uint4* wp = reinterpret_cast<const uint4*>(bw + ((size_t)variant * 64 + warp * 8 + g) * TB_K) + t * 2;
    const uint32_t a_addr = sm + ((lane & 7) + ((lane >> 3) & 1) * 8) * TB_STRIDE + (lane >> 4) * 16;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 5
    for (int tap = 0; tap < 25; ++tap) {
        const uint4 w0 = __ldg(wp + tap * 8), w1 = __ldg(wp + tap * 8 + 1);   // b[ks][w] for ks = 0,1 | 2,3
        const uint32_t bf[4][2] = {{w0.x, w0.y}, {w0.z, w0.w}, {w1.x, w1.y}, {w1.z, w1.w}};
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            uint32_t af[4];
            asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                         : "=r"(af[0]), "=r"(af[1]), "=r"(af[2]), "=r"(af[3]) : "r"(a_addr + tap * 128 + ks * 32));
            asm volatile(
                "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                : "+f"(acc[0]), "+f"(acc[1]), "+f"(acc[2]), "+f"(acc[3])
                : "r"(af[0]), "r"(af[1]), "r"(af[2]), "r"(af[3]), "r"(bf[ks][0]), "r"(bf[ks][1]));
        }
    }
    // accumulator (row g / g+8, columns 2t, 2t+1 of the warp's n-tile) -> image pixel (s*py + n/s, s*px + n%s)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int i = g + (e >> 1) * 8, n = warp * 8 + 2 * t + (e & 1);
        if (i < count && n < n_out) {
            int py, px;
            pixel(i, py, px);
            const int oy = py * s + n / s, ox = px * s + n % s;
            if (oy < Hc && ox < Wc)
                y[((size_t)bi * Hc + oy) * Wc + ox] = (acc[e] + bb[variant * 64 + n]) * out_scale;
        }
    }
}

// ---- bicubic baseline ---------------------------------------------------------------
// Interpolate.forward (dlib/utils/utils_trainer.py:120-147): F.interpolate(scale_factor = s, mode =
// 'bicubic', antialias = True) then clamp to [0, 1].  The anti-aliased bicubic is the separable PIL
// filter with a = -0.5; for an up-scaling its support is 2 input pixels each side, and at the image
// border the taps that do not exist are dropped and the rest renormalised (no edge replication).
__device__ __forceinline__ float cubic_aa(float x) {
    x = fabsf(x);
    const float a = -0.5f;
    if (x < 1.f) return ((a + 2.f) * x - (a + 3.f)) * x * x + 1.f;
    if (x < 2.f) return (((x - 5.f) * x + 8.f) * x - 4.f) * a;
    return 0.f;
}
// taps of output index o along an axis of n input pixels: first tap, count (<= 4), normalised weights
__device__ __forceinline__ void bicubic_taps(int o, int n, float inv_s, int& lo, int& cnt, float (&wt)[4]) {
    const float center = inv_s * ((float)o + 0.5f);
    lo = max(0, (int)(center - 2.f + 0.5f));
    const int hi = min(n, (int)(center + 2.f + 0.5f));
    cnt = hi - lo;
    float tot = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        wt[j] = j < cnt ? cubic_aa((float)(lo + j) - center + 0.5f) : 0.f;
        tot += wt[j];
    }
    const float inv = 1.f / tot;
#pragma unroll
    for (int j = 0; j < 4; ++j) wt[j] *= inv;
}

__global__ void __launch_bounds__(256)
bicubic_up_kernel(const float* __restrict__ x, int h, int w, int s, float* __restrict__ y) {
    const int W = w * s, Hh = h * s;
    const int ox = blockIdx.x * 32 + (threadIdx.x & 31), oy = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (ox >= W || oy >= Hh) return;
    const float inv_s = 1.f / (float)s;
    int x0, nx, y0, ny;
    float wx[4], wy[4];
    bicubic_taps(ox, w, inv_s, x0, nx, wx);
    bicubic_taps(oy, h, inv_s, y0, ny, wy);
    const float* xb = x + (size_t)blockIdx.z * h * w;
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (j < ny) {
            const float* row = xb + (size_t)(y0 + j) * w + x0;
            float r = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) if (i < nx) r = fmaf(wx[i], __ldg(row + i), r);
            acc = fmaf(wy[j], r, acc);
        }
    }
    y[((size_t)blockIdx.z * Hh + oy) * W + ox] = fminf(fmaxf(acc, 0.f), 1.f);
}

}  // namespace srk

using namespace srk;

extern "C" int srk_layernorm(const float* x, int ld32, int M, int C, const float* g, const float* b,
                             float eps, void* out16, int ld16, int out16_dtype, float* x32_out,
                             int H, int W, int win_shift, void* stream) {
    SRK_REQUIRE(x && (out16 || x32_out), "layernorm: null pointer");
    SRK_REQUIRE(M > 0 && C > 0 && C % 4 == 0 && C <= 32 * 4 * LN_MAXV, "layernorm: C=%d unsupported", C);
    SRK_REQUIRE(ld32 % 4 == 0 && ld32 >= C, "layernorm: bad ld32");
    SRK_REQUIRE(!out16 || (ld16 % 4 == 0 && ld16 >= C && ld16 <= 32 * 4 * LN_MAXV), "layernorm: bad ld16");
    SRK_REQUIRE((g == nullptr) == (b == nullptr), "layernorm: gamma/beta must both be set or null");
    int T = 1;
    if (win_shift >= 0) {
        SRK_REQUIRE(H > 0 && W > 0 && H % 8 == 0 && W % 8 == 0 && (win_shift == 0 || win_shift == 4),
                    "layernorm: bad window geometry");
        T = H * W;
        SRK_REQUIRE(M % T == 0, "layernorm: M must be a multiple of H*W");
    }
    ProfScope ps(SRK_PROF_LAYERNORM, stream);
    layernorm_kernel<<<ceil_div(M, 8), 256, 0, (cudaStream_t)stream>>>(
        x, ld32, M, C, g, b, eps, (uint16_t*)out16, ld16, out16_dtype, x32_out, T, H, W, win_shift);
    SRK_LAUNCH_CHECK("layernorm_kernel");
    return 0;
}

extern "C" int srk_conv_in(const float* x, int B, int h, int w, int H, int W, float in_scale,
                           const float* wgt, const float* bias, int C, float* out32, int ld32,
                           void* out16, int ld16, int out16_dtype, void* stream) {
    SRK_REQUIRE(x && wgt && bias && (out32 || out16), "conv_in: null pointer");
    SRK_REQUIRE(B > 0 && h > 0 && w > 0 && H >= h && W >= w, "conv_in: bad shape");
    SRK_REQUIRE(H - h < h && W - w < w, "conv_in: reflect pad must be smaller than the image");
    SRK_REQUIRE(!out32 || (ld32 % 4 == 0 && ld32 >= C), "conv_in: bad ld32");
    SRK_REQUIRE(!out16 || (ld16 % 4 == 0 && ld16 >= C), "conv_in: bad ld16");
    const int ldmax = (out32 ? ld32 : 0) > (out16 ? ld16 : 0) ? (out32 ? ld32 : 0) : (out16 ? ld16 : 0);
    const long long total = (long long)B * H * W * (ldmax / 4);
    const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    ProfScope ps(SRK_PROF_CONV_IN, stream);
    conv_in_kernel<<<grid, 256, (size_t)C * 10 * sizeof(float), (cudaStream_t)stream>>>(
        x, B, h, w, H, W, in_scale, wgt, bias, C, out32, ld32, (uint16_t*)out16, ld16, out16_dtype);
    SRK_LAUNCH_CHECK("conv_in_kernel");
    return 0;
}


extern "C" int srk_conv_in_ln(const float* x, int B, int h, int w, int H, int W, float in_scale,
                              const float* wgt, const float* bias, int C, float* f0, float* xa, int ld32,
                              const float* g1, const float* b1, const float* g2, const float* b2,
                              void* a16, int ld16, int out16_dtype, int win_shift, int ln_pad_one, void* stream) {
    SRK_REQUIRE(x && wgt && bias && f0 && xa && g1 && b1 && g2 && b2 && a16, "conv_in_ln: null pointer");
    SRK_REQUIRE(!ln_pad_one || C + 2 <= ld16, "conv_in_ln: ln_pad_one needs two pad columns");
    SRK_REQUIRE(B > 0 && h > 0 && w > 0 && H >= h && W >= w && H - h < h && W - w < w, "conv_in_ln: bad shape");
    SRK_REQUIRE(H % 8 == 0 && W % 8 == 0 && (win_shift == -1 || win_shift == 0 || win_shift == 4), "conv_in_ln: bad window geometry");
    SRK_REQUIRE(ld32 % 64 == 0 && ld32 == ld16 && ld32 >= C && ld32 <= 256 && C % 2 == 0, "conv_in_ln: channels must be padded to a multiple of 64 (<= 256)");
    const long long npix = (long long)B * H * W;
    SRK_REQUIRE(npix < (1ll << 31), "conv_in_ln: B*H*W must be below 2^31");
    const long long blocks = (npix + 7) / 8;
    const int grid = (int)(blocks < 148 * 8 ? blocks : 148 * 8);
    const size_t smem = (size_t)10 * ld32 * sizeof(float);
    ProfScope ps(SRK_PROF_CONV_IN, stream);
    cudaStream_t st = (cudaStream_t)stream;
#define L(NP) conv_in_ln_kernel<NP><<<grid, 256, smem, st>>>(x, B, h, w, H, W, in_scale, wgt, bias, C, f0, xa, ld32, g1, b1, g2, b2, (uint16_t*)a16, ld16, out16_dtype, win_shift, ln_pad_one)
    switch (ld32 / 64) { case 1: L(1); break; case 2: L(2); break; case 3: L(3); break; default: L(4); break; }
#undef L
    SRK_LAUNCH_CHECK("conv_in_ln_kernel");
    return 0;
}

extern "C" int srk_conv_out(const void* a, int lda, int B, int H, int W, int Cin, const void* wgt,
                            float bias, float out_scale, float* y, int Hc, int Wc, void* stream) {
    SRK_REQUIRE(a && wgt && y, "conv_out: null pointer");
    SRK_REQUIRE(B > 0 && H > 0 && W > 0 && Hc > 0 && Wc > 0 && Hc <= H && Wc <= W, "conv_out: bad shape");
    if (lda != 64 || Cin != 64)
        return fail(SRK_ERR_UNSUPPORTED, "conv_out: built for 64 input channels (got Cin=%d, lda=%d)", Cin, lda);
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope ps(SRK_PROF_CONV_OUT, stream);
    const size_t smem = 2 * (size_t)CO_TILE_BYTES + 8 * 3 * 48 * sizeof(float);
    static bool attr[64] = {};
    if (first_use_on_device(attr)) SRK_CUDA(cudaFuncSetAttribute(conv_out_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long tiles = (long long)B * ((Hc + CO_TH - 1) / CO_TH) * ((Wc + CO_TW - 1) / CO_TW);
    const int grid = (int)(tiles < 2 * 148 ? tiles : 2 * 148);
    conv_out_mma_kernel<<<grid, CO_THREADS, smem, st>>>((const __half*)a, B, H, W, (const __half*)wgt, bias, out_scale, y, Hc, Wc);
    SRK_LAUNCH_CHECK("conv_out_mma_kernel");
    return 0;
}

extern "C" int srk_tail_border(const void* a, int B, int H, int W, int s, const srk_tail_fold* f,
                               float out_scale, float* y, int Hc, int Wc, void* stream) {
    SRK_REQUIRE(a && f && f->border_w && f->border_b && y, "tail_border: null pointer");
    SRK_REQUIRE(B > 0 && H >= 3 && W >= 3 && s >= 1 && s * s <= 64, "tail_border: needs H, W >= 3 and s*s <= 64");
    SRK_REQUIRE(f->w_scale > 0.f, "tail_border: bad w_scale");
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope ps(SRK_PROF_CONV_OUT, stream);
    const size_t smem = 16 * (size_t)TB_STRIDE;
    static bool attr[64] = {};
    if (first_use_on_device(attr)) SRK_CUDA(cudaFuncSetAttribute(tail_border_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(2 * ((W - 2 + 15) / 16) + 2 * ((H - 2 + 15) / 16) + 4, B);
    tail_border_kernel<<<grid, TB_THREADS, smem, st>>>((const __half*)a, H, W, s, (const __half*)f->border_w, f->border_b,
                                                out_scale / f->w_scale, y, Hc, Wc);
    SRK_LAUNCH_CHECK("tail_border_kernel");
    return 0;
}

extern "C" int srk_bicubic_upsample(const float* x, int B, int h, int w, int s, float* y, void* stream) {
    SRK_REQUIRE(x && y, "bicubic_upsample: null pointer");
    SRK_REQUIRE(B > 0 && B <= 65535 && h > 0 && w > 0 && s >= 1 && s <= 16, "bicubic_upsample: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope ps(SRK_PROF_CONV_OUT, stream);
    dim3 grid(ceil_div((long long)w * s, 32), ceil_div((long long)h * s, 8), B);
    bicubic_up_kernel<<<grid, 256, 0, st>>>(x, h, w, s, y);
    SRK_LAUNCH_CHECK("bicubic_up_kernel");
    return 0;
}
