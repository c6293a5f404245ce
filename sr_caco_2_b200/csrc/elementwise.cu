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
// conv_out: 3x3, Cin (<= 64 per pass) -> 1 channel on an NHWC fp16 image, fp32 weights and
// accumulation (conv_last network_swinir.py:868,942; EDSR tail.1 network_nlsn.py:352-356).
// 8 lanes share one pixel (16 B = 8 channels each, so a warp reads 512 contiguous bytes per
// tap), a warp produces 32 consecutive output pixels and stores them as one 128 B line.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
conv_out_kernel(const __half* __restrict__ a, int lda, int B, int H, int W, int Cin,
                const float* __restrict__ wgt, float bias, float out_scale, float* __restrict__ y,
                int Hc, int Wc) {
    extern __shared__ __align__(16) float ws[];      // [9][cpad] weights, zero padded to 64-multiples
    const int cpad = ((Cin + 63) / 64) * 64;
    for (int i = threadIdx.x; i < 9 * cpad; i += blockDim.x) {
        const int tp = i / cpad, c = i - tp * cpad;
        ws[i] = c < Cin ? wgt[tp * Cin + c] : 0.f;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int sub = lane & 7, grp = lane >> 3;
    const int xsegs = (Wc + 31) >> 5;
    const long long nwarps = (long long)B * Hc * xsegs;
    const long long wstride = (long long)gridDim.x * (blockDim.x >> 5);
    const int cpasses = cpad / 64;
    for (long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
         wid < nwarps; wid += wstride) {
        const int xs = (int)(wid % xsegs);
        const int oy = (int)((wid / xsegs) % Hc);
        const int bi = (int)(wid / ((long long)xsegs * Hc));
        float mine = 0.f;
#pragma unroll 1
        for (int it = 0; it < 8; ++it) {
            const int ox = xs * 32 + it * 4 + grp;
            float acc = 0.f;
            for (int cp = 0; cp < cpasses; ++cp) {
                const int c0 = cp * 64 + sub * 8;
                if (ox < Wc && c0 < lda) {
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy) {
                        const int yy = oy + dy - 1;
                        if (yy < 0 || yy >= H) continue;
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx) {
                            const int xx = ox + dx - 1;
                            if (xx < 0 || xx >= W) continue;
                            const uint4 q = __ldg(reinterpret_cast<const uint4*>(
                                a + (((size_t)bi * H + yy) * W + xx) * lda + c0));
                            const __half2* hp = reinterpret_cast<const __half2*>(&q);
                            const float4 w0 = *reinterpret_cast<const float4*>(ws + (dy * 3 + dx) * cpad + c0);
                            const float4 w1 = *reinterpret_cast<const float4*>(ws + (dy * 3 + dx) * cpad + c0 + 4);
                            const float2 f0 = __half22float2(hp[0]), f1 = __half22float2(hp[1]);
                            const float2 f2 = __half22float2(hp[2]), f3 = __half22float2(hp[3]);
                            acc = fmaf(w0.x, f0.x, acc); acc = fmaf(w0.y, f0.y, acc);
                            acc = fmaf(w0.z, f1.x, acc); acc = fmaf(w0.w, f1.y, acc);
                            acc = fmaf(w1.x, f2.x, acc); acc = fmaf(w1.y, f2.y, acc);
                            acc = fmaf(w1.z, f3.x, acc); acc = fmaf(w1.w, f3.y, acc);
                        }
                    }
                }
            }
            acc += __shfl_xor_sync(0xffffffffu, acc, 1);
            acc += __shfl_xor_sync(0xffffffffu, acc, 2);
            acc += __shfl_xor_sync(0xffffffffu, acc, 4);
            // lane L must hold pixel L: it is produced at iteration L>>2 by lane group L&3
            const float tv = __shfl_sync(0xffffffffu, acc, (lane & 3) * 8);
            if ((lane >> 2) == it) mine = tv;
        }
        const int ox = xs * 32 + lane;
        if (ox < Wc) y[((size_t)bi * Hc + oy) * Wc + ox] = (mine + bias) * out_scale;
    }
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

extern "C" int srk_conv_out(const void* a, int lda, int B, int H, int W, int Cin, const float* wgt,
                            float bias, float out_scale, float* y, int Hc, int Wc, void* stream) {
    SRK_REQUIRE(a && wgt && y, "conv_out: null pointer");
    SRK_REQUIRE(B > 0 && H > 0 && W > 0 && Hc > 0 && Wc > 0 && Hc <= H && Wc <= W, "conv_out: bad shape");
    SRK_REQUIRE(lda % 8 == 0 && Cin <= lda, "conv_out: lda must be a multiple of 8 and >= Cin");
    const long long nwarps = (long long)B * Hc * ((Wc + 31) / 32);
    const long long blocks = (nwarps + 7) / 8;
    const int grid = (int)(blocks < 148 * 8 ? blocks : 148 * 8);
    const size_t smem = (size_t)9 * ((Cin + 63) / 64) * 64 * sizeof(float);
    SRK_REQUIRE(smem <= 48 * 1024, "conv_out: Cin too large");
    ProfScope ps(SRK_PROF_CONV_OUT, stream);
    conv_out_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>((const __half*)a, lda, B, H, W, Cin, wgt,
                                                            bias, out_scale, y, Hc, Wc);
    SRK_LAUNCH_CHECK("conv_out_kernel");
    return 0;
}
