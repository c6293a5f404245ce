// Single-pass PSNR / MSE / NRMSE / SSIM / PSNR_Y kernel (HBM bound: reads E and H once).
//
// Restates, for 1-channel images, dlib/utils/utils_image.py:369-372 (uint8 quantisation),
// :843-891 (PSNR), :894-934 (MSE), :937-1007 (NRMSE), :1010-1198 (SSIM, 11x11 sigma 1.5 valid
// Gaussian filter, fp32) and the PSNR_Y of dlib/utils/utils_trainer.py:1005-1012, for the full
// image and up to 8 ROI thresholds (roi = H8 >= th, utils_trainer.py:986) in the same pass.
//
// Three kernels share the finalize step:
//   metrics_stream_kernel  the hot path (quantised, nested level-set ROIs): row streaming, a CTA walks a strip of 128
//                          columns top to bottom with the vertical filter taps in registers (see its header)
//   metrics_fast_kernel    round 1's version of the same contract, kept as the in-library cross-check: one CTA owns a
//                          32x32 tile, loads the 42x42 halo tile of E and H once, accumulates the squared-error / min /
//                          max terms while loading, runs the separable Gaussian in shared memory, reduces with warp
//                          shuffles to one atomic per quantity
//   metrics_tile_kernel    the general contract (explicit ROI masks, unquantised inputs, unsorted thresholds), fp64 sums
#include "common.cuh"
#include <math.h>

namespace srk {

constexpr int MT = 32;            // tile of SSIM outputs / owned pixels
constexpr int MH = MT + 10;       // halo tile
constexpr int MAXV = 1 + SRK_MAX_ROI_THS;
constexpr int NTHREADS = 256;

// per (image, variant) accumulator record
struct MetAcc {
    double sse, sse_y, ssim;
    unsigned long long cnt, scnt;
    int mn_key, mx_key;     // ordered-int keys of min/max of y*roi
};
struct MetImg {
    int mn_all_key;         // min over the cropped target
    int range_bad;          // any input outside [0,255] (after optional quantisation: never)
};

__device__ __forceinline__ int f2key(float f) {
    int b = __float_as_int(f);
    return b >= 0 ? b : b ^ 0x7fffffff;
}
__device__ __forceinline__ float key2f(int k) {
    return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff);
}

__device__ __forceinline__ float quant255(float v) {
    // (x.clamp(0,1)*255).round().clamp(0,255) ; rintf = round half to even like torch.round
    v = fminf(fmaxf(v, 0.f), 1.f);
    return fminf(fmaxf(rintf(__fmul_rn(v, 255.f)), 0.f), 255.f);
}

// Y of a gray value replicated to RGB, same fp32 operation order as mb_gpu_rgb2ycbcr on
// (v/255) followed by *255 (utils_trainer.py:1005-1012); no fused multiply-adds.
__device__ __forceinline__ float luma255(float v) {
    float t = __fmul_rn(__fdiv_rn(v, 255.f), 255.f);
    float s = __fadd_rn(__fadd_rn(__fmul_rn(65.481f, t), __fmul_rn(128.553f, t)),
                        __fmul_rn(24.966f, t));
    float y = __fadd_rn(__fdiv_rn(s, 255.f), 16.f);
    float o = fminf(fmaxf(__fdiv_rn(y, 255.f), 0.f), 1.f);
    return __fmul_rn(o, 255.f);
}

__constant__ float c_gauss[11];

template <typename T>
__device__ __forceinline__ T block_reduce_sum(T v, T* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    T r = 0;
    if (wid == 0) {
        r = lane < NTHREADS / 32 ? red[lane] : T(0);
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    }
    return r;  // valid on thread 0
}

__device__ __forceinline__ int block_reduce_min(int v, int* red, bool is_max) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        int t = __shfl_xor_sync(0xffffffffu, v, o);
        v = is_max ? max(v, t) : min(v, t);
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    int r = v;
    if (wid == 0) {
        r = red[lane < NTHREADS / 32 ? lane : 0];
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            int t = __shfl_xor_sync(0xffffffffu, r, o);
            r = is_max ? max(r, t) : min(r, t);
        }
    }
    return r;
}

struct MetParams {
    const float* E; const float* H; const float* roi;
    int B, Hpx, Wpx, border, quantize, n_var;   // n_var = 1 + n_ths (or 1 with external roi)
    float ths[SRK_MAX_ROI_THS];
    MetAcc* acc; MetImg* img;
};

template <int NV, bool EXT_ROI>
__global__ void __launch_bounds__(NTHREADS)
metrics_tile_kernel(const MetParams p) {
    extern __shared__ __align__(16) unsigned char met_smem[];
    typedef float TileRow[MH + 1];
    typedef float HbRow[MT + 1];
    TileRow* sx = reinterpret_cast<TileRow*>(met_smem);            // E/255      [MH][MH+1]
    TileRow* sy = sx + MH;                                          // H/255
    TileRow* sq = sy + MH;                                          // H in [0,255] or external roi
    HbRow (*hb)[MH] = reinterpret_cast<HbRow (*)[MH]>(sq + MH);     // [5][MH][MT+1] moments
    __shared__ double red_d[NTHREADS / 32];
    __shared__ unsigned long long red_u[NTHREADS / 32];
    __shared__ int red_i[NTHREADS / 32];

    const int b = blockIdx.z;
    const int Hc = p.Hpx - 2 * p.border, Wc = p.Wpx - 2 * p.border;
    const int r0 = blockIdx.y * MT, c0 = blockIdx.x * MT;
    const float* Eb = p.E + (size_t)b * p.Hpx * p.Wpx;
    const float* Hb = p.H + (size_t)b * p.Hpx * p.Wpx;
    const float* Rb = EXT_ROI ? p.roi + (size_t)b * p.Hpx * p.Wpx : nullptr;

    double sse[NV], ssey[NV];
    unsigned int cnt[NV];
    int mnk[NV], mxk[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        sse[v] = 0.0; ssey[v] = 0.0; cnt[v] = 0;
        mnk[v] = 0x7fffffff; mxk[v] = (int)0x80000000;
    }
    int mn_all = 0x7fffffff;
    int bad = 0;

    // ---- load halo tile, accumulate owned-pixel terms -------------------------------------
    for (int i = threadIdx.x; i < MH * MH; i += NTHREADS) {
        const int lr = i / MH, lc = i - lr * MH;
        const int r = r0 + lr, c = c0 + lc;
        float e = 0.f, h = 0.f, w = 0.f;
        if (r < Hc && c < Wc) {
            const size_t off = (size_t)(r + p.border) * p.Wpx + (c + p.border);
            e = __ldg(Eb + off);
            h = __ldg(Hb + off);
            if (!isfinite(e) || !isfinite(h)) bad |= 2;        // before the clamp of quant255() hides a NaN
            if (p.quantize) { e = quant255(e); h = quant255(h); }
            if (EXT_ROI) w = __ldg(Rb + off);
            if (!(e >= 0.f && e <= 255.f) || !(h >= 0.f && h <= 255.f)) bad |= 1;
            if (lr < MT && lc < MT) {           // owned pixel
                const double d = (double)e - (double)h;
                const double dy = (double)luma255(e) - (double)luma255(h);
                mn_all = min(mn_all, f2key(h));
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    float wv;
                    if (EXT_ROI) wv = w;
                    else wv = (v == 0) ? 1.f : (h >= p.ths[v - 1 < 0 ? 0 : v - 1] ? 1.f : 0.f);
                    const double dw = d * (double)wv, dyw = dy * (double)wv;
                    sse[v] += dw * dw;
                    ssey[v] += dyw * dyw;
                    cnt[v] += (wv != 0.f) ? 1u : 0u;
                    const int k = f2key(h * wv);
                    mnk[v] = min(mnk[v], k);
                    mxk[v] = max(mxk[v], k);
                }
            }
        }
        sx[lr][lc] = __fdiv_rn(e, 255.f);
        sy[lr][lc] = __fdiv_rn(h, 255.f);
        sq[lr][lc] = EXT_ROI ? w : h;
    }
    __syncthreads();

    // ---- horizontal Gaussian of x, y, x*x, y*y, x*y ------------------------------------------
    for (int i = threadIdx.x; i < MH * MT; i += NTHREADS) {
        const int lr = i / MT, lc = i - lr * MT;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
#pragma unroll
        for (int t = 0; t < 11; ++t) {
            const float g = c_gauss[t];
            const float x = sx[lr][lc + t], y = sy[lr][lc + t];
            a0 = fmaf(g, x, a0); a1 = fmaf(g, y, a1);
            a2 = fmaf(g, x * x, a2); a3 = fmaf(g, y * y, a3); a4 = fmaf(g, x * y, a4);
        }
        hb[0][lr][lc] = a0; hb[1][lr][lc] = a1; hb[2][lr][lc] = a2;
        hb[3][lr][lc] = a3; hb[4][lr][lc] = a4;
    }
    __syncthreads();

    // ---- vertical Gaussian + SSIM map ---------------------------------------------------------
    float ssum[NV];
    unsigned int scnt[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) { ssum[v] = 0.f; scnt[v] = 0; }
    const int mapH = Hc - 10, mapW = Wc - 10;
    for (int i = threadIdx.x; i < MT * MT; i += NTHREADS) {
        const int lr = i / MT, lc = i - lr * MT;
        if (r0 + lr >= mapH || c0 + lc >= mapW) continue;
        float mx = 0.f, my = 0.f, xx = 0.f, yy = 0.f, xy = 0.f;
#pragma unroll
        for (int t = 0; t < 11; ++t) {
            const float g = c_gauss[t];
            mx = fmaf(g, hb[0][lr + t][lc], mx); my = fmaf(g, hb[1][lr + t][lc], my);
            xx = fmaf(g, hb[2][lr + t][lc], xx); yy = fmaf(g, hb[3][lr + t][lc], yy);
            xy = fmaf(g, hb[4][lr + t][lc], xy);
        }
        const float c1 = 1e-4f, c2 = 9e-4f;
        const float mxx = mx * mx, myy = my * my, mxy = mx * my;
        const float sxx = xx - mxx, syy = yy - myy, sxy = xy - mxy;
        const float cs = (2.f * sxy + c2) / (sxx + syy + c2);
        const float ss = ((2.f * mxy + c1) / (mxx + myy + c1)) * cs;
        const float q = sq[lr + 5][lc + 5];      // ROI is cropped by the filter radius
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            float wv;
            if (EXT_ROI) wv = q;
            else wv = (v == 0) ? 1.f : (q >= p.ths[v - 1 < 0 ? 0 : v - 1] ? 1.f : 0.f);
            ssum[v] += ss * wv;
            scnt[v] += (wv != 0.f) ? 1u : 0u;
        }
    }

    // ---- block reduction, one atomic per quantity -------------------------------------------
    MetAcc* acc = p.acc + (size_t)b * NV;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        double t0 = block_reduce_sum<double>(sse[v], red_d);
        if (threadIdx.x == 0) atomicAdd(&acc[v].sse, t0);
        double t1 = block_reduce_sum<double>(ssey[v], red_d);
        if (threadIdx.x == 0) atomicAdd(&acc[v].sse_y, t1);
        double t2 = block_reduce_sum<double>((double)ssum[v], red_d);
        if (threadIdx.x == 0) atomicAdd(&acc[v].ssim, t2);
        unsigned long long t3 = block_reduce_sum<unsigned long long>(cnt[v], red_u);
        if (threadIdx.x == 0) atomicAdd(&acc[v].cnt, t3);
        unsigned long long t4 = block_reduce_sum<unsigned long long>(scnt[v], red_u);
        if (threadIdx.x == 0) atomicAdd(&acc[v].scnt, t4);
        int k0 = block_reduce_min(mnk[v], red_i, false);
        if (threadIdx.x == 0) atomicMin(&acc[v].mn_key, k0);
        int k1 = block_reduce_min(mxk[v], red_i, true);
        if (threadIdx.x == 0) atomicMax(&acc[v].mx_key, k1);
    }
    int k2 = block_reduce_min(mn_all, red_i, false);
    if (threadIdx.x == 0) atomicMin(&p.img[b].mn_all_key, k2);
    const int any_range = __syncthreads_or(bad & 1), any_nonfinite = __syncthreads_or(bad & 2);
    if (threadIdx.x == 0 && (any_range || any_nonfinite)) atomicOr(&p.img[b].range_bad, (any_range ? 1 : 0) | (any_nonfinite ? 2 : 0));
}

__global__ void metrics_init_kernel(MetAcc* acc, MetImg* img, int B, int NV) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B * NV) {
        acc[i].sse = 0.0; acc[i].sse_y = 0.0; acc[i].ssim = 0.0;
        acc[i].cnt = 0ull; acc[i].scnt = 0ull;
        acc[i].mn_key = 0x7fffffff; acc[i].mx_key = (int)0x80000000;
    }
    if (i < B) { img[i].mn_all_key = 0x7fffffff; img[i].range_bad = 0; }
}

__global__ void metrics_finalize_kernel(const MetAcc* acc, const MetImg* img, int B, int NV,
                                        int Hc, int Wc, int ext_roi, double* out,
                                        int32_t* flags) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * NV) return;
    const int b = i / NV, v = i - b * NV;
    const MetAcc a = acc[i];
    const bool full = (v == 0) && !ext_roi;
    double n = full ? (double)Hc * (double)Wc : (double)a.cnt;
    if (n == 0.0) n = 1.0;                                   // empty ROI
    const double mse = a.sse / n, mse_y = a.sse_y / n;
    const double psnr = 20.0 * log10(255.0 / sqrt(fmax(mse, 1e-45)));
    const double psnr_y = 20.0 * log10(255.0 / sqrt(fmax(mse_y, 1e-45)));
    double mn = (double)key2f(a.mn_key), mx = (double)key2f(a.mx_key);
    if (!full) mn = fmax(mn, (double)key2f(img[b].mn_all_key));
    double den = mx - mn;
    if (den == 0.0) den = 1.0;
    const double nrmse = sqrt(mse) / den;
    double sn = full ? (double)(Hc - 10) * (double)(Wc - 10) : (double)a.scnt;
    if (sn == 0.0) sn = 1.0;
    // the reference keeps SSIM in fp32
    const double ssim = (double)(float)(a.ssim / sn);
    double* o = out + (size_t)i * SRK_MET_N;
    o[SRK_MET_PSNR] = psnr; o[SRK_MET_MSE] = mse; o[SRK_MET_NRMSE] = nrmse;
    o[SRK_MET_SSIM] = ssim; o[SRK_MET_PSNR_Y] = psnr_y;
    int f = 0;
    const double vals[5] = {psnr, mse, nrmse, ssim, psnr_y};
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        if (!isfinite(vals[k])) f |= 1;
        if (vals[k] < 0.0) f |= 2;
    }
    if (img[b].range_bad & 1) f |= 4;
    if (img[b].range_bad & 2) f |= 1;                      // non-finite input pixel
    if (f) atomicOr(&flags[b], f);
}


// ------------------------------------------------------------------------------------------
// Hot-path variant (quantize = 1, thresholded ROIs): integer arithmetic on the uint8 levels.
// The ROIs roi_v = (H8 >= th_v) are nested, so every pixel is dropped into ONE bucket
// k = #{v : H8 >= th_v} (thresholds sorted ascending by the host) and the per-variant sums are
// suffix sums over buckets, formed in the finalize kernel.  Squared errors are exact integers;
// PSNR_Y follows from Y = 219/255 * v + 16 (65.481 + 128.553 + 24.966 = 219), i.e.
// mse_y = mse * (219/255)^2, which equals the reference's fp32 luma path to ~1e-7 dB.
// The separable Gaussian runs as 4-output sliding windows held in registers.
// ------------------------------------------------------------------------------------------
struct BucketAcc {                       // per (image, bucket)
    unsigned long long sse, cnt, scnt;
    double ssim;
    int mn, mx;                          // min / max level inside the bucket (255 / -1 when empty)
};

__global__ void metrics_fast_init_kernel(BucketAcc* acc, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { acc[i].sse = 0ull; acc[i].cnt = 0ull; acc[i].scnt = 0ull; acc[i].ssim = 0.0; acc[i].mn = 255; acc[i].mx = -1; }
}

// v / 255 for an integer level v in [0, 255], rounded exactly like the reference's division
// (one Newton step on v * (1/255): checked against __fdiv_rn for all 256 levels, tests/test_gpu.py)
__device__ __forceinline__ float div255(float v) {
    const float r = 1.f / 255.f;
    const float q = v * r;
    return fmaf(fmaf(-q, 255.f, v), r, q);
}

struct FastParams {
    const float* E; const float* H;
    const unsigned char* H8;      // optional: the target as stored uint8 levels (H is then unused)
    int B, Hpx, Wpx, border, n_ths;
    float ths[SRK_MAX_ROI_THS];
    BucketAcc* acc;
    int32_t* flags;               // bit 0 is also raised for a non-finite INPUT pixel (the reference's NaN / Inf then reaches its metrics)
};

__global__ void __launch_bounds__(NTHREADS, 3)
metrics_fast_kernel(const FastParams p) {
    extern __shared__ __align__(16) unsigned char met_smem[];
    typedef float TileRow[MH + 2];
    typedef float HbRow[MT + 1];
    TileRow* sx = reinterpret_cast<TileRow*>(met_smem);             // E8/255   [MH][MH+2]
    TileRow* sy = sx + MH;                                           // H8/255
    HbRow (*hb)[MH] = reinterpret_cast<HbRow (*)[MH]>(sy + MH);      // [5][MH][MT+1]
    unsigned char* sk = reinterpret_cast<unsigned char*>(&hb[5][0][0]);   // bucket of every halo pixel [MH][MH]
    __shared__ unsigned long long r_sse[MAXV], r_cnt[MAXV], r_scnt[MAXV];
    __shared__ double r_ssim[MAXV];
    __shared__ int r_mn[MAXV], r_mx[MAXV];

    const int tid = threadIdx.x, lane = tid & 31;
    const int b = blockIdx.z;
    const int Hc = p.Hpx - 2 * p.border, Wc = p.Wpx - 2 * p.border;
    const int r0 = blockIdx.y * MT, c0 = blockIdx.x * MT;
    const float* Eb = p.E + (size_t)b * p.Hpx * p.Wpx;
    const float* Hb = p.H8 ? nullptr : p.H + (size_t)b * p.Hpx * p.Wpx;
    const unsigned char* H8b = p.H8 ? p.H8 + (size_t)b * p.Hpx * p.Wpx : nullptr;
    const int NB = p.n_ths + 1;

    if (tid < MAXV) { r_sse[tid] = 0ull; r_cnt[tid] = 0ull; r_scnt[tid] = 0ull; r_ssim[tid] = 0.0; r_mn[tid] = 255; r_mx[tid] = -1; }
    __syncthreads();

    unsigned int cnt[MAXV], sse[MAXV];
    int mn[MAXV], mx[MAXV];
#pragma unroll
    for (int k = 0; k < MAXV; ++k) { cnt[k] = 0; sse[k] = 0; mn[k] = 255; mx[k] = -1; }

    // all global loads of the thread first (7 halo pixels x 2 images in flight), then the arithmetic
    constexpr int NPIX = (MH * MH + NTHREADS - 1) / NTHREADS;
    float ev[NPIX], hv[NPIX];
#pragma unroll
    for (int u = 0; u < NPIX; ++u) {
        const int i = tid + u * NTHREADS;
        const int lr = i / MH, lc = i - lr * MH;
        const int r = r0 + lr, c = c0 + lc;
        ev[u] = 0.f; hv[u] = 0.f;
        if (i < MH * MH && r < Hc && c < Wc) {
            const size_t off = (size_t)(r + p.border) * p.Wpx + (c + p.border);
            ev[u] = __ldg(Eb + off);
            hv[u] = H8b ? (float)__ldg(H8b + off) : __ldg(Hb + off);
        }
    }
    {   // tensor2uint82float (utils_image.py:369-372) propagates NaN into every metric; the clamp in quant255() would
        // hide it (fminf / fmaxf return the non-NaN operand), so the raw loads are tested here
        bool nonfinite = false;
#pragma unroll
        for (int u = 0; u < NPIX; ++u) nonfinite = nonfinite || !isfinite(ev[u]) || !isfinite(hv[u]);
        if (__any_sync(0xffffffffu, nonfinite) && lane == 0) atomicOr(&p.flags[b], 1);
    }
#pragma unroll
    for (int u = 0; u < NPIX; ++u) {
        const int i = tid + u * NTHREADS;
        if (i >= MH * MH) break;
        const int lr = i / MH, lc = i - lr * MH;
        const int r = r0 + lr, c = c0 + lc;
        int e = 0, h = 0, kb = 0;
        if (r < Hc && c < Wc) {
            const float hf = H8b ? hv[u] : quant255(hv[u]);
            e = (int)quant255(ev[u]);
            h = (int)hf;
#pragma unroll
            for (int t = 0; t < SRK_MAX_ROI_THS; ++t) kb += (t < p.n_ths && hf >= p.ths[t]) ? 1 : 0;
            if (lr < MT && lc < MT) {                                // owned pixel
                const int d = e - h;
                const unsigned int d2 = (unsigned int)(d * d);
#pragma unroll
                for (int k = 0; k < MAXV; ++k) {
                    const bool in = kb == k;
                    cnt[k] += in ? 1u : 0u;
                    sse[k] += in ? d2 : 0u;
                    mn[k] = in ? min(mn[k], h) : mn[k];
                    mx[k] = in ? max(mx[k], h) : mx[k];
                }
            }
        }
        sx[lr][lc] = div255((float)e);
        sy[lr][lc] = div255((float)h);
        sk[lr * MH + lc] = (unsigned char)kb;
    }
    __syncthreads();

    // ---- horizontal pass: (row, 4 consecutive outputs) per work item ----
    for (int w = tid; w < MH * (MT / 4); w += NTHREADS) {
        const int lr = w / (MT / 4), seg = w - lr * (MT / 4);
        const int cb = seg * 4;
        float x[14], y[14];
#pragma unroll
        for (int t = 0; t < 14; ++t) { x[t] = sx[lr][cb + t]; y[t] = sy[lr][cb + t]; }
        float a[5][4];
#pragma unroll
        for (int q = 0; q < 5; ++q)
#pragma unroll
            for (int o = 0; o < 4; ++o) a[q][o] = 0.f;
#pragma unroll
        for (int t = 0; t < 14; ++t) {
            const float xx = x[t] * x[t], yy = y[t] * y[t], xy = x[t] * y[t];
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                const int tap = t - o;
                if (tap >= 0 && tap < 11) {
                    const float gq = c_gauss[tap];
                    a[0][o] = fmaf(gq, x[t], a[0][o]); a[1][o] = fmaf(gq, y[t], a[1][o]);
                    a[2][o] = fmaf(gq, xx, a[2][o]); a[3][o] = fmaf(gq, yy, a[3][o]); a[4][o] = fmaf(gq, xy, a[4][o]);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 5; ++q)
#pragma unroll
            for (int o = 0; o < 4; ++o) hb[q][lr][cb + o] = a[q][o];
    }
    __syncthreads();

    // ---- vertical pass + SSIM: (column, 4 consecutive rows) per thread ----
    float ssum[MAXV];
    unsigned int scnt[MAXV];
#pragma unroll
    for (int k = 0; k < MAXV; ++k) { ssum[k] = 0.f; scnt[k] = 0; }
    const int mapH = Hc - 10, mapW = Wc - 10;
    {
        const int lc = tid & 31, rb = (tid >> 5) * 4;
        float a[5][4];
#pragma unroll
        for (int q = 0; q < 5; ++q)
#pragma unroll
            for (int o = 0; o < 4; ++o) a[q][o] = 0.f;
#pragma unroll
        for (int t = 0; t < 14; ++t) {
            float v[5];
#pragma unroll
            for (int q = 0; q < 5; ++q) v[q] = hb[q][rb + t][lc];
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                const int tap = t - o;
                if (tap >= 0 && tap < 11) {
                    const float gq = c_gauss[tap];
#pragma unroll
                    for (int q = 0; q < 5; ++q) a[q][o] = fmaf(gq, v[q], a[q][o]);
                }
            }
        }
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            const int lr = rb + o;
            if (r0 + lr < mapH && c0 + lc < mapW) {
                const float c1 = 1e-4f, c2 = 9e-4f;
                const float mxm = a[0][o], mym = a[1][o];
                const float mxx = mxm * mxm, myy = mym * mym, mxy = mxm * mym;
                const float sxx = a[2][o] - mxx, syy = a[3][o] - myy, sxy = a[4][o] - mxy;
                const float cs = (2.f * sxy + c2) / (sxx + syy + c2);
                const float ss = ((2.f * mxy + c1) / (mxx + myy + c1)) * cs;
                const int kb = sk[(lr + 5) * MH + lc + 5];           // ROI is cropped by the filter radius
#pragma unroll
                for (int k = 0; k < MAXV; ++k) {
                    const bool in = kb == k;
                    ssum[k] += in ? ss : 0.f;
                    scnt[k] += in ? 1u : 0u;
                }
            }
        }
    }

    // ---- reduction: warp shuffles -> shared atomics -> one global atomic per quantity ----
#pragma unroll
    for (int k = 0; k < MAXV; ++k) {
        if (k < NB) {
            // integer quantities: one REDUX instruction each (sm_80+); only the fp32 SSIM sum needs the butterfly
            const unsigned int c_ = __reduce_add_sync(0xffffffffu, cnt[k]);
            const unsigned int s_ = __reduce_add_sync(0xffffffffu, sse[k]);      // <= 32 * 4 * 65025 fits 32 bits
            const unsigned int sc_ = __reduce_add_sync(0xffffffffu, scnt[k]);
            const int mn_ = __reduce_min_sync(0xffffffffu, mn[k]);
            const int mx_ = __reduce_max_sync(0xffffffffu, mx[k]);
            float f_ = ssum[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) f_ += __shfl_xor_sync(0xffffffffu, f_, o);
            if (lane == 0) {
                if (c_) { atomicAdd(&r_cnt[k], (unsigned long long)c_); atomicAdd(&r_sse[k], (unsigned long long)s_);
                          atomicMin(&r_mn[k], mn_); atomicMax(&r_mx[k], mx_); }
                if (sc_) { atomicAdd(&r_scnt[k], (unsigned long long)sc_); atomicAdd(&r_ssim[k], (double)f_); }
            }
        }
    }
    __syncthreads();
    if (tid < NB) {
        BucketAcc* a = p.acc + (size_t)b * MAXV + tid;
        if (r_cnt[tid]) { atomicAdd(&a->cnt, r_cnt[tid]); atomicAdd(&a->sse, r_sse[tid]);
                          atomicMin(&a->mn, r_mn[tid]); atomicMax(&a->mx, r_mx[tid]); }
        if (r_scnt[tid]) { atomicAdd(&a->scnt, r_scnt[tid]); atomicAdd(&a->ssim, r_ssim[tid]); }
    }
}

// ------------------------------------------------------------------------------------------
// Row-streaming version of the hot path (round 2).  A CTA of 128 threads owns a strip of 128 input columns
// (118 SSIM-map columns) of one image and streams a segment of its rows top to bottom:
//   vertical pass : thread = column.  The last 12 rows of the four maps (x, y), (x^2 + y^2, x y) live in registers as
//                   packed fp32x2 pairs (slot = row % 12, static through a 12-row unrolled body); every new row yields
//                   the column's vertically filtered values with 22 FFMA2, written as one 16 B chunk per column
//   horizontal    : every 4 rows, thread = (row of the batch, 4 adjacent outputs): 14 chunks stream through a register
//                   sliding window (88 FFMA2), then the SSIM formula and the ROI-bucket sums
// sigma_x^2 + sigma_y^2 only ever appears as a sum, so x^2 + y^2 is filtered as ONE map (4 maps instead of 5).
// Bookkeeping is kept off the FMA-bound path: quantisation rounds with the 1.5 * 2^23 add (same ties-to-even as
// rintf, and the level is then the low byte of the float's bits: no F2I / I2F / FRND, which run at a quarter of the
// rate), the ROI bucket of a level comes from a 256-byte table, and the per-bucket sums live in shared memory, one
// private word per (bucket, thread): squared errors as exact integers packed with their count (sse << 8 | cnt: a
// thread owns <= 254 rows), SSIM sums as (fp32 sum, count).  The finalize kernel and the scratch layout of the tile
// version are unchanged.  NRMSE only needs the extrema of the whole cropped target (the ROIs are upper level sets, see
// the finalize kernel), so every CTA writes the extrema of all pixels it read into the buckets it owns pixels of.
// ------------------------------------------------------------------------------------------
constexpr int ST_THREADS = 128;          // input columns per strip
constexpr int ST_OW = 118;               // SSIM-map columns per strip (128 - 10)
constexpr int ST_VSTRIDE = 170;          // 16 B chunks per staged row: column c at chunk c + c / 4 (<= 131 + 32 read); the
                                         // stride (2 mod 8) keeps the (2 groups x 4 rows) of a quarter warp on 8 banks
constexpr int ST_KROWS = 16, ST_KSTRIDE = 136;   // ring of ROI-bucket rows: column c at byte c + 3 (4-byte aligned reads)
constexpr float ST_MAGIC = 12582912.f;   // 1.5 * 2^23: t + MAGIC rounds t in [0, 2^22) to an integer, ties to even
constexpr int ST_MAGIC_BITS = 0x4B400000;

struct StreamParams {
    FastParams f;
    int seg_rows;                        // SSIM-map rows per row segment (<= 244)
};

template <bool H8IN>
__global__ void __launch_bounds__(ST_THREADS, 4)
metrics_stream_kernel(const StreamParams sp) {
    const FastParams& p = sp.f;
    __shared__ __align__(16) float4 vbuf[2][4][ST_VSTRIDE];         // [buffer][row of the batch][padded column]
    __shared__ __align__(16) unsigned char kring[ST_KROWS][ST_KSTRIDE];
    __shared__ __align__(16) unsigned char lut[256];                // level -> ROI bucket (# thresholds <= level)
    __shared__ unsigned int sacc[MAXV][ST_THREADS];                 // per (bucket, thread): sse << 8 | count of owned pixels
    __shared__ float2 ssacc[MAXV][ST_THREADS];                      // per (bucket, thread): SSIM sum, count (as bits)
    __shared__ unsigned long long r_sse[MAXV], r_cnt[MAXV], r_scnt[MAXV];
    __shared__ double r_ssim[MAXV];
    __shared__ int r_mn, r_mx;

    const int tid = threadIdx.x, lane = tid & 31;
    const int b = blockIdx.x, seg = blockIdx.y, strip = blockIdx.z;   // strips slowest: a ragged (cheap) last strip is scheduled last
    const int Hc = p.Hpx - 2 * p.border, Wc = p.Wpx - 2 * p.border;
    const int mapH = Hc - 10, mapW = Wc - 10;
    const int c0 = strip * ST_OW;                                     // first input / output column of the strip (cropped frame)
    const int r0 = seg * sp.seg_rows;                                 // first input / output row of the segment
    const bool last_strip = strip == (int)gridDim.z - 1, last_seg = seg == (int)gridDim.y - 1;
    const int out_rows = min(sp.seg_rows, mapH - r0);                 // SSIM rows of this segment
    const int in_rows = out_rows + 10;
    const int NB = p.n_ths + 1;
    const float* Eb = p.E + (size_t)b * p.Hpx * p.Wpx;
    const float* Hb = H8IN ? nullptr : p.H + (size_t)b * p.Hpx * p.Wpx;
    const unsigned char* H8b = H8IN ? p.H8 + (size_t)b * p.Hpx * p.Wpx : nullptr;

    if (tid < MAXV) { r_sse[tid] = 0ull; r_cnt[tid] = 0ull; r_scnt[tid] = 0ull; r_ssim[tid] = 0.0; }
    if (tid == 0) { r_mn = 255; r_mx = -1; }
#pragma unroll
    for (int k = 0; k < MAXV; ++k) { sacc[k][tid] = 0u; ssacc[k][tid] = make_float2(0.f, 0.f); }
    for (int l = tid; l < 256; l += ST_THREADS) {
        int kb = 0;
#pragma unroll
        for (int t = 0; t < SRK_MAX_ROI_THS; ++t) kb += (t < p.n_ths && (float)l >= p.ths[t]) ? 1 : 0;
        lut[l] = (unsigned char)kb;
    }

    // ---- vertical-pass state: this thread's column ----
    const int col = c0 + tid;                                         // cropped-frame column
    const bool col_in = col < Wc;
    // squared errors are owned by exactly one strip / segment (the halo columns / rows belong to the neighbours)
    const bool col_owned = col_in && (tid < ST_OW || last_strip);
    const int own_rows = col_owned ? (last_seg ? in_rows : sp.seg_rows) : 0;
    // rows past the segment / columns past the image re-read the last valid pixel: they only feed masked outputs, and
    // a real pixel keeps the extrema and the non-finite test right without a predicate on the load
    const size_t col_off = (size_t)(p.border + r0) * p.Wpx + (size_t)(col_in ? col : Wc - 1) + p.border;
    const float* Ec = Eb + col_off;                                   // row i of the segment: + i * Wpx (32-bit: Hpx * Wpx < 2^31 is checked at launch)
    const float* Hcp = H8IN ? nullptr : Hb + col_off;
    const unsigned char* H8c = H8IN ? H8b + col_off : nullptr;
    const bool v_active = c0 + (tid & ~31) < Wc;                      // warp-uniform: any column of this warp inside the image
    uint64_t ring[12][2];                                             // [row % 12][(x, y) | (x^2 + y^2, x y)]
#pragma unroll
    for (int j = 0; j < 12; ++j) { ring[j][0] = 0ull; ring[j][1] = 0ull; }
    int hmin = ST_MAGIC_BITS + 255, hmax = ST_MAGIC_BITS - 1;         // extrema of the target's float bits (monotonic in the level)
    float nf = 0.f;                                                   // sum of 0 * pixel: NaN iff any pixel read was NaN / Inf
    uint64_t g2[11];
#pragma unroll
    for (int k = 0; k < 11; ++k) g2[k] = dup2(c_gauss[k]);
    unsigned int* my_sacc = &sacc[0][tid];
    float2* my_ssacc = &ssacc[0][tid];

    // ---- horizontal-pass state: (row of the batch, 4 adjacent outputs); a warp = 8 groups x 4 rows ----
    const int hrow = tid & 3, hgrp = tid >> 2;
    const bool h_active = c0 + 4 * (hgrp & ~7) < mapW && 4 * (hgrp & ~7) < ST_OW;     // warp-uniform
    unsigned int omask = 0;                                           // which of this thread's 4 outputs exist
#pragma unroll
    for (int o = 0; o < 4; ++o) omask |= (4 * hgrp + o < ST_OW && c0 + 4 * hgrp + o < mapW) ? (1u << o) : 0u;

    // raw pixels of the next batch of 4 rows (requested one batch ahead)
    float e_nx[4];
    float h_nx[4];
    unsigned int h8_nx[4];
    auto request = [&](int i0) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int off = min(i0 + u, in_rows - 1) * p.Wpx;
            e_nx[u] = __ldg(Ec + off);
            if (H8IN) h8_nx[u] = __ldg(H8c + off); else h_nx[u] = __ldg(Hcp + off);
        }
    };
    if (v_active) request(0);
    __syncthreads();

    for (int base = 0; base < in_rows; base += 12) {
#pragma unroll
        for (int bq = 0; bq < 3; ++bq) {                              // three batches of 4 rows = one turn of the 12-row ring
            const int i0 = base + 4 * bq;
            if (i0 >= in_rows) break;
            const int buf = (i0 >> 2) & 1;
            // ---------------- vertical pass: 4 rows ----------------
            if (v_active) {
                float e_cur[4], h_cur[4];
                unsigned int h8_cur[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) { e_cur[u] = e_nx[u]; if (H8IN) h8_cur[u] = h8_nx[u]; else h_cur[u] = h_nx[u]; }
                request(i0 + 4);
                unsigned char* krow = &kring[i0 & (ST_KROWS - 1)][tid + 3];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = 4 * bq + u;                         // ring slot (static)
                    // tensor2uint82float (utils_image.py:369-372): (x.clamp(0,1) * 255).round()
                    nf = fmaf(e_cur[u], 0.f, nf);
                    const float em = __fadd_rn(__fmul_rn(__saturatef(e_cur[u]), 255.f), ST_MAGIC);
                    const int be = __float_as_int(em);
                    int bh;
                    float hm;
                    if (H8IN) {
                        bh = ST_MAGIC_BITS | (int)h8_cur[u];
                        hm = __int_as_float(bh);
                    } else {
                        nf = fmaf(h_cur[u], 0.f, nf);
                        hm = __fadd_rn(__fmul_rn(__saturatef(h_cur[u]), 255.f), ST_MAGIC);
                        bh = __float_as_int(hm);
                    }
                    const int kb = lut[bh & 0xff];
                    krow[u * ST_KSTRIDE] = (unsigned char)kb;
                    const int d = be - bh;
                    const unsigned int w = (i0 + u < own_rows) ? (unsigned int)(d * d) * 256u + 1u : 0u;
                    atomicAdd(my_sacc + kb * ST_THREADS, w);          // private word: no contention, no read-back dependency
                    hmin = min(hmin, bh); hmax = max(hmax, bh);
                    const float x = div255(__fadd_rn(em, -ST_MAGIC)), y = div255(__fadd_rn(hm, -ST_MAGIC));
                    ring[j][0] = pack64(__float_as_uint(x), __float_as_uint(y));
                    ring[j][1] = pack64(__float_as_uint(fmaf(x, x, y * y)), __float_as_uint(x * y));
                    // vertically filtered values of output row i - 10 (rows i-10 .. i = slots j+2 .. j+12 mod 12)
                    uint64_t v0 = 0ull, v1 = 0ull;
#pragma unroll
                    for (int k = 0; k < 11; ++k) {
                        const int sl = (j + 2 + k) % 12;
                        v0 = fma2(g2[k], ring[sl][0], v0);
                        v1 = fma2(g2[k], ring[sl][1], v1);
                    }
                    uint32_t a0, a1, a2, a3;
                    unpack64(v0, a0, a1); unpack64(v1, a2, a3);
                    vbuf[buf][u][tid + (tid >> 2)] = make_float4(__uint_as_float(a0), __uint_as_float(a1), __uint_as_float(a2), __uint_as_float(a3));
                }
            }
            __syncthreads();
            // ---------------- horizontal pass + SSIM: output row (i0 + hrow - 10), columns 4 hgrp .. 4 hgrp + 3 ----------------
            const int orow = i0 + hrow - 10;
            if (h_active && orow >= 0 && orow < out_rows && omask) {
                uint64_t acc[4][2];
#pragma unroll
                for (int o = 0; o < 4; ++o) { acc[o][0] = 0ull; acc[o][1] = 0ull; }
                const float4* vrow = &vbuf[buf][hrow][5 * hgrp];      // column 4 hgrp + k sits at chunk 5 hgrp + k + k / 4
#pragma unroll
                for (int k = 0; k < 14; ++k) {
                    const float4 ch = vrow[k + (k >> 2)];
                    const uint64_t w0 = pack64(__float_as_uint(ch.x), __float_as_uint(ch.y));
                    const uint64_t w1 = pack64(__float_as_uint(ch.z), __float_as_uint(ch.w));
#pragma unroll
                    for (int o = 0; o < 4; ++o) {
                        const int tap = k - o;
                        if (tap >= 0 && tap < 11) {
                            acc[o][0] = fma2(g2[tap], w0, acc[o][0]);
                            acc[o][1] = fma2(g2[tap], w1, acc[o][1]);
                        }
                    }
                }
                // ROI buckets of the 4 window centres: input row orow + 5, columns 4 hgrp + 5 .. + 8 (bytes 4 hgrp + 8 .. of the row)
                const unsigned int kb4 = *reinterpret_cast<const unsigned int*>(&kring[(i0 + hrow - 5) & (ST_KROWS - 1)][4 * hgrp + 8]);
                float ss[4];
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    uint32_t u0, u1, u2, u3;
                    unpack64(acc[o][0], u0, u1); unpack64(acc[o][1], u2, u3);
                    const float mxm = __uint_as_float(u0), mym = __uint_as_float(u1);
                    const float ess = __uint_as_float(u2), exy = __uint_as_float(u3);
                    const float c1 = 1e-4f, c2 = 9e-4f;
                    const float mxx = mxm * mxm, myy = mym * mym, mxy = mxm * mym;
                    const float svar = (ess - mxx) - myy, sxy = exy - mxy;
                    float rden;                                       // denominator >= c1 * c2 = 9e-8: no denormal fix-up needed
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rden) : "f"((mxx + myy + c1) * (svar + c2)));
                    const float v = (2.f * mxy + c1) * (2.f * sxy + c2) * rden;
                    ss[o] = (omask >> o & 1u) ? v : 0.f;
                }
                if (kb4 == (kb4 & 0xffu) * 0x01010101u) {             // one bucket (the common case inside / outside a ROI)
                    float2* a = my_ssacc + (kb4 & 0xffu) * ST_THREADS;
                    float2 t = *a;
                    t.x += (ss[0] + ss[1]) + (ss[2] + ss[3]);
                    t.y = __uint_as_float(__float_as_uint(t.y) + __popc(omask));
                    *a = t;
                } else {
#pragma unroll
                    for (int o = 0; o < 4; ++o) {
                        float2* a = my_ssacc + ((kb4 >> (8 * o)) & 0xffu) * ST_THREADS;
                        float2 t = *a;
                        t.x += ss[o];
                        t.y = __uint_as_float(__float_as_uint(t.y) + (omask >> o & 1u));
                        *a = t;
                    }
                }
            }
            // (the other buffer is written by the next batch; this one is rewritten two batches later, after the next barrier)
        }
    }
    if (__any_sync(0xffffffffu, !(nf == 0.f)) && lane == 0) atomicOr(&p.flags[b], 1);
    __syncthreads();                                                  // sacc atomics of the last batch

    // ---- reduction: warp -> shared atomics -> one global atomic per quantity ----
    if (!v_active) { hmin = ST_MAGIC_BITS + 255; hmax = ST_MAGIC_BITS - 1; }
    const int wmn = __reduce_min_sync(0xffffffffu, hmin) - ST_MAGIC_BITS, wmx = __reduce_max_sync(0xffffffffu, hmax) - ST_MAGIC_BITS;
    if (lane == 0) { atomicMin(&r_mn, wmn); atomicMax(&r_mx, wmx); }
    for (int k = 0; k < NB; ++k) {
        const unsigned int w = sacc[k][tid];
        const float2 sv = ssacc[k][tid];
        const unsigned int c_ = __reduce_add_sync(0xffffffffu, w & 0xffu);
        const unsigned int s_ = __reduce_add_sync(0xffffffffu, w >> 8);      // <= 32 * 254 * 65025 fits 32 bits
        const unsigned int sc_ = __reduce_add_sync(0xffffffffu, __float_as_uint(sv.y));
        float f_ = sv.x;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) f_ += __shfl_xor_sync(0xffffffffu, f_, o);
        if (lane == 0) {
            if (c_) { atomicAdd(&r_cnt[k], (unsigned long long)c_); atomicAdd(&r_sse[k], (unsigned long long)s_); }
            if (sc_) { atomicAdd(&r_scnt[k], (unsigned long long)sc_); atomicAdd(&r_ssim[k], (double)f_); }
        }
    }
    __syncthreads();
    if (tid < NB) {
        BucketAcc* a = p.acc + (size_t)b * MAXV + tid;
        if (r_cnt[tid]) { atomicAdd(&a->cnt, r_cnt[tid]); atomicAdd(&a->sse, r_sse[tid]);
                          atomicMin(&a->mn, r_mn); atomicMax(&a->mx, r_mx); }
        if (r_scnt[tid]) { atomicAdd(&a->scnt, r_scnt[tid]); atomicAdd(&a->ssim, r_ssim[tid]); }
    }
}

__global__ void metrics_fast_finalize_kernel(const BucketAcc* acc, int B, int n_ths, int Hc, int Wc,
                                             double* out, int32_t* flags) {
    const int NV = n_ths + 1;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * NV) return;
    const int b = i / NV, v = i - b * NV;
    const BucketAcc* a = acc + (size_t)b * MAXV;
    // variant v = union of buckets k >= v  (variant 0 = whole image)
    unsigned long long sse = 0, cnt = 0, scnt = 0, outside = 0;
    double ssim = 0.0;
    int mn_in = 255, mx_in = -1, mn_all = 255;
    for (int k = 0; k < NV; ++k) {
        if (a[k].cnt) mn_all = min(mn_all, a[k].mn);
        if (k >= v) {
            sse += a[k].sse; cnt += a[k].cnt; scnt += a[k].scnt; ssim += a[k].ssim;
            mn_in = min(mn_in, a[k].mn); mx_in = max(mx_in, a[k].mx);
        } else {
            outside += a[k].cnt;
        }
    }
    double n = (v == 0) ? (double)Hc * (double)Wc : (double)cnt;
    if (n == 0.0) n = 1.0;
    const double mse = (double)sse / n;
    const double ky = 219.0 / 255.0;
    const double mse_y = mse * ky * ky;
    const double psnr = 20.0 * log10(255.0 / sqrt(fmax(mse, 1e-45)));
    const double psnr_y = 20.0 * log10(255.0 / sqrt(fmax(mse_y, 1e-45)));
    // extrema of y*roi: pixels outside the ROI contribute 0; min additionally floored by min(y)
    double mx = cnt ? (double)mx_in : 0.0;
    double mn = cnt ? (double)mn_in : 0.0;
    if (outside) { mn = fmin(mn, 0.0); mx = fmax(mx, 0.0); }
    if (v > 0) mn = fmax(mn, (double)mn_all);
    double den = mx - mn;
    if (den == 0.0) den = 1.0;
    const double nrmse = sqrt(mse) / den;
    double sn = (v == 0) ? (double)(Hc - 10) * (double)(Wc - 10) : (double)scnt;
    if (sn == 0.0) sn = 1.0;
    const double ssv = (double)(float)(ssim / sn);
    double* o = out + (size_t)i * SRK_MET_N;
    o[SRK_MET_PSNR] = psnr; o[SRK_MET_MSE] = mse; o[SRK_MET_NRMSE] = nrmse;
    o[SRK_MET_SSIM] = ssv; o[SRK_MET_PSNR_Y] = psnr_y;
    int f = 0;
    const double vals[5] = {psnr, mse, nrmse, ssv, psnr_y};
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        if (!isfinite(vals[k])) f |= 1;
        if (vals[k] < 0.0) f |= 2;
    }
    if (f) atomicOr(&flags[b], f);
}

bool g_metrics_tile_path = false;            // srk_metrics_use_tile_kernel(): tests compare the two hot-path kernels
int g_metrics_ctas_per_sm = 6;               // row segments are sized for this many streaming CTAs per SM (bits 8.. of the same call)
static int num_sms_metrics() {
    int dev = 0, v = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    return v;
}

static int upload_gauss() {
    static bool done[64] = {false};
    int dev = 0;
    SRK_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && done[dev]) return 0;
    // utils_image.py:1102-1117: 2-D exp(-(x^2+y^2)/(2 sigma^2)) / sum  ==  outer(g1, g1)
    float e[11], s = 0.f, g[11];
    for (int i = 0; i < 11; ++i) { float c = (float)i - 5.f; e[i] = expf(-(c * c) / (2.f * 1.5f * 1.5f)); s += e[i]; }
    for (int i = 0; i < 11; ++i) g[i] = e[i] / s;
    SRK_CUDA(cudaMemcpyToSymbol(c_gauss, g, sizeof(g)));
    if (dev < 64) done[dev] = true;
    return 0;
}

constexpr size_t MET_SMEM = sizeof(float) * (3 * MH * (MH + 1) + 5 * MH * (MT + 1));

template <bool EXT>
static void launch_tile(int NV, dim3 grid, cudaStream_t st, const MetParams& p) {
    switch (NV) {
#define C(n) case n:                                                                          \
        cudaFuncSetAttribute(metrics_tile_kernel<n, EXT>,                                     \
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MET_SMEM);     \
        metrics_tile_kernel<n, EXT><<<grid, NTHREADS, MET_SMEM, st>>>(p); break;
        C(1) C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9)
#undef C
    }
}

static int run_metrics(const float* E, const float* H, const float* roi, int B, int Hpx, int Wpx,
                       int border, int quantize, const int* roi_ths, int n_ths, double* out,
                       int32_t* flags, void* scratch, cudaStream_t st, const unsigned char* H8 = nullptr) {
    SRK_REQUIRE(E && (H || H8) && out && flags && scratch, "metrics: null pointer");
    SRK_REQUIRE(B > 0 && Hpx > 0 && Wpx > 0 && border >= 0, "metrics: bad shape");
    SRK_REQUIRE(n_ths >= 0 && n_ths <= SRK_MAX_ROI_THS, "metrics: n_ths must be in [0,%d]",
                SRK_MAX_ROI_THS);
    const int Hc = Hpx - 2 * border, Wc = Wpx - 2 * border;
    // utils_image.py:1044-1047: the 11x11 window must fit
    SRK_REQUIRE(Hc >= 11 && Wc >= 11,
                "metrics: kernel size can't be greater than actual input size (%dx%d after border %d)",
                Hc, Wc, border);
    SRK_REQUIRE((long long)Hpx * Wpx < (1ll << 31), "metrics: image too large (%dx%d)", Hpx, Wpx);
    if (int rc = upload_gauss()) return rc;
    bool sorted = true;
    for (int i = 1; i < n_ths; ++i) sorted = sorted && roi_ths[i] > roi_ths[i - 1];
    if (quantize && !roi && sorted) {
        // hot path: integer / bucketed kernel
        FastParams fp{};
        fp.E = E; fp.H = H; fp.H8 = H8; fp.B = B; fp.Hpx = Hpx; fp.Wpx = Wpx; fp.border = border; fp.n_ths = n_ths;
        for (int i = 0; i < n_ths; ++i) fp.ths[i] = (float)roi_ths[i];
        fp.acc = reinterpret_cast<BucketAcc*>(scratch);
        fp.flags = flags;
        ProfScope ps(SRK_PROF_METRICS, st);
        SRK_CUDA(cudaMemsetAsync(flags, 0, sizeof(int32_t) * B, st));
        metrics_fast_init_kernel<<<ceil_div((long long)B * MAXV, 128), 128, 0, st>>>(fp.acc, B * MAXV);
        SRK_LAUNCH_CHECK("metrics_fast_init_kernel");
        if (g_metrics_tile_path) {                     // round-1 tile kernel: kept as the in-library cross-check of the streaming kernel
            const size_t smem = sizeof(float) * (2 * MH * (MH + 2) + 5 * MH * (MT + 1)) + MH * MH;
            static bool attr[64] = {};
            if (first_use_on_device(attr)) {
                SRK_CUDA(cudaFuncSetAttribute(metrics_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                SRK_CUDA(cudaFuncSetAttribute(metrics_fast_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            }
            dim3 grid(ceil_div(Wc, MT), ceil_div(Hc, MT), B);
            metrics_fast_kernel<<<grid, NTHREADS, smem, st>>>(fp);
            SRK_LAUNCH_CHECK("metrics_fast_kernel");
        } else {
            // streaming kernel: strips of 118 SSIM columns x row segments; segments sized for ~6 CTAs per SM (4 resident: the
            // rest balances the SMs; measured 82 / 77 / 76 / 78 us at 4 / 5 / 6 / 8 for 32 x 512 x 512), <= 244 rows each
            const int mapH = Hc - 10, mapW = Wc - 10;
            const int strips = ceil_div(mapW, ST_OW);
            int nseg = ceil_div(g_metrics_ctas_per_sm * num_sms_metrics(), (long long)strips * B);
            if (nseg < ceil_div(mapH, 244)) nseg = ceil_div(mapH, 244);
            if (nseg > ceil_div(mapH, 32)) nseg = ceil_div(mapH, 32);
            if (nseg < 1) nseg = 1;
            StreamParams sp{};
            sp.f = fp;
            sp.seg_rows = ceil_div(mapH, nseg);
            dim3 grid(B, ceil_div(mapH, sp.seg_rows), strips);
            if (H8) metrics_stream_kernel<true><<<grid, ST_THREADS, 0, st>>>(sp);
            else metrics_stream_kernel<false><<<grid, ST_THREADS, 0, st>>>(sp);
            SRK_LAUNCH_CHECK("metrics_stream_kernel");
        }
        metrics_fast_finalize_kernel<<<ceil_div((long long)B * (n_ths + 1), 128), 128, 0, st>>>(fp.acc, B, n_ths, Hc, Wc, out, flags);
        SRK_LAUNCH_CHECK("metrics_fast_finalize_kernel");
        return 0;
    }
    if (H8) return fail(SRK_ERR_UNSUPPORTED, "metrics: uint8 targets need the quantised path with ascending thresholds");
    const int NV = roi ? 1 : 1 + n_ths;
    MetParams p{};
    p.E = E; p.H = H; p.roi = roi; p.B = B; p.Hpx = Hpx; p.Wpx = Wpx; p.border = border;
    p.quantize = quantize; p.n_var = NV;
    for (int i = 0; i < n_ths; ++i) p.ths[i] = (float)roi_ths[i];
    p.acc = reinterpret_cast<MetAcc*>(scratch);
    p.img = reinterpret_cast<MetImg*>(reinterpret_cast<char*>(scratch) +
                                      align_up(sizeof(MetAcc) * (size_t)B * NV, 16));
    ProfScope ps(SRK_PROF_METRICS, st);
    SRK_CUDA(cudaMemsetAsync(flags, 0, sizeof(int32_t) * B, st));
    metrics_init_kernel<<<ceil_div((long long)B * NV, 128), 128, 0, st>>>(p.acc, p.img, B, NV);
    SRK_LAUNCH_CHECK("metrics_init_kernel");
    dim3 grid(ceil_div(Wc, MT), ceil_div(Hc, MT), B);
    if (roi) launch_tile<true>(NV, grid, st, p); else launch_tile<false>(NV, grid, st, p);
    SRK_LAUNCH_CHECK("metrics_tile_kernel");
    metrics_finalize_kernel<<<ceil_div((long long)B * NV, 128), 128, 0, st>>>(
        p.acc, p.img, B, NV, Hc, Wc, roi ? 1 : 0, out, flags);
    SRK_LAUNCH_CHECK("metrics_finalize_kernel");
    return 0;
}

}  // namespace srk

extern "C" int srk_metrics_use_tile_kernel(int on) {
    srk::g_metrics_tile_path = (on & 1) != 0;
    srk::g_metrics_ctas_per_sm = (on >> 8) > 0 ? (on >> 8) : 6;
    return 0;
}

extern "C" size_t srk_metrics_scratch_bytes(int B, int n_ths) {
    const int NV = 1 + (n_ths < 0 ? 0 : n_ths);
    const size_t generic = srk::align_up(sizeof(srk::MetAcc) * (size_t)B * NV, 16) + sizeof(srk::MetImg) * (size_t)B + 64;
    const size_t fast = sizeof(srk::BucketAcc) * (size_t)B * srk::MAXV + 64;
    return generic > fast ? generic : fast;
}

extern "C" int srk_metrics(const float* E, const float* H, int B, int Hpx, int Wpx, int border,
                           int quantize, const int* roi_ths, int n_ths, double* out,
                           int32_t* flags, void* scratch, void* stream) {
    return srk::run_metrics(E, H, nullptr, B, Hpx, Wpx, border, quantize, roi_ths, n_ths, out,
                            flags, scratch, (cudaStream_t)stream);
}

extern "C" int srk_metrics_h8(const float* E, const unsigned char* H8, int B, int Hpx, int Wpx, int border,
                              const int* roi_ths, int n_ths, double* out, int32_t* flags, void* scratch,
                              void* stream) {
    if (!H8) return srk::fail(SRK_ERR_INVALID, "metrics_h8: null target");
    return srk::run_metrics(E, nullptr, nullptr, B, Hpx, Wpx, border, 1, roi_ths, n_ths, out, flags, scratch,
                            (cudaStream_t)stream, H8);
}

extern "C" int srk_metrics_roi(const float* E, const float* H, const float* roi, int B, int Hpx,
                               int Wpx, int border, int quantize, double* out, int32_t* flags,
                               void* scratch, void* stream) {
    if (!roi) return srk::fail(SRK_ERR_INVALID, "metrics_roi: null roi");
    return srk::run_metrics(E, H, roi, B, Hpx, Wpx, border, quantize, nullptr, 0, out, flags,
                            scratch, (cudaStream_t)stream);
}
