// Fused 8x8-window multi-head attention:  S = q k^T * scale + rel-pos bias + shift mask,
// softmax (fp32, warp-shuffle row reductions), O = P v.   One CTA per window, the window's
// q|k|v rows are staged once in shared memory (cp.async, coalesced 16 B chunks); each warp owns
// a 16-query strip of one head at a time and keeps S / P / O in registers (mma.sync m16n8k16,
// bf16 operands, fp32 accumulate); O overwrites the strip's own q slot and the CTA writes the
// window's output rows back coalesced.
//
// Restates WindowAttention.forward (dlib/models/network_swinir.py:150-176): q*scale, q@k^T,
// + relative_position_bias_table[relative_position_index] (:116-129, :156-162), + mask
// (calculate_mask :260-285: -100 where the 3x3 region labels of the shifted frame differ),
// softmax(-1), @ v, head merge (:176).  The attention matrix is never materialised in HBM.
#include "attention.cuh"
#include <stdlib.h>

namespace srk {

constexpr int ATT_THREADS = 128;          // 4 warps = the four 16-query strips of a window

// DP = padded head dim (16, 32, 48 or 64).  grid = (windows, head groups): a CTA stages only the
// q|k|v columns of its HG heads (small footprint -> many CTAs per SM hide the load latency).
template <int DP>
__global__ void __launch_bounds__(ATT_THREADS)
window_attention_kernel(const __nv_bfloat16* __restrict__ qkv, int ldq,
                        __nv_bfloat16* __restrict__ out, int ldo, const float* __restrict__ rel_table, int H, int W, int nH,
                        int HG, float scale, int shift) {
    extern __shared__ __align__(16) unsigned char att_smem[];
    const int h0 = blockIdx.y * HG;              // first head of this CTA
    const int nhl = min(HG, nH - h0);            // heads handled here
    const int seg = HG * DP;                     // local elements per q / k / v section
    const int RS = 3 * seg * 2 + 16;             // padded smem row stride (bytes)
    unsigned char* rows = att_smem;              // [64][RS]  local layout [q(HG x DP) | k | v]
    float* tab = reinterpret_cast<float*>(att_smem + 64 * RS);   // [HG][225]
    int* lab = reinterpret_cast<int*>(tab + HG * 225);            // [64]

    const int nW = (H >> 3) * (W >> 3);
    const int win_g = blockIdx.x;                // global window index (b * nW + win)
    const int win = win_g % nW;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- stage q|k|v of the window -------------------------------------------------------
    const __nv_bfloat16* src = qkv + (size_t)win_g * 64 * ldq;
    const int cps = nhl * DP / 8;                // 16 B chunks per section actually present
    const uint32_t rows_s = (uint32_t)__cvta_generic_to_shared(rows);
    {
        // (row, chunk) decomposition: with cps a power of two (the usual case) every thread keeps a
        // fixed 16 B column and walks down the rows with pointer increments
        const bool pow2 = (cps & (cps - 1)) == 0 && cps <= ATT_THREADS;
        if (pow2) {
            const int sh = 31 - __clz(cps);
            const int r0 = tid >> sh, cc = tid & (cps - 1), rstep = ATT_THREADS >> sh;
#pragma unroll
            for (int w = 0; w < 3; ++w) {
                const __nv_bfloat16* sp = src + (w * nH + h0) * DP + (size_t)r0 * ldq + cc * 8;
                uint32_t dp = rows_s + w * seg * 2 + r0 * RS + cc * 16;
                for (int r = r0; r < 64; r += rstep) {
                    cp_async16(dp, sp);
                    sp += (size_t)rstep * ldq;
                    dp += rstep * RS;
                }
            }
        } else {
            for (int w = 0; w < 3; ++w) {
                const __nv_bfloat16* sw = src + (w * nH + h0) * DP;
                const uint32_t dw = rows_s + w * seg * 2;
                for (int i = tid; i < 64 * cps; i += ATT_THREADS) {
                    const int r = i / cps, cc = i - r * cps;
                    cp_async16(dw + r * RS + cc * 16, sw + (size_t)r * ldq + cc * 8);
                }
            }
        }
    }
    asm volatile("cp.async.commit_group;\n");
    for (int i = tid; i < nhl * 225; i += ATT_THREADS) tab[i] = __ldg(rel_table + h0 * 225 + i);
    const int wpr = W >> 3;
    const int wi = win / wpr, wj = win - wi * wpr;
    const bool masked = shift > 0 && (wi == (H >> 3) - 1 || wj == wpr - 1);
    if (tid < 64) lab[tid] = masked ? win_pos_label(win, tid, H, W, shift) : 0;
    asm volatile("cp.async.wait_group 0;\n");
    __syncthreads();

    for (int h = 0; h < nhl; ++h)
        attn_unit<DP>(rows_s, rows, RS, warp & 3, h * DP, seg + h * DP, 2 * seg + h * DP, tab + h * 225, lab, masked,
                      scale, lane);
    __syncthreads();
    // ---- write the window's output rows (coalesced 16 B chunks, pad columns zeroed) -------------
    __nv_bfloat16* dst = out + (size_t)win_g * 64 * ldo + h0 * DP;
    const bool last = h0 + nhl >= nH;            // the last group also zeroes the pad columns
    const int valid = nhl * DP / 8;
    const int oc = last ? (ldo - h0 * DP) / 8 : valid;
    if ((oc & (oc - 1)) == 0 && oc <= ATT_THREADS) {
        const int sh = 31 - __clz(oc);
        const int r0w = tid >> sh, c = tid & (oc - 1), rstep = ATT_THREADS >> sh;
        const unsigned char* sp = rows + (size_t)r0w * RS + c * 16;
        __nv_bfloat16* dp = dst + (size_t)r0w * ldo + c * 8;
        for (int r = r0w; r < 64; r += rstep) {
            uint4 v = make_uint4(0, 0, 0, 0);
            if (c < valid) v = *reinterpret_cast<const uint4*>(sp);
            *reinterpret_cast<uint4*>(dp) = v;
            sp += (size_t)rstep * RS;
            dp += (size_t)rstep * ldo;
        }
    } else {
        for (int i = tid; i < 64 * oc; i += ATT_THREADS) {
            const int r = i / oc, c = i - r * oc;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (c < valid) v = *reinterpret_cast<const uint4*>(rows + (size_t)r * RS + c * 16);
            *reinterpret_cast<uint4*>(dst + (size_t)r * ldo + c * 8) = v;
        }
    }
}

}  // namespace srk

using namespace srk;

extern "C" int srk_window_attention(const void* qkv, int ldq, void* out, int ldo,
                                    const float* rel_table, int nB, int H, int W, int nH, int dp, float scale, int shift,
                                    void* stream) {
    SRK_REQUIRE(qkv && out && rel_table, "window_attention: null pointer");
    SRK_REQUIRE(nB > 0 && H > 0 && W > 0 && H % 8 == 0 && W % 8 == 0,
                "window_attention: H, W must be positive multiples of the window size 8");
    SRK_REQUIRE(shift == 0 || shift == 4, "window_attention: shift must be 0 or 4");
    SRK_REQUIRE(dp == 16 || dp == 32 || dp == 48 || dp == 64,
                "window_attention: padded head dim %d not in {16,32,48,64}", dp);
    SRK_REQUIRE(nH >= 1 && ldo % 8 == 0 && ldo >= nH * dp, "window_attention: bad nH/ldo");
    const int nq = 3 * nH * dp;
    SRK_REQUIRE(ldq % 8 == 0 && ldq >= nq, "window_attention: bad ldq");
    int HG = nH % 2 == 0 ? 2 : (nH % 3 == 0 ? 3 : 1);                  // heads per CTA
    if (HG > nH) HG = nH;
    const size_t smem = (size_t)64 * (3 * HG * dp * 2 + 16) + (size_t)HG * 225 * 4 + 64 * 4;
    SRK_REQUIRE(smem <= 227 * 1024, "window_attention: window does not fit shared memory");
    const dim3 grid(nB * (H / 8) * (W / 8), (nH + HG - 1) / HG);
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope ps(SRK_PROF_ATTENTION, stream);
#define LAUNCH(D)                                                                             \
    do {                                                                                      \
        SRK_CUDA(cudaFuncSetAttribute(window_attention_kernel<D>,                             \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        window_attention_kernel<D><<<grid, ATT_THREADS, smem, st>>>(                          \
            (const __nv_bfloat16*)qkv, ldq, (__nv_bfloat16*)out, ldo, rel_table, H, W, nH, HG, scale,  \
            shift);                                                                           \
    } while (0)
    if (dp == 16) LAUNCH(16);
    else if (dp == 32) LAUNCH(32);
    else if (dp == 48) LAUNCH(48);
    else LAUNCH(64);
#undef LAUNCH
    SRK_LAUNCH_CHECK("window_attention_kernel");
    return 0;
}
