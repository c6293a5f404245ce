// Legacy tensor-core GEMM engine (mma.sync m16n8k16, ldmatrix, cp.async 3-stage ring).
// It is NOT the product path: it is the in-library cross-check for the tcgen05 engine
// (tests run both engines on the same operands) and the "recompiled pre-Blackwell kernel"
// data point that bench.py --engine mma_sync reports.
#include "gemm_common.cuh"

namespace srk {

constexpr int MBM = 128, MBN = 64, MBK = 64, MSTAGES = 3, MTHREADS = 256;
constexpr int MSTAGE_BYTES = (MBM + MBN) * MBK * 2;

__device__ __forceinline__ void cp16_zfill(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(sz));
}
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
template <bool FP16>
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    if (FP16)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <bool FP16>
__global__ void __launch_bounds__(MTHREADS)
gemm_mma_kernel(const GemmP p) {
    extern __shared__ __align__(128) unsigned char msm[];
    const uint32_t smem = (uint32_t)__cvta_generic_to_shared(msm);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp & 3, wn = warp >> 2;
    const int m0 = blockIdx.x * MBM, n0 = blockIdx.y * MBN;
    const int nkb = p.K / MBK;

    // per-thread load coordinates: chunk c of rows r_j = tid/8 + 32 j
    const int lc = tid & 7, lr = tid >> 3;
    int py[4], px[4], pb[4];
    bool rv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int m = m0 + lr + 32 * j;
        rv[j] = m < p.M;
        py[j] = px[j] = pb[j] = 0;
        if (p.a_mode == SRK_A_CONV3X3 && rv[j]) {
            pb[j] = m / p.T;
            const int rem = m - pb[j] * p.T;
            py[j] = rem / p.W; px[j] = rem - py[j] * p.W;
        }
    }

    auto load_stage = [&](int kb, int stage) {
        const uint32_t sa = smem + stage * MSTAGE_BYTES, sb = sa + MBM * MBK * 2;
        int dy = 0, dx = 0, cb = kb;
        if (p.a_mode == SRK_A_CONV3X3) {
            const int tap = kb / p.cpb;
            cb = kb - tap * p.cpb;
            dy = tap / p.kt - (p.kt >> 1); dx = tap - (tap / p.kt) * p.kt - (p.kt >> 1);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = lr + 32 * j;
            const uint16_t* src = p.A;
            bool ok = rv[j];
            if (p.a_mode == SRK_A_CONV3X3) {
                const int yy = py[j] + dy, xx = px[j] + dx;
                ok = ok && yy >= 0 && yy < p.H && xx >= 0 && xx < p.W;
                if (ok) src = p.A + (((size_t)pb[j] * p.H + yy) * p.W + xx) * p.lda + cb * 64 + lc * 8;
            } else if (ok) {
                src = p.A + (size_t)(m0 + r) * p.lda + kb * 64 + lc * 8;
            }
            cp16_zfill(sa + r * 128 + ((lc ^ (r & 7)) << 4), src, ok);
        }
#pragma unroll
        for (int j = 0; j < MBN / 32; ++j) {
            const int r = lr + 32 * j;
            cp16_zfill(sb + r * 128 + ((lc ^ (r & 7)) << 4),
                       p.Wt + (size_t)(n0 + r) * p.K + kb * 64 + lc * 8, true);
        }
    };

    float acc[2][4][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;

#pragma unroll
    for (int s = 0; s < MSTAGES - 1; ++s) {
        if (s < nkb) load_stage(s, s);
        asm volatile("cp.async.commit_group;\n");
    }
    for (int kb = 0; kb < nkb; ++kb) {
        asm volatile("cp.async.wait_group %0;\n" ::"n"(MSTAGES - 2));
        __syncthreads();
        {
            const int nk = kb + MSTAGES - 1;
            if (nk < nkb) load_stage(nk, nk % MSTAGES);
            asm volatile("cp.async.commit_group;\n");
        }
        const int stage = kb % MSTAGES;
        const uint32_t sa = smem + stage * MSTAGE_BYTES, sb = sa + MBM * MBK * 2;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            uint32_t a[2][4], b[2][4];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi) {
                const int r = wm * 32 + mi * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                const int c = ks * 2 + (lane >> 4);
                ldsm4(a[mi], sa + r * 128 + ((c ^ (r & 7)) << 4));
            }
#pragma unroll
            for (int nj = 0; nj < 2; ++nj) {
                const int r = wn * 32 + nj * 16 + (lane & 7) + (lane >> 4) * 8;
                const int c = ks * 2 + ((lane >> 3) & 1);
                ldsm4(b[nj], sb + r * 128 + ((c ^ (r & 7)) << 4));
            }
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                for (int nj = 0; nj < 2; ++nj) {
                    mma16816<FP16>(acc[mi][2 * nj], a[mi], b[nj][0], b[nj][1]);
                    mma16816<FP16>(acc[mi][2 * nj + 1], a[mi], b[nj][2], b[nj][3]);
                }
        }
    }
    asm volatile("cp.async.wait_group 0;\n");

    // ---- epilogue ----------------------------------------------------------------------------
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            const int m = m0 + wm * 32 + mi * 16 + g + hf * 8;
            if (m >= p.M) continue;
            const int r32 = (p.res || p.out32) ? row32_of(p, m) : 0;
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                const int n = n0 + wn * 32 + ni * 8 + 2 * t;
                float v0 = acc[mi][ni][hf * 2 + 0] + p.bias[n];
                float v1 = acc[mi][ni][hf * 2 + 1] + p.bias[n + 1];
                v0 = apply_act(v0, p.act); v1 = apply_act(v1, p.act);
                if (p.res) {
                    const float2 rr = *reinterpret_cast<const float2*>(p.res + (size_t)r32 * p.ld32 + n);
                    v0 = v0 * p.res_scale + rr.x; v1 = v1 * p.res_scale + rr.y;
                }
                if (p.out32)
                    *reinterpret_cast<float2*>(p.out32 + (size_t)r32 * p.ld32 + n) = make_float2(v0, v1);
                if (p.out16)
                    *reinterpret_cast<uint32_t*>(p.out16 + off16_of(p, m, n)) = pack2(v0, v1, p.out16_dtype);
                if (p.img) {
                    size_t off;
                    if (offimg_of(p, m, n, off)) p.img[off] = v0 * p.img_scale;
                    if (offimg_of(p, m, n + 1, off)) p.img[off] = v1 * p.img_scale;
                }
            }
        }
    }
}

int validate_gemm(const srk_gemm_args* g) {
    SRK_REQUIRE(g && g->A && g->Wt && (g->bias || g->attn_table), "gemm: null operand");
    SRK_REQUIRE(g->M > 0 && g->N > 0 && g->K > 0, "gemm: bad M/N/K");
    SRK_REQUIRE(g->N % 16 == 0, "gemm: N=%d must be padded to a multiple of 16", g->N);
    SRK_REQUIRE(g->K % 64 == 0, "gemm: K=%d must be padded to a multiple of 64", g->K);
    SRK_REQUIRE(g->dtype == SRK_BF16 || g->dtype == SRK_FP16, "gemm: bad dtype");
    SRK_REQUIRE(g->lda % 8 == 0, "gemm: lda must be a multiple of 8");
    if (g->a_mode == SRK_A_CONV3X3) {
        SRK_REQUIRE(g->conv_k == 0 || g->conv_k == 3 || g->conv_k == 5, "gemm: conv_k must be 3 or 5");
        const int taps = g->conv_k == 5 ? 25 : 9;
        SRK_REQUIRE(g->lda % 64 == 0 && g->K == taps * g->lda, "gemm: conv needs lda %% 64 == 0 and K == %d*lda", taps);
        SRK_REQUIRE(g->nB > 0 && g->H > 0 && g->W > 0 && g->M == g->nB * g->H * g->W, "gemm: conv needs M == nB*H*W");
    } else {
        SRK_REQUIRE(g->a_mode == SRK_A_ROWS && g->lda >= g->K, "gemm: rows mode needs lda >= K");
    }
    const bool needs_hw = g->win_shift >= 0 || g->out16_mode == SRK_O16_PIXSHUF2 || g->img || g->ln_win_shift >= 0;
    if (needs_hw) {
        SRK_REQUIRE(g->H > 0 && g->W > 0 && g->M % (g->H * g->W) == 0, "gemm: H, W required and M %% (H*W) == 0");
    }
    if (g->win_shift >= 0 || g->ln_win_shift >= 0)
        SRK_REQUIRE(g->H % 8 == 0 && g->W % 8 == 0, "gemm: window mapping needs H, W multiples of 8");
    SRK_REQUIRE(g->win_shift == -1 || g->win_shift == 0 || g->win_shift == 4, "gemm: win_shift must be -1, 0 or 4");
    if (g->res || g->out32) SRK_REQUIRE(g->ld32 >= g->N && g->ld32 % 4 == 0, "gemm: bad ld32");
    if (g->out16) {
        if (g->out16_mode == SRK_O16_PIXSHUF2)
            SRK_REQUIRE(g->N % 4 == 0 && g->ld16 >= g->N / 4 && g->ld16 % 8 == 0 && (g->N / 4) % 8 == 0, "gemm: bad pixel-shuffle output");
        else
            SRK_REQUIRE(g->out16_mode == SRK_O16_ROWS && (g->ld16 >= g->N || g->attn_table) && g->ld16 % 8 == 0, "gemm: bad ld16");
    }
    if (g->img) SRK_REQUIRE(g->img_s > 0 && g->img_s * g->img_s <= g->N && g->img_hc > 0 && g->img_wc > 0, "gemm: bad image output");
    SRK_REQUIRE(g->out32 || g->out16 || g->img, "gemm: no output");
    if (g->attn_table) {
        SRK_REQUIRE(g->a_mode == SRK_A_ROWS && g->attn_heads >= 2 && g->attn_heads % 2 == 0 && g->N == 3 * g->attn_heads * 32,
                    "gemm: fused attention needs an even head count and N == 3*heads*32");
        SRK_REQUIRE(g->out16 && g->ld16 == g->attn_heads * 32 && g->out16_mode == SRK_O16_ROWS && g->out16_dtype == SRK_BF16 &&
                    g->dtype == SRK_BF16 && !g->res && !g->out32 && !g->ln_g && !g->img && g->act == SRK_ACT_NONE,
                    "gemm: fused attention writes only the bf16 attention output (ld16 == heads*32)");
        SRK_REQUIRE(g->H > 0 && g->W > 0 && g->H % 8 == 0 && g->W % 8 == 0 && g->M % 64 == 0 && g->M % (g->H * g->W) == 0 &&
                    (g->attn_shift == 0 || g->attn_shift == 4), "gemm: fused attention needs window-major rows of whole images");
    }
    return 0;
}

int gemm_mma_sync(const srk_gemm_args* g, cudaStream_t st) {
    if (g->ln_g) return fail(SRK_ERR_UNSUPPORTED, "gemm(mma.sync): fused LayerNorm epilogue is tcgen05-only");
    if (g->attn_table) return fail(SRK_ERR_UNSUPPORTED, "gemm(mma.sync): fused attention epilogue is tcgen05-only");
    SRK_REQUIRE(g->N % MBN == 0, "gemm(mma.sync): N=%d must be a multiple of %d", g->N, MBN);
    GemmP p = make_gemm_params(g);
    const size_t smem = (size_t)MSTAGES * MSTAGE_BYTES;
    dim3 grid(ceil_div(g->M, MBM), g->N / MBN);
    if (g->dtype == SRK_FP16) {
        SRK_CUDA(cudaFuncSetAttribute(gemm_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        gemm_mma_kernel<true><<<grid, MTHREADS, smem, st>>>(p);
    } else {
        SRK_CUDA(cudaFuncSetAttribute(gemm_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        gemm_mma_kernel<false><<<grid, MTHREADS, smem, st>>>(p);
    }
    SRK_LAUNCH_CHECK("gemm_mma_kernel");
    return 0;
}

}  // namespace srk
