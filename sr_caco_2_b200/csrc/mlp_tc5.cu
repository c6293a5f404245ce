// Fused Swin MLP:  x' = x + fc2(GELU(fc1(A) + b1)) + b2,  then LayerNorm (or a plain cast) of x'
// for the next GEMM -- ONE persistent tcgen05 kernel, the 128 x hidden activation never leaves
// the SM.  Restates Mlp.forward (dlib/models/network_swinir.py:39-45) + the residual add and the
// next norm1 (:335, :293) of SwinTransformerBlock.forward.
//
// Per 128-token tile (hidden processed in chunks of 128 columns, c = 0 .. NC-1):
//   TMA        : A tile (LN2 output, bf16) once; W1 / W2 tiles stream through a 4-slot ring in
//                exactly the order the MMA warp consumes them.
//   MMA warp   : fc1(c):  D1[c&1] (TMEM, 128 cols)  = A . W1[128c.., :]^T
//                fc2(c):  D2      (TMEM, CP cols)  += H[c&1] . W2[:, 128c..]^T
//                issued as fc1(0), fc1(1), fc2(0), fc1(2), fc2(1), ... so the GELU of chunk c
//                overlaps the fc1 MMAs of chunk c+1.
//   16 epilogue warps: GELU stage: tcgen05.ld D1 -> + b1 -> GELU -> bf16 -> written into shared
//                memory in the 128B-swizzled K-major operand layout (H[c&1]) that fc2 reads;
//                final stage: D2 -> padded fp32 staging (aliases the H buffers, 64 rows at a
//                time) -> + b2 + residual (fp32 stream, register-prefetched one tile ahead) ->
//                x' store + LayerNorm / cast store, all as coalesced row segments.
#include "tc5_ptx.cuh"
#include <stdlib.h>

namespace srk {

constexpr int ML_EPI_WARPS = 8;            // 2 warps per TMEM lane group
constexpr int ML_QN = ML_EPI_WARPS / 4;    // warps per lane group
constexpr int ML_RPW = 64 / ML_EPI_WARPS;  // rows per warp in each 64-row half of the final stage
constexpr int ML_THREADS = 64 + 32 * ML_EPI_WARPS;
constexpr int ML_WSLOT = 24576, ML_WSLOTS = 4;
constexpr int ML_SMEM_TOTAL = 227 * 1024;

struct MlpP {
    int M, C, hid_p, NC, m_tiles;
    int H, W, T;
    const float* b1; const float* b2;
    const float* res; float* out32; int ld32;
    uint16_t* out16; int ld16; int out16_dtype;
    const float* ln_g; const float* ln_b; int ln_C; int ln_win_shift;
};

template <int CP>
struct MlCfg {
    static constexpr int KB1 = CP / 64;
    static constexpr int A_BYTES = KB1 * 16384;
    static constexpr int H_BYTES = 2 * 32768;
    static constexpr int SROW = CP + 4;                          // fp32 staging row stride (floats)
    static constexpr int W_BYTES = ML_WSLOTS * ML_WSLOT;
    static constexpr int AUX = 512;                              // barriers + tmem slot
    static_assert(64 * SROW * 4 <= H_BYTES, "staging must fit in the H buffers");
    static_assert(CP * 128 <= ML_WSLOT, "W2 tile must fit a ring slot");
};

template <int CP>
__global__ void __launch_bounds__(ML_THREADS, 1)
mlp_tc5_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w1,
               const __grid_constant__ CUtensorMap map_w2, const MlpP p) {
    using Cfg = MlCfg<CP>;
    constexpr int KB1 = Cfg::KB1;
    extern __shared__ unsigned char ml_smem_raw[];
    const uint32_t raw = smem_u32(ml_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t sA = base, sH = sA + Cfg::A_BYTES, sW = sH + Cfg::H_BYTES, bars = sW + Cfg::W_BYTES;
    // barrier map
    const uint32_t a_full = bars, a_empty = bars + 8;
    auto w_full = [&](int s) { return bars + 16 + 8u * s; };
    auto w_empty = [&](int s) { return bars + 48 + 8u * s; };
    auto d1_full = [&](int b) { return bars + 80 + 8u * b; };
    auto d1_empty = [&](int b) { return bars + 96 + 8u * b; };
    auto h_full = [&](int b) { return bars + 112 + 8u * b; };
    auto h_empty = [&](int b) { return bars + 128 + 8u * b; };
    const uint32_t d2_full = bars + 144, d2_empty = bars + 152, tmem_slot = bars + 160;
    float* sb1 = reinterpret_cast<float*>(ml_smem_raw + (bars + Cfg::AUX - raw));   // [hid_p]
    float* sb2 = sb1 + p.hid_p;                                                     // [CP]
    float* sg = sb2 + CP;                                                            // [CP] LayerNorm gamma (0 beyond ln_C)
    float* sbt = sg + CP;                                                            // [CP] LayerNorm beta

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NC = p.NC;

    if (threadIdx.x == 0) {
        mbar_init(a_full, 1); mbar_init(a_empty, 1);
        for (int s = 0; s < ML_WSLOTS; ++s) { mbar_init(w_full(s), 1); mbar_init(w_empty(s), 1); }
        for (int b = 0; b < 2; ++b) {
            mbar_init(d1_full(b), 1); mbar_init(d1_empty(b), ML_EPI_WARPS);
            mbar_init(h_full(b), ML_EPI_WARPS); mbar_init(h_empty(b), 1);
        }
        mbar_init(d2_full, 1); mbar_init(d2_empty, ML_EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w1) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w2) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < p.hid_p; i += ML_THREADS) sb1[i] = p.b1[i];
    for (int i = threadIdx.x; i < CP; i += ML_THREADS) {
        sb2[i] = p.b2[i];
        const bool in = p.ln_g != nullptr && i < p.ln_C;
        sg[i] = in ? p.ln_g[i] : 0.f;
        sbt[i] = in ? p.ln_b[i] : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(ml_smem_raw + (tmem_slot - raw));
    const uint32_t tD1 = tmem_base, tD2 = tmem_base + 256;

    if (warp == 0) {
        // ===================================== TMA producer =====================================
        if (lane == 0) {
            int ws = 0, wph = 0, tc = 0;
            auto w_slot = [&](uint32_t bytes) {
                mbar_wait(w_empty(ws), wph ^ 1);
                mbar_expect_tx(w_full(ws), bytes);
                return sW + ws * ML_WSLOT;
            };
            auto w_next = [&]() { if (++ws == ML_WSLOTS) { ws = 0; wph ^= 1; } };
            auto load_w1 = [&](int c) {
                for (int kb = 0; kb < KB1; ++kb) {
                    const uint32_t dst = w_slot(16384);
                    tma_load_2d(dst, &map_w1, w_full(ws), kb * 64, c * 128);
                    w_next();
                }
            };
            auto load_w2 = [&](int c) {
                for (int k2 = 0; k2 < 2; ++k2) {
                    const uint32_t dst = w_slot(CP * 128);
                    tma_load_2d(dst, &map_w2, w_full(ws), (c * 2 + k2) * 64, 0);
                    w_next();
                }
            };
            for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++tc) {
                mbar_wait(a_empty, (tc & 1) ^ 1);
                mbar_expect_tx(a_full, Cfg::A_BYTES);
                for (int kb = 0; kb < KB1; ++kb) tma_load_2d(sA + kb * 16384, &map_a, a_full, kb * 64, tile * 128);
                load_w1(0);
                for (int c = 0; c < NC; ++c) {
                    if (c + 1 < NC) load_w1(c + 1);
                    load_w2(c);
                }
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer =====================================
        if (lane == 0) {
            const uint32_t idesc1 = umma_idesc(1, 128, 128), idesc2 = umma_idesc(1, 128, CP);
            int ws = 0, wph = 0, tc = 0;
            int use_d1[2] = {0, 0}, use_h[2] = {0, 0};
            auto w_next = [&]() { if (++ws == ML_WSLOTS) { ws = 0; wph ^= 1; } };
            auto fc1 = [&](int c) {
                const int b = c & 1;
                mbar_wait(d1_empty(b), (use_d1[b] & 1) ^ 1);
                tc_fence_after();
                for (int kb = 0; kb < KB1; ++kb) {
                    mbar_wait(w_full(ws), wph);
                    tc_fence_after();
                    const uint64_t da = umma_desc_sw128(sA + kb * 16384), db = umma_desc_sw128(sW + ws * ML_WSLOT);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        tc_mma_f16(tD1 + b * 128, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc1, (kb | k) != 0 ? 1u : 0u);
                    tc_commit(w_empty(ws));
                    w_next();
                }
                tc_commit(d1_full(b));
                use_d1[b]++;
            };
            auto fc2 = [&](int c) {
                const int b = c & 1;
                mbar_wait(h_full(b), use_h[b] & 1);
                tc_fence_after();
                if (c == 0) { mbar_wait(d2_empty, (tc & 1) ^ 1); tc_fence_after(); }
                for (int k2 = 0; k2 < 2; ++k2) {
                    mbar_wait(w_full(ws), wph);
                    tc_fence_after();
                    const uint64_t da = umma_desc_sw128(sH + b * 32768 + k2 * 16384), db = umma_desc_sw128(sW + ws * ML_WSLOT);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        tc_mma_f16(tD2, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc2, (c | k2 | k) != 0 ? 1u : 0u);
                    tc_commit(w_empty(ws));
                    w_next();
                }
                tc_commit(h_empty(b));
                use_h[b]++;
            };
            for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++tc) {
                mbar_wait(a_full, tc & 1);
                tc_fence_after();
                fc1(0);
                if (NC == 1) tc_commit(a_empty);
                for (int c = 0; c < NC; ++c) {
                    if (c + 1 < NC) {
                        fc1(c + 1);
                        if (c + 2 == NC) tc_commit(a_empty);        // last fc1 of the tile issued: A may be refilled
                    }
                    fc2(c);
                }
                tc_commit(d2_full);
            }
        }
    } else {
        // ===================================== epilogue warps =====================================
        const int ew = warp - 2, lg = warp & 3, q = ew >> 2;          // q: 32-column quarter of a 128-col chunk
        constexpr int NP = CP / 64;
        float* stg = reinterpret_cast<float*>(ml_smem_raw + (sH - raw));   // [64][SROW] fp32, aliases H
        const bool has_ln = p.ln_g != nullptr;
        const float inv_c = 1.f / (float)(p.ln_C > 0 ? p.ln_C : 1);

        // this warp finishes rows  tile*128 + half*64 + ew*RPW + i  (i < RPW) in half `half`
        auto row_of = [&](int tile_, int j) { return tile_ * 128 + (j / ML_RPW) * 64 + ew * ML_RPW + (j % ML_RPW); };
        float2 resv[2 * ML_RPW][NP];
#pragma unroll
        for (int j = 0; j < 2 * ML_RPW; ++j) {
            const int m = row_of(blockIdx.x, j);
            const bool ok = (int)blockIdx.x < p.m_tiles && m < p.M;
#pragma unroll
            for (int k = 0; k < NP; ++k)
                resv[j][k] = ok ? __ldg(reinterpret_cast<const float2*>(p.res + (size_t)m * p.ld32 + 64 * k + 2 * lane))
                                : make_float2(0.f, 0.f);
        }
        int n_d1[2] = {0, 0}, tc = 0;
        for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++tc) {
            // ---------------- GELU stage: D1 chunk -> bf16 operand tile H[c&1] ----------------
            for (int c = 0; c < NC; ++c) {
                const int b = c & 1;
                mbar_wait(d1_full(b), n_d1[b] & 1);
                tc_fence_after();
                // this warp: rows of lane group lg, hidden columns [q*W32*32, (q+1)*W32*32) of the chunk
                constexpr int W32 = 4 / ML_QN;                           // 32-column blocks per warp
                const int r = lg * 32 + lane;
#pragma unroll
                for (int sbk = 0; sbk < W32; ++sbk) {
                    const int col0 = (q * W32 + sbk) * 32;               // column offset inside the 128-col chunk
                    uint32_t v[32];
                    tc_ld32(tD1 + ((uint32_t)(lg * 32) << 16) + (uint32_t)(b * 128 + col0), v);
                    if (sbk == W32 - 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(d1_empty(b));         // D1[b] may be overwritten
                    }
                    if (sbk == 0) mbar_wait(h_empty(b), (n_d1[b] & 1) ^ 1);   // fc2 of the previous user of H[b] retired
                    const float4* bp = reinterpret_cast<const float4*>(sb1 + c * 128 + col0);
                    // K-major, 128B-swizzled operand layout: row r, 16 B chunk j -> r*128 + ((j ^ (r & 7)) << 4)
                    unsigned char* hrow = ml_smem_raw + (sH - raw) + b * 32768 + (col0 >> 6) * 16384 + r * 128;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 b0 = bp[2 * j], b1v = bp[2 * j + 1];
                        const uint32_t p0 = packf<SRK_BF16>(gelu_erf(__uint_as_float(v[8 * j]) + b0.x), gelu_erf(__uint_as_float(v[8 * j + 1]) + b0.y));
                        const uint32_t p1 = packf<SRK_BF16>(gelu_erf(__uint_as_float(v[8 * j + 2]) + b0.z), gelu_erf(__uint_as_float(v[8 * j + 3]) + b0.w));
                        const uint32_t p2 = packf<SRK_BF16>(gelu_erf(__uint_as_float(v[8 * j + 4]) + b1v.x), gelu_erf(__uint_as_float(v[8 * j + 5]) + b1v.y));
                        const uint32_t p3 = packf<SRK_BF16>(gelu_erf(__uint_as_float(v[8 * j + 6]) + b1v.z), gelu_erf(__uint_as_float(v[8 * j + 7]) + b1v.w));
                        const int chunk = ((col0 >> 5) & 1) * 4 + j;
                        *reinterpret_cast<uint4*>(hrow + ((chunk ^ (r & 7)) << 4)) = make_uint4(p0, p1, p2, p3);
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to UMMA
                __syncwarp();
                if (lane == 0) mbar_arrive(h_full(b));
                n_d1[b]++;
            }
            // ---------------- final stage: D2 -> x', LayerNorm / cast ----------------
            mbar_wait(d2_full, tc & 1);
            tc_fence_after();
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                if ((lg >> 1) == half) {
                    float* srow = stg + (size_t)((lg & 1) * 32 + lane) * Cfg::SROW;
#pragma unroll
                    for (int jj = 0; jj < CP / (16 * ML_QN); ++jj) {
                        const int c16 = q + ML_QN * jj;
                        uint32_t v[16];
                        asm volatile(
                            "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                            : "r"(tD2 + ((uint32_t)(lg * 32) << 16) + (uint32_t)(c16 * 16)));
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int j = 0; j < 16; j += 4)
                            *reinterpret_cast<uint4*>(srow + c16 * 16 + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(d2_empty);                // this warp's part of D2 is drained
                }
                asm volatile("bar.sync 5, %0;" ::"n"(32 * ML_EPI_WARPS) : "memory");   // staging complete
#pragma unroll
                for (int i = 0; i < ML_RPW; ++i) {
                    const int j = half * ML_RPW + i;
                    const int m = row_of(tile, j);
                    const bool valid = m < p.M;
                    const float* srow = stg + (size_t)(ew * ML_RPW + i) * Cfg::SROW;
                    float2 v[NP];
#pragma unroll
                    for (int k = 0; k < NP; ++k) {
                        v[k] = *reinterpret_cast<const float2*>(srow + 64 * k + 2 * lane);
                        const float2 b2v = *reinterpret_cast<const float2*>(sb2 + 64 * k + 2 * lane);
                        v[k].x += b2v.x + resv[j][k].x;
                        v[k].y += b2v.y + resv[j][k].y;
                    }
                    {   // request row j of the next tile into the same registers
                        const int tn = tile + gridDim.x;
                        const int mn = row_of(tn, j);
                        const bool ok = tn < p.m_tiles && mn < p.M;
#pragma unroll
                        for (int k = 0; k < NP; ++k)
                            resv[j][k] = ok ? __ldg(reinterpret_cast<const float2*>(p.res + (size_t)mn * p.ld32 + 64 * k + 2 * lane))
                                            : make_float2(0.f, 0.f);
                    }
                    if (valid) {
                        float* oo = p.out32 + (size_t)m * p.ld32 + 2 * lane;
#pragma unroll
                        for (int k = 0; k < NP; ++k) *reinterpret_cast<float2*>(oo + 64 * k) = v[k];
                    }
                    if (has_ln) {
                        float sm = 0.f;
#pragma unroll
                        for (int k = 0; k < NP; ++k) sm += v[k].x + v[k].y;
                        const float mean = warp_sum(sm) * inv_c;
                        float qq = 0.f;
#pragma unroll
                        for (int k = 0; k < NP; ++k)
                            if (64 * k + 2 * lane < p.ln_C) { const float a0 = v[k].x - mean, a1 = v[k].y - mean; qq += a0 * a0 + a1 * a1; }
                        const float rstd = rsqrtf(warp_sum(qq) * inv_c + 1e-5f);
                        int r16 = m;
                        if (valid && p.ln_win_shift >= 0) {
                            const int bi = m / p.T;
                            r16 = bi * p.T + token_to_win_pos(m - bi * p.T, p.H, p.W, p.ln_win_shift);
                        }
                        if (valid) {
                            uint16_t* o16 = p.out16 + (size_t)r16 * p.ld16 + 2 * lane;
#pragma unroll
                            for (int k = 0; k < NP; ++k) {
                                const float2 gg = *reinterpret_cast<const float2*>(sg + 64 * k + 2 * lane);
                                const float2 bb = *reinterpret_cast<const float2*>(sbt + 64 * k + 2 * lane);
                                const float y0 = (v[k].x - mean) * rstd * gg.x + bb.x;     // gamma = beta = 0 on pad columns
                                const float y1 = (v[k].y - mean) * rstd * gg.y + bb.y;
                                *reinterpret_cast<uint32_t*>(o16 + 64 * k) = pack2(y0, y1, p.out16_dtype);
                            }
                        }
                    } else if (p.out16 && valid) {
                        uint16_t* o16 = p.out16 + (size_t)m * p.ld16 + 2 * lane;
#pragma unroll
                        for (int k = 0; k < NP; ++k) *reinterpret_cast<uint32_t*>(o16 + 64 * k) = pack2(v[k].x, v[k].y, p.out16_dtype);
                    }
                }
                asm volatile("bar.sync 5, %0;" ::"n"(32 * ML_EPI_WARPS) : "memory");   // staging free
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

template <int CP>
static int launch_mlp(const srk_mlp_args* a, cudaStream_t st) {
    using Cfg = MlCfg<CP>;
    MlpP p{};
    p.M = a->M; p.C = a->C; p.hid_p = a->hid_p; p.NC = a->hid_p / 128; p.m_tiles = ceil_div(a->M, 128);
    p.H = a->H; p.W = a->W; p.T = (a->H > 0 && a->W > 0) ? a->H * a->W : 0;
    p.b1 = a->b1; p.b2 = a->b2; p.res = a->res; p.out32 = a->out32; p.ld32 = a->ld32;
    p.out16 = (uint16_t*)a->out16; p.ld16 = a->ld16; p.out16_dtype = a->out16_dtype;
    p.ln_g = a->ln_g; p.ln_b = a->ln_b; p.ln_C = a->ln_C; p.ln_win_shift = a->ln_win_shift;
    CUtensorMap ma, mw1, mw2;
    {
        cuuint64_t dims[2] = {(cuuint64_t)CP, (cuuint64_t)a->M};
        cuuint64_t strides[1] = {(cuuint64_t)a->lda * 2};
        cuuint32_t box[2] = {64, 128};
        if (int rc = encode_map(&ma, SRK_BF16, 2, a->A, dims, strides, box)) return rc;
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)CP, (cuuint64_t)a->hid_p};
        cuuint64_t strides[1] = {(cuuint64_t)CP * 2};
        cuuint32_t box[2] = {64, 128};
        if (int rc = encode_map(&mw1, SRK_BF16, 2, a->W1, dims, strides, box)) return rc;
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)a->hid_p, (cuuint64_t)CP};
        cuuint64_t strides[1] = {(cuuint64_t)a->hid_p * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)CP};
        if (int rc = encode_map(&mw2, SRK_BF16, 2, a->W2, dims, strides, box)) return rc;
    }
    const size_t smem = (size_t)Cfg::A_BYTES + Cfg::H_BYTES + Cfg::W_BYTES + Cfg::AUX + (size_t)(a->hid_p + 3 * CP) * 4 + 1024;
    SRK_REQUIRE(smem <= (size_t)ML_SMEM_TOTAL, "mlp: hidden dim %d needs too much shared memory", a->hid_p);
    static bool attr[64] = {};
    if (first_use_on_device(attr)) SRK_CUDA(cudaFuncSetAttribute(mlp_tc5_kernel<CP>, cudaFuncAttributeMaxDynamicSharedMemorySize, ML_SMEM_TOTAL));
    const int grid = p.m_tiles < num_sms() ? p.m_tiles : num_sms();
    mlp_tc5_kernel<CP><<<grid, ML_THREADS, smem, st>>>(ma, mw1, mw2, p);
    SRK_LAUNCH_CHECK("mlp_tc5_kernel");
    return 0;
}

}  // namespace srk

using namespace srk;

extern "C" int srk_mlp(const srk_mlp_args* a, void* stream) {
    SRK_REQUIRE(a && a->A && a->W1 && a->W2 && a->b1 && a->b2 && a->res && a->out32, "mlp: null pointer");
    SRK_REQUIRE(a->M > 0 && a->Cp % 64 == 0 && a->Cp >= 64 && a->Cp <= 192, "mlp: Cp must be 64, 128 or 192");
    SRK_REQUIRE(a->hid_p % 128 == 0 && a->hid_p >= 128, "mlp: hidden dim must be padded to a multiple of 128");
    SRK_REQUIRE(a->lda >= a->Cp && a->lda % 8 == 0 && a->ld32 >= a->Cp && a->ld32 % 4 == 0, "mlp: bad leading dims");
    SRK_REQUIRE(!a->out16 || (a->ld16 >= a->Cp && a->ld16 % 8 == 0), "mlp: bad ld16");
    if (a->ln_g) {
        SRK_REQUIRE(a->ln_b && a->out16 && a->ln_C > 0 && a->ln_C <= a->Cp, "mlp: bad LayerNorm arguments");
        if (a->ln_win_shift >= 0)
            SRK_REQUIRE(a->H > 0 && a->W > 0 && a->H % 8 == 0 && a->W % 8 == 0 && a->M % (a->H * a->W) == 0 &&
                        (a->ln_win_shift == 0 || a->ln_win_shift == 4), "mlp: bad window geometry");
    }
    if (srk_get_engine() != SRK_ENGINE_TCGEN05)
        return fail(SRK_ERR_UNSUPPORTED, "mlp: the fused MLP kernel exists for the tcgen05 engine only");
    ProfScope ps(SRK_PROF_GEMM, stream);
    cudaStream_t st = (cudaStream_t)stream;
    switch (a->Cp) {
        case 64: return launch_mlp<64>(a, st);
        case 128: return launch_mlp<128>(a, st);
        default: return launch_mlp<192>(a, st);
    }
}
