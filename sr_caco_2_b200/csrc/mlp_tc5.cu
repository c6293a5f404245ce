// Fused Swin MLP:  x' = x + fc2(GELU(fc1(A) + b1)) + b2,  then LayerNorm (or a plain cast) of x'
// for the next GEMM -- ONE persistent tcgen05 kernel, the 128 x hidden activation never leaves
// the SM.  Restates Mlp.forward (dlib/models/network_swinir.py:39-45) + the residual add and the
// next norm1 (:335, :293) of SwinTransformerBlock.forward.
//
// Per 128-token tile the hidden dimension is processed in chunks of 64 columns, c = 0 .. NC-1:
//   warp 0       TMA: the A tile (LN2 output, bf16) once per tile; W1 / W2 chunk tiles stream
//                through a 3-slot ring in exactly the order the MMA warp consumes them.
//   warp 1       MMA: fc1(c):  D1[c&1] (TMEM, 64 cols)   = A . W1[64c.., :]^T
//                     fc2(c):  D2[t&1] (TMEM, CP cols)  += H[c%3] . W2[:, 64c..]^T
//                issued as fc1(0) fc1(1) | fc1(2) fc2(0) | fc1(3) fc2(1) | ... so the tensor pipe
//                always has the fc1 of a later chunk queued while a chunk is in the GELU warps.
//   warps 4..11  GELU warps, two sets of 4 (one warp per TMEM lane group) on alternate chunks:
//                tcgen05.ld D1 -> + b1 -> exact-erf GELU (packed fp32x2, one MUFU per element) -> bf16 ->
//                written into shared memory in the 128B-swizzled K-major operand layout fc2 reads.
//   warps 12..19 final warps: D2 -> padded fp32 staging (64 rows at a time) -> + b2 + residual
//                (fp32 stream, register-prefetched half a tile ahead) -> x' store + LayerNorm /
//                cast store, all as coalesced row segments.  D2 is double buffered, so the final
//                stage of tile t runs under the GELU / MMA work of tile t+1.
#include "tc5_ptx.cuh"

namespace srk {

#ifdef SRK_MLP_TRACE
// experiment builds only (scripts/micro/trace_mlp.py): per-role clock64 stamps of CTA 0
__device__ long long g_ml_trace[4 * 64 * 8];
#define ML_TR(role, idx, ev) do { if (blockIdx.x == 0 && lane == 0 && (idx) < 64) g_ml_trace[((role) * 64 + (idx)) * 8 + (ev)] = clock64(); } while (0)
#else
#define ML_TR(role, idx, ev) do { } while (0)
#endif

constexpr int ML_CH = 64;                  // hidden columns per chunk
constexpr int ML_G_WARPS = 8;              // GELU warps
#ifndef SRK_ML_F_WARPS
#define SRK_ML_F_WARPS 8
#endif
constexpr int ML_F_WARPS = SRK_ML_F_WARPS;  // final-stage warps (1 or 2 per TMEM lane group)
constexpr int ML_THREADS = 128 + 32 * (ML_G_WARPS + ML_F_WARPS);
constexpr int ML_NH = 3;                   // H (GELU output) operand stages
constexpr int ML_NW = 3;                   // weight ring slots
constexpr int ML_RPW = 64 / ML_F_WARPS;    // rows per final warp in each 64-row half
constexpr int ML_RB = 2;                   // row PAIRS per final-stage step (a warp finishes two rows at a time)
constexpr int ML_NBK = 2;                  // residual banks of RB row pairs in flight (prefetch distance 2 * NBK * RB rows)
constexpr int ML_SMEM_TOTAL = 227 * 1024;

struct MlpP {
    int M, C, hid_p, NC, m_tiles;
    int H, W, T;
    const float* b1; const float* b2;
    const float* res; float* out32; int ld32;
    uint16_t* out16; int ld16; int out16_dtype;
    const float* ln_g; const float* ln_b; int ln_C; int ln_win_shift; int ln_pad_one;
};

template <int CP>
struct MlCfg {
    static constexpr int KB1 = CP / 64;
    static constexpr int A_BYTES = KB1 * 16384;
    static constexpr int H_BYTES = ML_NH * 16384;
    static constexpr int WSLOT = CP * 128;                       // one W1 chunk (KB1 boxes of 64 x 64) or one W2 chunk (CP x 64)
    static constexpr int W_BYTES = ML_NW * WSLOT;
    static constexpr int SROW = CP + 4;                          // fp32 staging row stride (floats)
    static constexpr int STG_BYTES = 64 * SROW * 4;
    static constexpr int AUX = 256;                              // barriers + tmem slot
    static constexpr int TMEM_D2 = 128;                          // D1: 2 x 64 columns, D2: 2 x CP columns
    static_assert(128 + 2 * CP <= 512, "TMEM budget");
};

template <int CP, bool LN>
__global__ void __launch_bounds__(ML_THREADS, 1)
mlp_tc5_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w1,
               const __grid_constant__ CUtensorMap map_w2, const MlpP p) {
    using Cfg = MlCfg<CP>;
    constexpr int KB1 = Cfg::KB1;
    extern __shared__ unsigned char ml_smem_raw[];
    const uint32_t raw = smem_u32(ml_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t sA = base, sH = sA + Cfg::A_BYTES, sW = sH + Cfg::H_BYTES, sStg = sW + Cfg::W_BYTES,
                   bars = sStg + Cfg::STG_BYTES;
    // barrier map
    const uint32_t a_full = bars, a_empty = bars + 8;
    auto w_full = [&](int s) { return bars + 16 + 8u * s; };
    auto w_empty = [&](int s) { return bars + 40 + 8u * s; };
    auto d1_full = [&](int b) { return bars + 64 + 8u * b; };
    auto d1_empty = [&](int b) { return bars + 80 + 8u * b; };
    auto h_full = [&](int b) { return bars + 96 + 8u * b; };
    auto h_empty = [&](int b) { return bars + 120 + 8u * b; };
    auto d2_full = [&](int b) { return bars + 144 + 8u * b; };
    auto d2_empty = [&](int b) { return bars + 160 + 8u * b; };
    const uint32_t tmem_slot = bars + 176;
    float* sb1 = reinterpret_cast<float*>(ml_smem_raw + (bars + Cfg::AUX - raw));   // [hid_p] fc1 bias, HALVED (gelu2)
    float* sb2 = sb1 + p.hid_p;                                                     // [CP]
    float* sg = sb2 + CP;                                                            // [CP] LayerNorm gamma (0 beyond ln_C)
    float* sbt = sg + CP;                                                            // [CP] LayerNorm beta (pad: 0, or 1 in ln_C, ln_C+1)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NC = p.NC;

    if (threadIdx.x == 0) {
        mbar_init(a_full, 1); mbar_init(a_empty, 1);
        for (int s = 0; s < ML_NW; ++s) { mbar_init(w_full(s), 1); mbar_init(w_empty(s), 1); }
        for (int b = 0; b < 2; ++b) {
            mbar_init(d1_full(b), 1); mbar_init(d1_empty(b), ML_G_WARPS / 2);
            mbar_init(d2_full(b), 1); mbar_init(d2_empty(b), ML_F_WARPS);
        }
        for (int s = 0; s < ML_NH; ++s) { mbar_init(h_full(s), ML_G_WARPS / 2); mbar_init(h_empty(s), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w1) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w2) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < p.hid_p; i += ML_THREADS) sb1[i] = 0.5f * p.b1[i];
    for (int i = threadIdx.x; i < CP; i += ML_THREADS) {
        sb2[i] = p.b2[i];
        const bool in = p.ln_g != nullptr && i < p.ln_C;
        sg[i] = in ? p.ln_g[i] : 0.f;
        sbt[i] = in ? p.ln_b[i] : ((p.ln_pad_one && (i == p.ln_C || i == p.ln_C + 1)) ? 1.f : 0.f);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(ml_smem_raw + (tmem_slot - raw));
    const uint32_t tD1 = tmem_base, tD2 = tmem_base + Cfg::TMEM_D2;

    // 20 warps start with 96 registers each; the final warps keep their residual window in registers and need
    // more, TMA / MMA / the two idle warps almost none: setmaxnreg moves registers between warp groups of 4 (it is
    // the first instruction of each role branch, so that the register allocator sees one limit per region).
    // An increase can only take what THIS CTA released: 128 x (96 - 48) + 256 x (96 - 80) = 10240 >= 256 x (128 - 96).
    if (warp < 4) {
    if constexpr (ML_THREADS == 640) asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
    if (warp == 0) {
        // ===================================== TMA producer =====================================
        if (lane == 0) {
            int ws = 0, wph = 0, tc = 0;
            auto w_next = [&]() { if (++ws == ML_NW) { ws = 0; wph ^= 1; } };
            auto load_w1 = [&](int c) {
                mbar_wait(w_empty(ws), wph ^ 1);
                mbar_expect_tx(w_full(ws), Cfg::WSLOT);
#pragma unroll
                for (int kb = 0; kb < KB1; ++kb)
                    tma_load_2d(sW + ws * Cfg::WSLOT + kb * 8192, &map_w1, w_full(ws), kb * 64, c * ML_CH);
                w_next();
            };
            auto load_w2 = [&](int c) {
                mbar_wait(w_empty(ws), wph ^ 1);
                mbar_expect_tx(w_full(ws), Cfg::WSLOT);
                tma_load_2d(sW + ws * Cfg::WSLOT, &map_w2, w_full(ws), c * ML_CH, 0);
                w_next();
            };
            // The CTA's chunks form ONE sequence q = tile_iteration * NC + c that runs across tile boundaries: the
            // weights arrive in the MMA warp's issue order  W1(0) W1(1) | W1(q+2) W2(q) ...,  and the A tile of the next
            // tile iteration is requested right in front of its first W1 (its bytes were prefetched into L2 a tile earlier).
            const int n_my = (int)blockIdx.x < p.m_tiles ? (p.m_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
            const int Q = n_my * NC;
            auto load_a = [&](int ti) {
                const int tile = (int)blockIdx.x + ti * (int)gridDim.x;
                mbar_wait(a_empty, (ti & 1) ^ 1);
                mbar_expect_tx(a_full, Cfg::A_BYTES);
#pragma unroll
                for (int kb = 0; kb < KB1; ++kb) tma_load_2d(sA + kb * 16384, &map_a, a_full, kb * 64, tile * 128);
                if (ti + 1 < n_my) {
#pragma unroll
                    for (int kb = 0; kb < KB1; ++kb) tma_prefetch_2d(&map_a, kb * 64, (tile + (int)gridDim.x) * 128);
                }
            };
            int c1 = 0, t1 = 0;
            auto load_fc1 = [&](int) {
                if (c1 == 0) load_a(t1);
                load_w1(c1);
                if (++c1 == NC) { c1 = 0; ++t1; }
            };
            if (Q > 0) load_fc1(0);
            if (Q > 1) load_fc1(1);
            for (int q = 0, c2 = 0; q < Q; ++q) {
                if (q + 2 < Q) load_fc1(q + 2);
                load_w2(c2);
                if (++c2 == NC) c2 = 0;
            }
            (void)tc;
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer =====================================
        {   // the whole warp runs the loop (uniform control flow); one elected lane issues the MMAs and commits
            const uint32_t idesc1 = umma_idesc(1, 128, ML_CH), idesc2 = umma_idesc(1, 128, CP);
            int ws = 0, wph = 0, tc = 0;
            int d1s = 0, d1ph = 0, hs = 0, hph = 0;
            auto w_next = [&]() { if (++ws == ML_NW) { ws = 0; wph ^= 1; } };
            int trq = 0;
            auto fc1 = [&]() {
                mbar_wait(d1_empty(d1s), d1ph ^ 1);
                ML_TR(0, trq, 4);
                mbar_wait(w_full(ws), wph);
                ML_TR(0, trq, 5);
                tc_fence_after();
#pragma unroll
                for (int kb = 0; kb < KB1; ++kb) {
                    const uint64_t da = umma_desc_sw128(sA + kb * 16384), db = umma_desc_sw128(sW + ws * Cfg::WSLOT + kb * 8192);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            tc_mma_f16(tD1 + d1s * ML_CH, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc1, (kb | k) != 0 ? 1u : 0u);
                    }
                }
                if (elect_one()) { tc_commit(w_empty(ws)); tc_commit(d1_full(d1s)); }
                __syncwarp();
                w_next();
                if (++d1s == 2) { d1s = 0; d1ph ^= 1; }
            };
            const int n_my = (int)blockIdx.x < p.m_tiles ? (p.m_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
            const int Q = n_my * NC;
            // fc1 of chunk q (tile iteration q / NC): the first chunk of a tile waits for its A tile, the last one hands
            // the A buffer back to the producer
            // (chunk, tile iteration) of the fc1 and fc2 sequences are carried as counters: no divisions in this loop
            int c1 = 0, t1 = 0;
            auto fc1_q = [&](int) {
                if (c1 == 0) { mbar_wait(a_full, t1 & 1); tc_fence_after(); }
                fc1();
                if (c1 == NC - 1 && elect_one()) tc_commit(a_empty);
                __syncwarp();
                if (++c1 == NC) { c1 = 0; ++t1; }
            };
            if (Q > 0) fc1_q(0);
            if (Q > 1) fc1_q(1);
            for (int q = 0, c = 0, ti = 0; q < Q; ++q) {
                ML_TR(0, q, 0);
                trq = q;
                if (q + 2 < Q) fc1_q(q + 2);                      // two chunks ahead, also across the tile boundary
                ML_TR(0, q, 1);
                const uint32_t d2 = tD2 + (uint32_t)((ti & 1) * CP);
                // fc2(q)
                mbar_wait(h_full(hs), hph);
                if (c == 0) mbar_wait(d2_empty(ti & 1), ((ti >> 1) & 1) ^ 1);
                ML_TR(0, q, 2);
                mbar_wait(w_full(ws), wph);
                tc_fence_after();
                ML_TR(0, q, 3);
                const uint64_t da = umma_desc_sw128(sH + hs * 16384), db = umma_desc_sw128(sW + ws * Cfg::WSLOT);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        tc_mma_f16(d2, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc2, (c | k) != 0 ? 1u : 0u);
                    tc_commit(w_empty(ws));
                    tc_commit(h_empty(hs));
                    if (c == NC - 1) tc_commit(d2_full(ti & 1));
                }
                __syncwarp();
                w_next();
                if (++hs == ML_NH) { hs = 0; hph ^= 1; }
                if (++c == NC) { c = 0; ++ti; }
            }
            (void)tc;
        }
    }
    } else if (warp < 4 + ML_G_WARPS) {
        // ===================================== GELU warps =====================================
        if constexpr (ML_THREADS == 640) asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
        // Two sets of 4 warps (one per TMEM lane group) take alternate chunks of the CTA's chunk sequence
        // q = tile_iteration * NC + c: set s owns the chunks with q % 2 == s, i.e. always D1 stage s, so one set
        // drains / converts a chunk while the MMAs of the other set's chunk run.  A warp converts the 64 columns of
        // its 32 rows in four 16-column pieces; the tcgen05.ld of a piece is in flight while the previous one is
        // converted (the TMEM read port is shared by every warp of the SM: an exposed load costs hundreds of cycles).
        const int lg = warp & 3, set = (warp - 4) >> 2;
        const int r = lg * 32 + lane;                                // tile row of this thread
        const int n_my_tiles = (int)blockIdx.x < p.m_tiles ? (p.m_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
        const int n_chunks = n_my_tiles * NC;
        const uint32_t t_row = tD1 + ((uint32_t)(lg * 32) << 16) + (uint32_t)(set * ML_CH);
        auto convert = [&](const uint32_t (&v)[16], const float* bias16, unsigned char* hrow, int piece) {
            const float4* bp = reinterpret_cast<const float4*>(bias16);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const float4 b0 = bp[2 * j], b1v = bp[2 * j + 1];
                uint32_t a0, a1, pk[4];
                unpack64(gelu2(pack64(v[8 * j], v[8 * j + 1]), pack64(__float_as_uint(b0.x), __float_as_uint(b0.y))), a0, a1);
                pk[0] = packf<SRK_BF16>(__uint_as_float(a0), __uint_as_float(a1));
                unpack64(gelu2(pack64(v[8 * j + 2], v[8 * j + 3]), pack64(__float_as_uint(b0.z), __float_as_uint(b0.w))), a0, a1);
                pk[1] = packf<SRK_BF16>(__uint_as_float(a0), __uint_as_float(a1));
                unpack64(gelu2(pack64(v[8 * j + 4], v[8 * j + 5]), pack64(__float_as_uint(b1v.x), __float_as_uint(b1v.y))), a0, a1);
                pk[2] = packf<SRK_BF16>(__uint_as_float(a0), __uint_as_float(a1));
                unpack64(gelu2(pack64(v[8 * j + 6], v[8 * j + 7]), pack64(__float_as_uint(b1v.z), __float_as_uint(b1v.w))), a0, a1);
                pk[3] = packf<SRK_BF16>(__uint_as_float(a0), __uint_as_float(a1));
                // K-major, 128B-swizzled operand layout: row r, 16 B chunk j -> r*128 + ((j ^ (r & 7)) << 4)
                const int chunk = piece * 2 + j;
                *reinterpret_cast<uint4*>(hrow + ((chunk ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
        };
        for (int q = set, k = 0, c = set % NC; q < n_chunks; q += 2, ++k) {
            const int hs = q % ML_NH, hu = q / ML_NH;
            if (lg == 0) ML_TR(1, q, 0);
            mbar_wait(d1_full(set), k & 1);
            tc_fence_after();
            if (lg == 0) ML_TR(1, q, 1);
            uint32_t va[16], vb[16];
            tc_ld16_nowait(t_row, va);
            tc_wait_ld16(va);
            tc_ld16_nowait(t_row + 16, vb);
            mbar_wait(h_empty(hs), (hu & 1) ^ 1);                    // fc2 of the previous user of this H stage retired
            if (lg == 0) ML_TR(1, q, 4);
            const float* bias = sb1 + c * ML_CH;
            unsigned char* hrow = ml_smem_raw + (sH - raw) + hs * 16384 + r * 128;
            convert(va, bias, hrow, 0);
            if (lg == 0) ML_TR(1, q, 5);
            tc_wait_ld16(vb);
            tc_ld16_nowait(t_row + 32, va);
            if (lg == 0) ML_TR(1, q, 6);
            convert(vb, bias + 16, hrow, 1);
            if (lg == 0) ML_TR(1, q, 7);
            tc_wait_ld16(va);
            tc_ld16_nowait(t_row + 48, vb);
            convert(va, bias + 32, hrow, 2);
            tc_wait_ld16(vb);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(d1_empty(set));               // D1 stage may be overwritten
            if (lg == 0) ML_TR(1, q, 2);
            convert(vb, bias + 48, hrow, 3);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to UMMA
            __syncwarp();
            if (lane == 0) mbar_arrive(h_full(hs));
            if (lg == 0) ML_TR(1, q, 3);
            c += 2;                                                  // chunk index of this set's next chunk
            while (c >= NC) c -= NC;
        }
    } else {
        // ===================================== final warps =====================================
        if constexpr (ML_THREADS == 640) asm volatile("setmaxnreg.inc.sync.aligned.u32 128;");
        constexpr int QN = ML_F_WARPS / 4;                                   // final warps per TMEM lane group
        const int fw = warp - 4 - ML_G_WARPS, lg = warp & 3, qh = fw >> 2;   // qh: column slice drained in phase T
        constexpr int NP = CP / 64;
        constexpr int RB = ML_RB;                                            // rows whose LayerNorm butterflies are interleaved
        constexpr int NBK = ML_NBK;
        static_assert(ML_RPW % (2 * NBK * RB) == 0 && (CP / 32) % QN == 0, "final-stage tiling");
        float* stg = reinterpret_cast<float*>(ml_smem_raw + (sStg - raw));   // [64][SROW] fp32
        const float inv_c = 1.f / (float)(p.ln_C > 0 ? p.ln_C : 1);

        // This warp finishes rows  tile*128 + half*64 + fw*RPW + i  (i < RPW) of each half.  Its rows form one
        // sequence over (tile, half, i); position pos = (2 * tile_iteration + half) * RPW + i.
        auto row_at = [&](int pos) {
            const int th = pos / ML_RPW, i = pos - th * ML_RPW;
            const int tile_ = (int)blockIdx.x + (th >> 1) * (int)gridDim.x;
            return tile_ < p.m_tiles ? tile_ * 128 + (th & 1) * 64 + fw * ML_RPW + i : p.M;
        };
        // Phase R layout: a warp finishes TWO rows at a time, one per half-warp; lane hl = lane & 15 owns the four
        // consecutive columns 64 k + 4 hl (k < NP) of its row, so every access is a 16 B vector and a half-warp
        // covers 256 contiguous bytes of the row.
        const int hl = lane & 15, hh = lane >> 4;
        // (rows beyond the problem are clamped to the last row: an unconditional load keeps the request in flight in
        //  the slot's own registers -- a predicated load + select makes the compiler wait for it on the spot)
        auto load_res = [&](float4 (&dst)[NP], int m) {
            const float* rp = p.res + (size_t)(m < p.M ? m : p.M - 1) * p.ld32 + 4 * hl;
#pragma unroll
            for (int k = 0; k < NP; ++k)
                asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(dst[k].x), "=f"(dst[k].y), "=f"(dst[k].z), "=f"(dst[k].w) : "l"(rp + 64 * k));
        };
        // residual rows: a rolling window of NBK banks x RB row pairs, NBK * RB pairs ahead of the pair being finished
        float4 resv[NBK][RB][NP];
#pragma unroll
        for (int i = 0; i < NBK * RB; ++i) load_res(resv[i / RB][i % RB], row_at(2 * i + hh));
        // lane i (< RPW) describes row i of the half being finished: its GEMM row and the row of the 16-bit output
        // (my_m, my_r16) and the GEMM row 2 * NBK * RB positions further on (my_mn: the residual row to request)
        int my_m = 0, my_r16 = 0, my_mn = 0;
        bool colin[NP];                                                      // this lane's column group lies inside the LayerNorm width
#pragma unroll
        for (int k = 0; k < NP; ++k) colin[k] = 64 * k + 4 * hl < p.ln_C;
        const bool o_bf16 = p.out16_dtype == SRK_BF16;
        auto pack = [&](float a, float b) { return o_bf16 ? packf<SRK_BF16>(a, b) : packf<SRK_FP16>(a, b); };

        // one step: RB row pairs starting at row i0 of the staged half, residuals in bank `rs`
        auto step = [&](float4 (&rs)[RB][NP], int i0) {
            int m[RB], r16[RB];
            float4 v[RB][NP];
#pragma unroll
            for (int rr = 0; rr < RB; ++rr) {
                const int i = i0 + 2 * rr + hh;                              // this half-warp's row of the staged half
                m[rr] = __shfl_sync(0xffffffffu, my_m, i);
                r16[rr] = __shfl_sync(0xffffffffu, my_r16, i);
                const float* srow = stg + (size_t)(fw * ML_RPW + i) * Cfg::SROW + 4 * hl;
#pragma unroll
                for (int k = 0; k < NP; ++k) {
                    v[rr][k] = *reinterpret_cast<const float4*>(srow + 64 * k);
                    const float4 bb = *reinterpret_cast<const float4*>(sb2 + 64 * k + 4 * hl);
                    v[rr][k].x += bb.x + rs[rr][k].x; v[rr][k].y += bb.y + rs[rr][k].y;
                    v[rr][k].z += bb.z + rs[rr][k].z; v[rr][k].w += bb.w + rs[rr][k].w;
                }
                load_res(rs[rr], __shfl_sync(0xffffffffu, my_mn, i));        // the row NBK * RB pairs ahead, same slot
                if (m[rr] < p.M) {
                    float* oo = p.out32 + (size_t)m[rr] * p.ld32 + 4 * hl;
#pragma unroll
                    for (int k = 0; k < NP; ++k) *reinterpret_cast<float4*>(oo + 64 * k) = v[rr][k];
                }
            }
            if constexpr (LN) {
                float sm[RB], mean[RB], qq[RB];
#pragma unroll
                for (int rr = 0; rr < RB; ++rr) {
                    sm[rr] = 0.f;
#pragma unroll
                    for (int k = 0; k < NP; ++k) sm[rr] += (v[rr][k].x + v[rr][k].y) + (v[rr][k].z + v[rr][k].w);   // pad columns are exactly 0
                }
#pragma unroll
                for (int o = 8; o > 0; o >>= 1)
#pragma unroll
                    for (int rr = 0; rr < RB; ++rr) sm[rr] += __shfl_xor_sync(0xffffffffu, sm[rr], o);
#pragma unroll
                for (int rr = 0; rr < RB; ++rr) {
                    mean[rr] = sm[rr] * inv_c; qq[rr] = 0.f;
#pragma unroll
                    for (int k = 0; k < NP; ++k) {
                        v[rr][k].x -= mean[rr]; v[rr][k].y -= mean[rr]; v[rr][k].z -= mean[rr]; v[rr][k].w -= mean[rr];
                        const float q4 = (v[rr][k].x * v[rr][k].x + v[rr][k].y * v[rr][k].y) + (v[rr][k].z * v[rr][k].z + v[rr][k].w * v[rr][k].w);
                        qq[rr] += colin[k] ? q4 : 0.f;
                    }
                }
#pragma unroll
                for (int o = 8; o > 0; o >>= 1)
#pragma unroll
                    for (int rr = 0; rr < RB; ++rr) qq[rr] += __shfl_xor_sync(0xffffffffu, qq[rr], o);
#pragma unroll
                for (int rr = 0; rr < RB; ++rr) {
                    const float rstd = rsqrtf(qq[rr] * inv_c + 1e-5f);
                    if (m[rr] < p.M) {
                        uint16_t* o16 = p.out16 + (size_t)r16[rr] * p.ld16 + 4 * hl;
#pragma unroll
                        for (int k = 0; k < NP; ++k) {
                            const float4 gg = *reinterpret_cast<const float4*>(sg + 64 * k + 4 * hl);    // gamma = 0 on pad columns
                            const float4 bb = *reinterpret_cast<const float4*>(sbt + 64 * k + 4 * hl);
                            *reinterpret_cast<uint2*>(o16 + 64 * k) =
                                make_uint2(pack(v[rr][k].x * rstd * gg.x + bb.x, v[rr][k].y * rstd * gg.y + bb.y),
                                           pack(v[rr][k].z * rstd * gg.z + bb.z, v[rr][k].w * rstd * gg.w + bb.w));
                        }
                    }
                }
            } else {
#pragma unroll
                for (int rr = 0; rr < RB; ++rr)
                    if (p.out16 && m[rr] < p.M) {
                        uint16_t* o16 = p.out16 + (size_t)m[rr] * p.ld16 + 4 * hl;
#pragma unroll
                        for (int k = 0; k < NP; ++k)
                            *reinterpret_cast<uint2*>(o16 + 64 * k) = make_uint2(pack(v[rr][k].x, v[rr][k].y), pack(v[rr][k].z, v[rr][k].w));
                    }
            }
        };

        int tc = 0, pos = 0;
        for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++tc) {
            const uint32_t d2 = tD2 + (uint32_t)((tc & 1) * CP);
            if (fw == 0) ML_TR(2, tc, 0);
            mbar_wait(d2_full(tc & 1), (tc >> 1) & 1);
            tc_fence_after();
            if (fw == 0) ML_TR(2, tc, 1);
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                // ---- phase T: this half's two lane groups drain D2 into the staging tile (thread = row) ----
                if ((lg >> 1) == half) {
                    float* srow = stg + (size_t)((lg & 1) * 32 + lane) * Cfg::SROW;
                    const uint32_t t_row = d2 + ((uint32_t)(lg * 32) << 16);
#pragma unroll 1
                    for (int jj = 0; jj < CP / 16 / QN; ++jj) {              // x16 loads, not unrolled: the residual window owns the registers
                        const int c16 = qh * (CP / 16 / QN) + jj;
                        uint32_t v[16];
                        tc_ld16_nowait(t_row + (uint32_t)(c16 * 16), v);
                        tc_wait_ld16(v);
#pragma unroll
                        for (int j = 0; j < 16; j += 4)
                            *reinterpret_cast<uint4*>(srow + c16 * 16 + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(d2_empty(tc & 1));            // this warp's part of D2 is drained
                }
                {   // row bookkeeping of this half: one row per lane
                    my_m = row_at(pos + (lane < ML_RPW ? lane : 0));
                    my_mn = row_at(pos + 2 * NBK * RB + (lane < ML_RPW ? lane : 0));
                    my_r16 = my_m;
                    if (LN && my_m < p.M && p.ln_win_shift >= 0) {
                        const int bi = my_m / p.T;
                        my_r16 = bi * p.T + token_to_win_pos(my_m - bi * p.T, p.H, p.W, p.ln_win_shift);
                    }
                }
                asm volatile("bar.sync 1, %0;" ::"n"(32 * ML_F_WARPS) : "memory");   // staging complete
                // ---- phase R: RPW rows per warp, lane = column pair, RB rows per step ----
#pragma unroll 1
                for (int i0 = 0; i0 < ML_RPW; i0 += 2 * NBK * RB, pos += 2 * NBK * RB) {
#pragma unroll
                    for (int bk = 0; bk < NBK; ++bk) step(resv[bk], i0 + 2 * bk * RB);
                }
                asm volatile("bar.sync 1, %0;" ::"n"(32 * ML_F_WARPS) : "memory");   // staging free
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

template <int CP, bool LN>
static int launch_mlp(const srk_mlp_args* a, cudaStream_t st) {
    using Cfg = MlCfg<CP>;
    MlpP p{};
    p.M = a->M; p.C = a->C; p.hid_p = a->hid_p; p.NC = a->hid_p / ML_CH; p.m_tiles = ceil_div(a->M, 128);
    p.H = a->H; p.W = a->W; p.T = (a->H > 0 && a->W > 0) ? a->H * a->W : 0;
    p.b1 = a->b1; p.b2 = a->b2; p.res = a->res; p.out32 = a->out32; p.ld32 = a->ld32;
    p.out16 = (uint16_t*)a->out16; p.ld16 = a->ld16; p.out16_dtype = a->out16_dtype;
    p.ln_g = a->ln_g; p.ln_b = a->ln_b; p.ln_C = a->ln_C; p.ln_win_shift = a->ln_win_shift; p.ln_pad_one = a->ln_pad_one;
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
        cuuint32_t box[2] = {64, ML_CH};
        if (int rc = encode_map(&mw1, SRK_BF16, 2, a->W1, dims, strides, box)) return rc;
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)a->hid_p, (cuuint64_t)CP};
        cuuint64_t strides[1] = {(cuuint64_t)a->hid_p * 2};
        cuuint32_t box[2] = {ML_CH, (cuuint32_t)CP};
        if (int rc = encode_map(&mw2, SRK_BF16, 2, a->W2, dims, strides, box)) return rc;
    }
    const size_t smem = (size_t)Cfg::A_BYTES + Cfg::H_BYTES + Cfg::W_BYTES + Cfg::STG_BYTES + Cfg::AUX +
                        (size_t)(a->hid_p + 3 * CP) * 4 + 1024;
    SRK_REQUIRE(smem <= (size_t)ML_SMEM_TOTAL, "mlp: hidden dim %d needs too much shared memory", a->hid_p);
    static bool attr[64] = {};
    if (first_use_on_device(attr)) SRK_CUDA(cudaFuncSetAttribute(mlp_tc5_kernel<CP, LN>, cudaFuncAttributeMaxDynamicSharedMemorySize, ML_SMEM_TOTAL));
    const int grid = p.m_tiles < num_sms() ? p.m_tiles : num_sms();
    mlp_tc5_kernel<CP, LN><<<grid, ML_THREADS, smem, st>>>(ma, mw1, mw2, p);
    SRK_LAUNCH_CHECK("mlp_tc5_kernel");
    return 0;
}

}  // namespace srk

using namespace srk;

#ifdef SRK_MLP_TRACE
extern "C" int srk_debug_mlp_trace(long long* out) {
    return cudaMemcpyFromSymbol(out, g_ml_trace, sizeof(long long) * 4 * 64 * 8) == cudaSuccess ? 0 : -1;
}
#endif

extern "C" int srk_mlp(const srk_mlp_args* a, void* stream) {
    SRK_REQUIRE(a && a->A && a->W1 && a->W2 && a->b1 && a->b2 && a->res && a->out32, "mlp: null pointer");
    SRK_REQUIRE(a->M > 0 && a->Cp % 64 == 0 && a->Cp >= 64 && a->Cp <= 192, "mlp: Cp must be 64, 128 or 192");
    SRK_REQUIRE(a->hid_p % ML_CH == 0 && a->hid_p >= ML_CH, "mlp: hidden dim must be padded to a multiple of 64");
    SRK_REQUIRE(a->lda >= a->Cp && a->lda % 8 == 0 && a->ld32 >= a->Cp && a->ld32 % 4 == 0, "mlp: bad leading dims");
    SRK_REQUIRE(((uintptr_t)a->res & 15) == 0 && ((uintptr_t)a->out32 & 15) == 0 && ((uintptr_t)a->out16 & 7) == 0, "mlp: res / out32 must be 16 B aligned, out16 8 B aligned");
    SRK_REQUIRE(!a->out16 || (a->ld16 >= a->Cp && a->ld16 % 8 == 0), "mlp: bad ld16");
    if (a->ln_g) {
        SRK_REQUIRE(a->ln_b && a->out16 && a->ln_C > 0 && a->ln_C <= a->Cp && a->ln_C % 4 == 0, "mlp: bad LayerNorm arguments (ln_C must be a multiple of 4)");
        SRK_REQUIRE(!a->ln_pad_one || (a->ln_C % 2 == 0 && a->ln_C + 2 <= a->Cp), "mlp: ln_pad_one needs two pad columns");
        if (a->ln_win_shift >= 0)
            SRK_REQUIRE(a->H > 0 && a->W > 0 && a->H % 8 == 0 && a->W % 8 == 0 && a->M % (a->H * a->W) == 0 &&
                        (a->ln_win_shift == 0 || a->ln_win_shift == 4), "mlp: bad window geometry");
    }
    if (srk_get_engine() != SRK_ENGINE_TCGEN05)
        return fail(SRK_ERR_UNSUPPORTED, "mlp: the fused MLP kernel exists for the tcgen05 engine only");
    ProfScope ps(SRK_PROF_MLP, stream);
    cudaStream_t st = (cudaStream_t)stream;
    const bool ln = a->ln_g != nullptr;
    switch (a->Cp) {
        case 64: return ln ? launch_mlp<64, true>(a, st) : launch_mlp<64, false>(a, st);
        case 128: return ln ? launch_mlp<128, true>(a, st) : launch_mlp<128, false>(a, st);
        default: return ln ? launch_mlp<192, true>(a, st) : launch_mlp<192, false>(a, st);
    }
}
