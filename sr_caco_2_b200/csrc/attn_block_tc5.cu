// The attention half of a Swin block as ONE persistent tcgen05 kernel:
//
//     x' = x + proj(window_attention(qkv(LN1(x)))) + b_proj ;   a2 = LN2(x')
//
// per 128-token tile (two 8x8 windows, window-major rows under the block's cyclic shift).  q, k, v, the
// attention output and the proj accumulator never leave the SM: the kernel reads the LN1 rows (bf16, TMA)
// and the fp32 residual rows and writes x' (fp32) and LN2(x') (bf16) -- 2304 B per token instead of the
// 3072 B (+ 2 launches) of the qkv+attention kernel followed by the proj GEMM.
// Restates SwinTransformerBlock.forward (dlib/models/network_swinir.py:287-337, first half),
// WindowAttention.forward (:148-176) and calculate_mask (:260-285).
//
//   warp 0       TMA producer: the A tile (LN1 rows) once per tile; the weights stream through a ring of
//                six 12 KB slots in exactly the order the MMA warp consumes them: per head the 96 rows
//                [q | k | v] x 32 of the head-major packed qkv weight (3 k-blocks), and the proj weight
//                (3 k-blocks x 2 slots).  The qkv bias rides in the two pad K columns (fold_qkv_bias).
//   warp 1       MMA issuer.  Head job h:  Q[j % 3] (TMEM, 96 cols) = A . Wqkv[h]^T   (M128 N96  K192)
//                proj job:               PD (TMEM, 192 cols)      = AO . Wproj^T     (M128 N192 K192)
//                issued as  h0 h1 proj(t-1) h2 h3 h4 h5  per tile, so that the first heads of a tile are
//                in flight before the previous tile's proj (which has to wait for its last head).
//   warps 4..19  two groups of 8.  Group g runs the head units h = g, g+2, g+4 of every tile:
//                  drain Q (thread = row) into a bf16 staging tile [q | k | v] of 128 rows,
//                  8 x (window, 16-query strip) attention units (attention.cuh: mma.sync S and PV, softmax
//                  in registers, rel-pos bias + shift mask arithmetic), O written straight into the
//                  128B-swizzled K-major operand tile AO that the proj MMAs read.
//                After the first unit of tile t both groups together finish tile t-1: PD -> fp32 staging
//                (aliases the two bf16 staging tiles) -> + b_proj + residual (gathered through the
//                window_reverse + roll map) -> x' store, LayerNorm, bf16 store -- two rows per warp at a
//                time, 16 B per lane, like the final stage of mlp_tc5.cu.
#include "tc5_ptx.cuh"
#include "attention.cuh"

namespace srk {

#ifdef SRK_AB_TRACE
// experiment builds only (scripts/micro/trace_ab.py): per-role clock64 stamps of CTA 0
__device__ long long g_ab_trace[4 * 64 * 8];
#define AB_TR(role, idx, ev) do { if (blockIdx.x == 0 && lane == 0 && (idx) < 64) g_ab_trace[((role) * 64 + (idx)) * 8 + (ev)] = clock64(); } while (0)
#else
#define AB_TR(role, idx, ev) do { } while (0)
#endif

constexpr int AB_CP = 192;                  // padded embedding = heads * 32 (the only instantiation: 6 heads)
constexpr int AB_KB = AB_CP / 64;
constexpr int AB_NH = 6;
constexpr int AB_THREADS = 128 + 32 * 16;
constexpr int AB_WSLOT = 96 * 128;          // 96 weight rows x 64 k, bf16
constexpr int AB_NW = 6;
constexpr int AB_SROW16 = 96 * 2 + 16;      // staged [q | k | v] row stride (bytes): an odd multiple of 16 B
constexpr int AB_STG = 128 * AB_SROW16;     // one group's staging tile
constexpr int AB_SROW32 = AB_CP + 4;        // fp32 staging row stride (floats) of the final stage
constexpr int AB_A_OFF = 0;
constexpr int AB_W_OFF = AB_A_OFF + AB_KB * 16384;
constexpr int AB_AO_OFF = AB_W_OFF + AB_NW * AB_WSLOT;
constexpr int AB_STG_OFF = AB_AO_OFF + AB_KB * 16384;
constexpr int AB_BAR_OFF = AB_STG_OFF + 2 * AB_STG;
constexpr int AB_TAB = (AB_NH * 225 * 4 + 15) / 16 * 16;         // rel-pos tables, padded to 16 B
constexpr int AB_AUX = 256 + AB_TAB + 256 + AB_CP * 4 + 192;   // barriers, rel-pos tables, shift-mask labels, proj bias, window geometry
constexpr size_t AB_SMEM = (size_t)AB_BAR_OFF + AB_AUX;         // the dynamic shared memory window itself is 1024 B aligned
static_assert(64 * AB_SROW32 * 4 <= 2 * AB_STG, "the fp32 staging of the final stage aliases the two bf16 staging tiles");
static_assert(AB_SMEM <= 232448, "shared memory plan exceeds the 227 KB per-CTA limit");

struct AbP {
    int M, m_tiles, C, H, W, T, shift;
    float scale;
    const float* table;                       // (heads, 225)
    const float* b_proj;
    const float* res; float* out32; int ld32;
    uint16_t* out16; int ld16; int out16_dtype;
    const float* ln_g; const float* ln_b; int ln_C;
};

__global__ void __launch_bounds__(AB_THREADS, 1)
attn_block_tc5_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_wq,
                      const __grid_constant__ CUtensorMap map_wp, const AbP p) {
    extern __shared__ __align__(1024) unsigned char ab_raw[];
    const uint32_t base = smem_u32(ab_raw);
    if ((base & 1023u) != 0) __trap();                               // SW128 operand tiles need 1024 B alignment (no slack is allocated)
    unsigned char* sm = ab_raw;
    const uint32_t sA = base + AB_A_OFF, sW = base + AB_W_OFF, sAO = base + AB_AO_OFF, sStg = base + AB_STG_OFF,
                   bars = base + AB_BAR_OFF;
    const uint32_t a_full = bars, a_empty = bars + 8;
    auto w_full = [&](int s) { return bars + 16 + 8u * s; };
    auto w_empty = [&](int s) { return bars + 64 + 8u * s; };
    auto q_full = [&](int s) { return bars + 112 + 8u * s; };
    auto q_empty = [&](int s) { return bars + 136 + 8u * s; };
    const uint32_t ao_full = bars + 160, ao_empty = bars + 168, pd_full = bars + 176, pd_empty = bars + 184,
                   tmem_slot = bars + 192;
    float* stab = reinterpret_cast<float*>(sm + AB_BAR_OFF + 256);                    // [heads][225]
    unsigned char* slab = sm + AB_BAR_OFF + 256 + AB_TAB;                   // [group][window][64] region labels
    float* sbp = reinterpret_cast<float*>(slab + 256);                               // [192] proj bias
    int4* sgeo_all = reinterpret_cast<int4*>(slab + 256 + AB_CP * 4);                 // [group][tile iteration % 3][window]: see compute_geo

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_my = (int)blockIdx.x < p.m_tiles ? (p.m_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    if (threadIdx.x == 0) {
        mbar_init(a_full, 1); mbar_init(a_empty, 1);
        for (int s = 0; s < AB_NW; ++s) { mbar_init(w_full(s), 1); mbar_init(w_empty(s), 1); }
        for (int s = 0; s < 3; ++s) { mbar_init(q_full(s), 1); mbar_init(q_empty(s), 8); }
        mbar_init(ao_full, 8 * AB_NH); mbar_init(ao_empty, 1);
        mbar_init(pd_full, 1); mbar_init(pd_empty, 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wq) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wp) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < AB_NH * 225; i += AB_THREADS) stab[i] = p.table[i];
    for (int i = threadIdx.x; i < AB_CP; i += AB_THREADS) sbp[i] = p.b_proj[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(sm + AB_BAR_OFF + 192);
    const uint32_t tQ = tmem_base, tPD = tmem_base + 288;          // Q: 3 stages x 96 columns, PD: 192 columns

    if (warp < 4) {
    // 20 warps start with 96 registers; the 16 unit warps take what the TMA / MMA / idle warps release:
    // 128 x (96 - 56) = 5120 >= 512 x (104 - 96)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
        // ======================================= TMA producer =======================================
        if (lane == 0) {
            int ws = 0, wph = 0;
            auto slot = [&]() {
                mbar_wait(w_empty(ws), wph ^ 1);
                mbar_expect_tx(w_full(ws), AB_WSLOT);
                return sW + ws * AB_WSLOT;
            };
            auto w_next = [&]() { if (++ws == AB_NW) { ws = 0; wph ^= 1; } };
            auto load_head = [&](int h) {
                for (int kb = 0; kb < AB_KB; ++kb) { const uint32_t d = slot(); tma_load_2d(d, &map_wq, w_full(ws), kb * 64, h * 96); w_next(); }
            };
            auto load_proj = [&]() {
                for (int kb = 0; kb < AB_KB; ++kb)
                    for (int half = 0; half < 2; ++half) { const uint32_t d = slot(); tma_load_2d(d, &map_wp, w_full(ws), kb * 64, half * 96); w_next(); }
            };
            int it = 0;
            for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++it) {
                mbar_wait(a_empty, (it & 1) ^ 1);
                mbar_expect_tx(a_full, AB_KB * 16384);
                for (int kb = 0; kb < AB_KB; ++kb) tma_load_2d(sA + kb * 16384, &map_a, a_full, kb * 64, tile * 128);
                load_head(0); load_head(1);
                if (it > 0) load_proj();
                for (int h = 2; h < AB_NH; ++h) load_head(h);
            }
            if (n_my > 0) load_proj();
        }
    } else if (warp == 1) {
        // ======================================= MMA issuer =======================================
        if (lane == 0) {
            const uint32_t id_q = umma_idesc(1, 128, 96), id_p = umma_idesc(1, 128, AB_CP);
            int ws = 0, wph = 0, j = 0;
            auto w_next = [&]() { if (++ws == AB_NW) { ws = 0; wph ^= 1; } };
            auto head_job = [&]() {
                const int s = j % 3;
                mbar_wait(q_empty(s), ((j / 3) & 1) ^ 1);
                tc_fence_after();
                for (int kb = 0; kb < AB_KB; ++kb) {
                    mbar_wait(w_full(ws), wph);
                    tc_fence_after();
                    const uint64_t da = umma_desc_sw128(sA + kb * 16384), db = umma_desc_sw128(sW + ws * AB_WSLOT);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        tc_mma_f16(tQ + (uint32_t)(s * 96), da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), id_q, (kb | k) != 0 ? 1u : 0u);
                    tc_commit(w_empty(ws));
                    w_next();
                }
                tc_commit(q_full(s));
                ++j;
            };
            auto proj_job = [&](int tp) {
                mbar_wait(ao_full, tp & 1);                         // every head unit of the tile wrote its O
                mbar_wait(pd_empty, (tp & 1) ^ 1);                  // the previous tile's accumulator was drained
                tc_fence_after();
                for (int kb = 0; kb < AB_KB; ++kb) {
                    mbar_wait(w_full(ws), wph);
                    const int ws1 = ws + 1;                         // ws is even here: the two halves are adjacent slots
                    mbar_wait(w_full(ws1), wph);
                    tc_fence_after();
                    const uint64_t da = umma_desc_sw128(sAO + kb * 16384), db = umma_desc_sw128(sW + ws * AB_WSLOT);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        tc_mma_f16(tPD, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), id_p, (kb | k) != 0 ? 1u : 0u);
                    tc_commit(w_empty(ws)); w_next();
                    tc_commit(w_empty(ws)); w_next();
                }
                tc_commit(pd_full);
                tc_commit(ao_empty);
            };
            for (int it = 0; it < n_my; ++it) {
                mbar_wait(a_full, it & 1);
                tc_fence_after();
#pragma unroll 1
                for (int h = 0; h <= AB_NH; ++h) {
                    if (h == 2 || h == AB_NH) {                      // proj of the previous tile after the first two heads;
                        const int tp = h == 2 ? it - 1 : (it == n_my - 1 ? it : -1);   // the last tile's own proj at the end
                        if (tp >= 0) proj_job(tp);
                        if (h == AB_NH) break;
                    }
                    head_job();
                    if (h == AB_NH - 1) tc_commit(a_empty);         // the A tile may be refilled once the last head retires
                }
            }
        }
    }
    } else {
        // ======================================= unit warps =======================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
        const int uw = warp - 4, grp = uw >> 3, gw = uw & 7;          // group, warp inside the group
        const int lg = warp & 3;                                      // TMEM lane group
        const int hq = gw >> 2;                                       // drain: which 48 of the 96 accumulator columns
        const int window = gw >> 2, strip = gw & 3;                   // attention: (window, 16-query strip) of this warp
        unsigned char* stg = sm + AB_STG_OFF + grp * AB_STG;          // this group's bf16 staging tile
        const uint32_t stg_s = sStg + grp * AB_STG;
        unsigned char* lab = slab + grp * 128;
        const int nWy = p.H >> 3, wpr = p.W >> 3, nW = nWy * wpr;
        const int bar_g = 1 + grp;
        unsigned char* ao = sm + AB_AO_OFF;

        // ---- head unit: head h of tile iteration it (job index 6 it + h) ----
        // per tile, once per group: the shift-mask region labels of the tile's two windows (threads 0..127 of the group;
        // the group barrier inside the tile's first unit publishes them) and this warp's window flags
        // Window geometry (two integer divisions per window) is worked out ONCE per tile and group, by two threads, an
        // iteration ahead (three slots: tile it + 1 is written while tiles it and it - 1 are still read; the group
        // barriers of the unit that follows publish it); everybody else reads the 16 B record.
        //   x: first token row of the window's image, or -1 beyond the problem   y, z: window origin (rows, columns)
        //   w: 1 if the shift mask applies (last window row / column of a shifted block)
        int4* sgeo = sgeo_all + grp * 6;
        auto compute_geo = [&](int it, int tile) {
            const int tg = gw * 32 + lane;
            if (tg < 2) {
                const int wg = tile * 2 + tg;
                const int bi = wg / nW, win = wg - bi * nW;
                const int wi = win / wpr, wj = win - wi * wpr;
                sgeo[(it % 3) * 2 + tg] = make_int4((long long)wg * 64 < p.M ? bi * p.T : -1, wi << 3, wj << 3,
                                                    (p.shift > 0 && (wi == nWy - 1 || wj == wpr - 1)) ? 1 : 0);
            }
        };
        bool w_valid = false, w_masked = false;
        auto tile_setup = [&](int it) {
            const int tg = gw * 32 + lane;
            if (tg < 128) {
                const int4 g = sgeo[(it % 3) * 2 + (tg >> 6)];
                const int pos = tg & 63;
                lab[tg] = (unsigned char)(g.w ? region_label(g.y + (pos >> 3), p.H, p.shift) * 3 + region_label(g.z + (pos & 7), p.W, p.shift) : 0);
            }
            const int4 g2 = sgeo[(it % 3) * 2 + window];
            w_valid = g2.x >= 0;
            w_masked = g2.w != 0;
        };
        auto unit = [&](int it, int h) {
            const int j = it * AB_NH + h, s = j % 3;
            const int tri = it * 3 + (h >> 1);
            if (gw == 0) AB_TR(grp, tri, 0);
            mbar_wait(q_full(s), (j / 3) & 1);
            tc_fence_after();
            if (gw == 0) AB_TR(grp, tri, 1);
            {   // drain: thread = row, 48 columns -> 96 B of the staged row [q 32 | k 32 | v 32]
                const uint32_t t_row = tQ + ((uint32_t)(lg * 32) << 16) + (uint32_t)(s * 96 + hq * 48);
                unsigned char* srow = stg + (size_t)(lg * 32 + lane) * AB_SROW16 + hq * 96;
                uint32_t va[16], vb[16];
                tc_ld16_nowait(t_row, va);
                tc_wait_ld16(va);
                tc_ld16_nowait(t_row + 16, vb);
                auto put = [&](const uint32_t (&v)[16], int c) {
                    uint32_t pk[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) pk[e] = packf<SRK_BF16>(__uint_as_float(v[2 * e]), __uint_as_float(v[2 * e + 1]));
                    *reinterpret_cast<uint4*>(srow + c * 32) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    *reinterpret_cast<uint4*>(srow + c * 32 + 16) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                };
                put(va, 0);
                tc_wait_ld16(vb);
                tc_ld16_nowait(t_row + 32, va);
                put(vb, 1);
                tc_wait_ld16(va);
                put(va, 2);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(q_empty(s));                       // accumulator stage drained
            if (gw == 0) AB_TR(grp, tri, 2);
            asm volatile("bar.sync %0, 256;" ::"r"(bar_g) : "memory");   // the head is staged (+ labels)
            if (gw == 0) AB_TR(grp, tri, 3);
            if (it > 0) mbar_wait(ao_empty, (it - 1) & 1);                // the previous tile's proj MMAs have read AO
            if (gw == 0) AB_TR(grp, tri, 4);
            if (w_valid)
                attn_unit<32, true, unsigned char>(stg_s + window * 64 * AB_SROW16, stg + (size_t)window * 64 * AB_SROW16, AB_SROW16,
                                                   strip, 0, 32, 64, stab + h * 225, lab + window * 64, w_masked, p.scale, lane,
                                                   ao, window * 64, h);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes of AO -> visible to UMMA
            __syncwarp();
            if (lane == 0) mbar_arrive(ao_full);
            if (gw == 0) AB_TR(grp, tri, 5);
            asm volatile("bar.sync %0, 256;" ::"r"(bar_g) : "memory");   // staging (k, v) free for the group's next unit
            if (gw == 0) AB_TR(grp, tri, 6);
        };

        // ---- final stage, one QUARTER of a tile at a time: PD rows of one TMEM lane group -> x', LN2 ----
        // Group g finishes window g of the previous tile (rows 64 g .. 64 g + 63 = lane groups 2 g, 2 g + 1) in two
        // quarters of 32 rows, each between two of its head units, with its own staging tile and its own barrier: no
        // barrier spans the two groups, so one group's (latency bound) final stage runs under the other group's
        // attention arithmetic.  Phase T: the two warps of the quarter's lane group drain 96 accumulator columns each
        // into the fp32 staging tile (aliases the group's bf16 staging tile); phase R: all 8 warps, 4 rows each, two
        // rows at a time: half-warp hh owns a row, lane hl the columns 64 k + 4 hl.
        const int hl = lane & 15, hh = lane >> 4;
        float* stg32 = reinterpret_cast<float*>(stg);                 // [32][AB_SROW32] fp32
        const float inv_c = 1.f / (float)p.ln_C;
        bool colin[AB_KB];
#pragma unroll
        for (int k = 0; k < AB_KB; ++k) colin[k] = 64 * k + 4 * hl < p.ln_C;
        const bool o_bf16 = p.out16_dtype == SRK_BF16;
        auto pack = [&](float a, float b) { return o_bf16 ? packf<SRK_BF16>(a, b) : packf<SRK_FP16>(a, b); };
        // token row (fp32 stream / LN2 output) of window position pos of window `grp` of tile iteration it, or -1
        auto token_row = [&](int it, int pos) {
            const int4 g = sgeo[(it % 3) * 2 + grp];
            int y = g.y + p.shift + (pos >> 3), x = g.z + p.shift + (pos & 7);
            if (y >= p.H) y -= p.H;
            if (x >= p.W) x -= p.W;
            return g.x < 0 ? -1 : g.x + y * p.W + x;
        };
        // L2 prefetch of the residual rows of this group's window (issued a tile ahead: no registers held)
        auto prefetch_res = [&](int it) {
            const int tg = gw * 32 + lane;
            const int row = token_row(it, tg >> 2);
            if (row >= 0) {
                const float* rp = p.res + (size_t)row * p.ld32 + 32 * (tg & 3);
                asm volatile("prefetch.global.L2 [%0];" ::"l"(rp));
                if ((tg & 3) < 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + 128));
            }
        };
        auto quarter = [&](int tp, int qi) {
            // residual rows of this warp (L2 hits thanks to prefetch_res), requested before anything else
            float4 resv[2][AB_KB];
            int row[2];
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                row[rr] = token_row(tp, 32 * qi + gw * 4 + 2 * rr + hh);
                const float* rp = p.res + (size_t)(row[rr] >= 0 ? row[rr] : 0) * p.ld32 + 4 * hl;
#pragma unroll
                for (int k = 0; k < AB_KB; ++k)
                    asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                                 : "=f"(resv[rr][k].x), "=f"(resv[rr][k].y), "=f"(resv[rr][k].z), "=f"(resv[rr][k].w)
                                 : "l"(rp + 64 * k));
            }
            if (gw == 0) AB_TR(2 + grp, 2 * tp + qi, 0);
            // phase T: lane group 2 grp + qi = warps gw = 2 grp + qi (columns 0..95) and gw + 4 (columns 96..191)
            if ((gw & 3) == 2 * grp + qi) {
                mbar_wait(pd_full, tp & 1);
                tc_fence_after();
                const int hq2 = gw >> 2;
                float* srow = stg32 + (size_t)lane * AB_SROW32 + hq2 * 96;
                const uint32_t t_row = tPD + ((uint32_t)(lg * 32) << 16) + (uint32_t)(hq2 * 96);
                // three 32-column loads (one round trip each) instead of six 16-column ones: under the other group's
                // attention arithmetic a tcgen05.ld + wait costs ~250 cycles, and only two warps can drain a lane group
#pragma unroll 1
                for (int c = 0; c < 3; ++c) {
                    uint32_t v[32];
                    tc_ld32(t_row + 32 * c, v);
#pragma unroll
                    for (int e = 0; e < 32; e += 4) *reinterpret_cast<uint4*>(srow + 32 * c + e) = make_uint4(v[e], v[e + 1], v[e + 2], v[e + 3]);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(pd_empty);
            }
            // LayerNorm parameters of this lane's columns
            float4 gg[AB_KB], bt[AB_KB];
#pragma unroll
            for (int k = 0; k < AB_KB; ++k) {
                gg[k] = make_float4(0.f, 0.f, 0.f, 0.f); bt[k] = gg[k];                   // pad columns: 0
                if (colin[k]) {
                    gg[k] = __ldg(reinterpret_cast<const float4*>(p.ln_g + 64 * k + 4 * hl));
                    bt[k] = __ldg(reinterpret_cast<const float4*>(p.ln_b + 64 * k + 4 * hl));
                }
            }
            asm volatile("bar.sync %0, 256;" ::"r"(bar_g) : "memory");   // staging complete
            if (gw == 0) AB_TR(2 + grp, 2 * tp + qi, 1);
            // phase R
            float4 v[2][AB_KB];
            float sm_[2], mean[2], qq[2];
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const float* srow = stg32 + (size_t)(gw * 4 + 2 * rr + hh) * AB_SROW32 + 4 * hl;
                sm_[rr] = 0.f;
#pragma unroll
                for (int k = 0; k < AB_KB; ++k) {
                    v[rr][k] = *reinterpret_cast<const float4*>(srow + 64 * k);
                    const float4 bb = *reinterpret_cast<const float4*>(sbp + 64 * k + 4 * hl);
                    v[rr][k].x = (v[rr][k].x + bb.x) + resv[rr][k].x; v[rr][k].y = (v[rr][k].y + bb.y) + resv[rr][k].y;
                    v[rr][k].z = (v[rr][k].z + bb.z) + resv[rr][k].z; v[rr][k].w = (v[rr][k].w + bb.w) + resv[rr][k].w;
                    sm_[rr] += (v[rr][k].x + v[rr][k].y) + (v[rr][k].z + v[rr][k].w);      // pad columns are exactly 0
                }
                if (row[rr] >= 0) {
                    float* oo = p.out32 + (size_t)row[rr] * p.ld32 + 4 * hl;
#pragma unroll
                    for (int k = 0; k < AB_KB; ++k) *reinterpret_cast<float4*>(oo + 64 * k) = v[rr][k];
                }
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1)
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) sm_[rr] += __shfl_xor_sync(0xffffffffu, sm_[rr], o);
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                mean[rr] = sm_[rr] * inv_c; qq[rr] = 0.f;
#pragma unroll
                for (int k = 0; k < AB_KB; ++k) {
                    v[rr][k].x -= mean[rr]; v[rr][k].y -= mean[rr]; v[rr][k].z -= mean[rr]; v[rr][k].w -= mean[rr];
                    const float q4 = (v[rr][k].x * v[rr][k].x + v[rr][k].y * v[rr][k].y) + (v[rr][k].z * v[rr][k].z + v[rr][k].w * v[rr][k].w);
                    qq[rr] += colin[k] ? q4 : 0.f;
                }
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1)
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) qq[rr] += __shfl_xor_sync(0xffffffffu, qq[rr], o);
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const float rstd = rsqrtf(qq[rr] * inv_c + 1e-5f);
                if (row[rr] >= 0) {
                    uint16_t* o16 = p.out16 + (size_t)row[rr] * p.ld16 + 4 * hl;
#pragma unroll
                    for (int k = 0; k < AB_KB; ++k)
                        *reinterpret_cast<uint2*>(o16 + 64 * k) =
                            make_uint2(pack(v[rr][k].x * rstd * gg[k].x + bt[k].x, v[rr][k].y * rstd * gg[k].y + bt[k].y),
                                       pack(v[rr][k].z * rstd * gg[k].z + bt[k].z, v[rr][k].w * rstd * gg[k].w + bt[k].w));
                }
            }
            asm volatile("bar.sync %0, 256;" ::"r"(bar_g) : "memory");   // staging free for the next unit's drain
            if (gw == 0) AB_TR(2 + grp, 2 * tp + qi, 2);
        };

        // Per tile and group: three head units with the two final-stage quarters of the PREVIOUS tile between them; the
        // groups use different slots (0: u q u q u, 1: u u q u q) so that their final stages do not coincide.
        int tile = blockIdx.x;
        compute_geo(0, tile);
        asm volatile("bar.sync %0, 256;" ::"r"(bar_g) : "memory");
#pragma unroll 1
        for (int it = 0; it <= n_my; ++it, tile += gridDim.x) {
            const bool live = it < n_my, fin = it > 0;
            if (live) {
                tile_setup(it);
                compute_geo(it + 1, tile + gridDim.x);
                unit(it, grp);
                unit(it, grp + 2);
                if (grp == 1) unit(it, grp + 4);
            }
            if (fin) quarter(it - 1, 0);
            if (live && grp == 0) unit(it, grp + 4);
            if (fin) quarter(it - 1, 1);
            if (live) prefetch_res(it);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

}  // namespace srk

using namespace srk;

#ifdef SRK_AB_TRACE
extern "C" int srk_debug_ab_trace(long long* out) {
    return cudaMemcpyFromSymbol(out, g_ab_trace, sizeof(long long) * 4 * 64 * 8) == cudaSuccess ? 0 : -1;
}
#endif

extern "C" int srk_attn_block(const srk_attn_block_args* a, void* stream) {
    SRK_REQUIRE(a && a->A && a->Wqkv && a->Wproj && a->b_proj && a->rel_table && a->res && a->out32 && a->out16 && a->ln_g && a->ln_b,
                "attn_block: null pointer");
    if (srk_get_engine() != SRK_ENGINE_TCGEN05)
        return fail(SRK_ERR_UNSUPPORTED, "attn_block: the fused attention block exists for the tcgen05 engine only");
    if (a->Cp != AB_CP || a->num_heads != AB_NH)
        return fail(SRK_ERR_UNSUPPORTED, "attn_block: built for a padded embedding of 192 = 6 heads x 32 (Cp=%d, heads=%d)", a->Cp, a->num_heads);
    SRK_REQUIRE(a->M > 0 && a->H > 0 && a->W > 0 && a->H % 8 == 0 && a->W % 8 == 0 && a->M % (a->H * a->W) == 0,
                "attn_block: rows must be whole images of 8x8 windows");
    SRK_REQUIRE(a->shift == 0 || a->shift == 4, "attn_block: shift must be 0 or window_size / 2");
    SRK_REQUIRE(a->lda >= AB_CP && a->lda % 8 == 0 && a->ld32 >= AB_CP && a->ld32 % 4 == 0 && a->ld16 >= AB_CP && a->ld16 % 8 == 0,
                "attn_block: bad leading dims");
    SRK_REQUIRE(a->ln_C > 0 && a->ln_C <= AB_CP && a->ln_C % 4 == 0 && a->C == a->ln_C, "attn_block: bad LayerNorm width");
    SRK_REQUIRE(((uintptr_t)a->A & 15) == 0 && ((uintptr_t)a->res & 15) == 0 && ((uintptr_t)a->out32 & 15) == 0 && ((uintptr_t)a->out16 & 7) == 0 &&
                ((uintptr_t)a->b_proj & 15) == 0 && ((uintptr_t)a->ln_g & 15) == 0 && ((uintptr_t)a->ln_b & 15) == 0, "attn_block: misaligned pointer");
    AbP p{};
    p.M = a->M; p.m_tiles = ceil_div(a->M, 128); p.C = a->C; p.H = a->H; p.W = a->W; p.T = a->H * a->W; p.shift = a->shift;
    p.scale = a->scale; p.table = a->rel_table; p.b_proj = a->b_proj;
    p.res = a->res; p.out32 = a->out32; p.ld32 = a->ld32;
    p.out16 = (uint16_t*)a->out16; p.ld16 = a->ld16; p.out16_dtype = a->out16_dtype;
    p.ln_g = a->ln_g; p.ln_b = a->ln_b; p.ln_C = a->ln_C;
    CUtensorMap ma, mq, mp;
    {
        cuuint64_t dims[2] = {(cuuint64_t)AB_CP, (cuuint64_t)a->M};
        cuuint64_t strides[1] = {(cuuint64_t)a->lda * 2};
        cuuint32_t box[2] = {64, 128};
        if (int rc = encode_map(&ma, SRK_BF16, 2, a->A, dims, strides, box)) return rc;
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)AB_CP, (cuuint64_t)(AB_NH * 96)};
        cuuint64_t strides[1] = {(cuuint64_t)AB_CP * 2};
        cuuint32_t box[2] = {64, 96};
        if (int rc = encode_map(&mq, SRK_BF16, 2, a->Wqkv, dims, strides, box)) return rc;
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)AB_CP, (cuuint64_t)AB_CP};
        cuuint64_t strides[1] = {(cuuint64_t)AB_CP * 2};
        cuuint32_t box[2] = {64, 96};
        if (int rc = encode_map(&mp, SRK_BF16, 2, a->Wproj, dims, strides, box)) return rc;
    }
    static bool attr[64] = {};
    if (first_use_on_device(attr)) SRK_CUDA(cudaFuncSetAttribute(attn_block_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AB_SMEM));
    ProfScope ps(SRK_PROF_ATTN_BLOCK, stream);
    const int grid = p.m_tiles < num_sms() ? p.m_tiles : num_sms();
    attn_block_tc5_kernel<<<grid, AB_THREADS, AB_SMEM, (cudaStream_t)stream>>>(ma, mq, mp, p);
    SRK_LAUNCH_CHECK("attn_block_tc5_kernel");
    return 0;
}
