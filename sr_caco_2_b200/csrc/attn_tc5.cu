// Fused qkv projection + 8x8-window attention with ALL matrix products on tcgen05 (TMEM):
//   per 128-token tile (2 windows) and head pair:   QKV = A . Wpair^T        (M128 N192 K192)
//                                                   S_h = Q'_h K'_h^T        (M128 N64  K64,  per head)
//                                                   O_h = P'_h [V_w0; V_w1]  (M128 N32  K128, per head, A from TMEM)
// The two windows of a tile share the 128-row MMAs; the window selection is folded into the K
// dimension of both products.  Row r of Q'_h holds q(r) in the 32 columns of ITS window and zeros in
// the other 32, row j of K'_h holds [k_w0(j) | k_w1(j)], so S_h[r][j] = q(r) . k_{window(r)}(j) with
// only 64 accumulator columns.  Row r of P'_h holds its 64 probabilities in the key range of its
// window and zeros in the other window's range, and the K = 128 keys of [V_w0; V_w1] are contracted
// in one chain.  P'_h is written by the softmax threads straight into the TMEM columns of S_h
// (tcgen05.st) and read from there as the A operand, so P never touches shared memory.
//
//   warp 0       TMA producer (A tiles; the head pair's 192 weight rows are loaded once and stay resident)
//   warp 1       tcgen05.mma issuer of the projection (free running, one accumulator)
//   warp 2       tcgen05.mma issuer of S(t) and PV(t-1)
//   warps 3..10  "T" warps: drain the QKV accumulator of tile t (+ bias, bf16) into the UMMA operand
//                tiles (Q', K': 128B-swizzled K-major rows; V: transposed, keys along K, double buffered);
//                they also read the finished O_h rows of tile t-2 out of TMEM (* 1/sum, bf16, 64 B store)
//   warps 11..26 softmax warps, two per (TMEM lane group, head): thread = (row, 32 of its 64 keys):
//                tcgen05.ld S_h, * scale + rel-pos bias + shift mask, row max exchanged with the partner
//                warp, exp2, tcgen05.st P'_h (bf16) over S_h, partial row sum to shared memory
// Two tiles are in flight (S/P' accumulators and V^T are double buffered), so every
// MMA -> mbarrier -> warp hop of one tile is covered by work on the other.  q, k, v, S and P never
// leave the SM.  Restates WindowAttention.forward (dlib/models/network_swinir.py:148-176) and
// calculate_mask (:260-285).
#include "tc5_ptx.cuh"
#include <stdlib.h>

namespace srk {

__device__ __forceinline__ long long gtime() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define QA_TR(role, t, ev) do { if (p.trace && blockIdx.x == 0 && lane == 0 && (t) < 16) p.trace[((role) * 16 + (t)) * 8 + (ev)] = gtime(); } while (0)

constexpr int QA_THREADS = 96 + 32 * 24;        // TMA, 2 MMA issuers, 8 T warps, 16 softmax warps
constexpr int QA_ASTAGES = 4;
constexpr int QA_B_BYTES = 3 * 24576;                  // resident weight tile: 3 k-blocks x (192 rows x 128 B)
constexpr int QA_A_OFF = QA_B_BYTES;
constexpr int QA_Q_OFF = QA_A_OFF + QA_ASTAGES * 16384;    // [head] Q'_h [128][64]
constexpr int QA_K_OFF = QA_Q_OFF + 32768;                 // [head] K'_h [64][64]
constexpr int QA_VT_OFF = QA_K_OFF + 16384;                // [set][window] V^T [64 d-rows (2 heads x 32)][64 keys]
constexpr int QA_BAR_OFF = QA_VT_OFF + 2 * 16384;
constexpr int QA_AUX = 9728;                               // barriers, bias, tables, labels, max exchange, row sums
constexpr size_t QA_SMEM = (size_t)QA_BAR_OFF + QA_AUX + 1024;
static_assert(QA_SMEM <= 232448, "shared memory plan exceeds the 227 KB per-CTA limit");

struct QaP {
    int M, m_tiles, n_pairs, nH, H, W, shift;
    float scale;
    const float* bias; const float* table;
    uint16_t* out; int ldo;
    long long* trace;          // SRK_QA_TRACE: [role 4][tile 16][event 8] globaltimer stamps of CTA 0
};

__global__ void __launch_bounds__(QA_THREADS, 1)
qkv_attn_tc5_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const QaP p) {
    extern __shared__ unsigned char qa_raw[];
    const uint32_t raw = smem_u32(qa_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* sm = qa_raw + (base - raw);
    const uint32_t sB = base, sA = base + QA_A_OFF, sQ = base + QA_Q_OFF, sK = base + QA_K_OFF,
                   sVT = base + QA_VT_OFF, bars = base + QA_BAR_OFF;
    const uint32_t bfull = bars;
    auto a_full = [&](int s) { return bars + 8 + 8u * s; };
    auto a_empty = [&](int s) { return bars + 40 + 8u * s; };
    const uint32_t qkv_full = bars + 72, qkv_empty = bars + 80;
    auto stg_full = [&](int s) { return bars + 88 + 8u * s; };            // T warps -> MMA : Q', K', V^T[s] written
    auto s_full = [&](int s) { return bars + 104 + 8u * s; };             // MMA -> softmax / T warps : S_0, S_1 of set s
    auto p_full = [&](int s, int h) { return bars + 120 + 8u * (2 * s + h); };   // softmax -> MMA : P'_h of set s in TMEM
    auto o_full = [&](int h) { return bars + 152 + 8u * h; };             // MMA -> T warps : O_h
    auto o_empty = [&](int h) { return bars + 168 + 8u * h; };            // T warps -> MMA : O_h read out
    const uint32_t tmem_slot = bars + 184;
    auto vt_free = [&](int s) { return bars + 192 + 8u * s; };            // MMA -> T warps : P V chains of set s retired
    float* sbias = reinterpret_cast<float*>(sm + QA_BAR_OFF + 256);     // [192] local column order
    float* stab = sbias + 192;                                          // [2 heads][225], pre-multiplied by log2(e)
    unsigned char* slab = reinterpret_cast<unsigned char*>(stab + 450); // [2 sets][2 windows][64] shift-mask region labels
    float* smax = reinterpret_cast<float*>(slab + 256);                 // [8 warp pairs][2 halves][32] row-max exchange
    float* ssum = smax + 512;                                           // [2 sets][2 heads][2 halves][128] partial row sums

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pair = blockIdx.x % p.n_pairs;
    const int first_mt = blockIdx.x / p.n_pairs, mt_step = gridDim.x / p.n_pairs;
    const int n_my = first_mt < p.m_tiles ? (p.m_tiles - first_mt + mt_step - 1) / mt_step : 0;
    const float LOG2E = 1.4426950408889634f;

    if (threadIdx.x == 0) {
        mbar_init(bfull, 1);
        for (int s = 0; s < QA_ASTAGES; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); }
        mbar_init(qkv_full, 1); mbar_init(qkv_empty, 8);
        for (int s = 0; s < 2; ++s) {
            mbar_init(stg_full(s), 8); mbar_init(s_full(s), 1); mbar_init(vt_free(s), 1);
            mbar_init(o_full(s), 1); mbar_init(o_empty(s), 8);
            for (int h = 0; h < 2; ++h) mbar_init(p_full(s, h), 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 192; i += QA_THREADS) sbias[i] = p.bias ? p.bias[(i >> 6) * p.nH * 32 + pair * 64 + (i & 63)] : 0.f;
    for (int i = threadIdx.x; i < 450; i += QA_THREADS) stab[i] = p.table[pair * 450 + i] * LOG2E;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(sm + QA_BAR_OFF + 184);
    const uint32_t tQKV = tmem_base;
    auto tSP = [&](int s, int h) { return tmem_base + 192u + 128u * s + 64u * h; };   // S_h fp32, later P'_h bf16 (128 keys)
    auto tO = [&](int h) { return tmem_base + 448u + 32u * h; };

    if (warp == 0) {
        // ======================================= TMA producer =======================================
        if (lane == 0 && n_my > 0) {
            mbar_expect_tx(bfull, QA_B_BYTES);
            for (int kb = 0; kb < 3; ++kb)
#pragma unroll
                for (int w = 0; w < 3; ++w)
                    tma_load_2d(sB + kb * 24576 + w * 8192, &map_b, bfull, kb * 64, w * p.nH * 32 + pair * 64);
            int stage = 0, phase = 0;
            for (int mt = first_mt; mt < p.m_tiles; mt += mt_step) {
                for (int kb = 0; kb < 3; ++kb) {
                    mbar_wait(a_empty(stage), phase ^ 1);
                    mbar_expect_tx(a_full(stage), 16384);
                    tma_load_2d(sA + stage * 16384, &map_a, a_full(stage), kb * 64, mt * 128);
                    if (++stage == QA_ASTAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ======================================= projection MMA issuer =======================================
        if (lane == 0 && n_my > 0) {
            const uint32_t id_qkv = umma_idesc(1, 128, 192);
            int stage = 0, phase = 0;
            mbar_wait(bfull, 0);
            tc_fence_after();
            for (int t = 0; t < n_my; ++t) {
                QA_TR(0, t, 0);
                mbar_wait(qkv_empty, (t & 1) ^ 1);              // accumulator drained by the T warps (tile t-1)
                tc_fence_after();
                QA_TR(0, t, 1);
                for (int kb = 0; kb < 3; ++kb) {
                    mbar_wait(a_full(stage), phase);
                    tc_fence_after();
                    const uint64_t da = umma_desc_sw128(sA + stage * 16384), db = umma_desc_sw128(sB + kb * 24576);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        tc_mma_f16(tQKV, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), id_qkv, (kb | k) != 0 ? 1u : 0u);
                    tc_commit(a_empty(stage));
                    if (++stage == QA_ASTAGES) { stage = 0; phase ^= 1; }
                }
                tc_commit(qkv_full);
                QA_TR(0, t, 2);
            }
        }
    } else if (warp == 2) {
        // ======================================= attention MMA issuer =======================================
        if (lane == 0 && n_my > 0) {
            const uint32_t id_s = umma_idesc(1, 128, 64), id_o = umma_idesc(1, 128, 32);
            auto issue_s = [&](int t) {                         // S_h = Q'_h K'_h^T, K = 64 = 4 x UMMA_K
                const int s = t & 1;
                QA_TR(1, t, 0);
                mbar_wait(stg_full(s), (t >> 1) & 1);
                tc_fence_after();
                QA_TR(1, t, 1);
                // the columns of set s held P'(t-2): its P V chain was issued by this thread before, in order
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint64_t dq = umma_desc_sw128(sQ + h * 16384), dk = umma_desc_sw128(sK + h * 8192);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        tc_mma_f16(tSP(s, h), dq + (uint64_t)(2 * k), dk + (uint64_t)(2 * k), id_s, k != 0 ? 1u : 0u);
                }
                tc_commit(s_full(s));
            };
            auto issue_pv = [&](int t) {                        // O_h = P'_h [V_w0; V_w1], K = 128 keys, A from TMEM
                const int s = t & 1;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (h == 0) QA_TR(1, t, 2);
                    mbar_wait(p_full(s, h), (t >> 1) & 1);
                    QA_TR(1, t, 3 + 2 * h);
                    mbar_wait(o_empty(h), (t & 1) ^ 1);         // O_h of tile t-1 has been read out
                    tc_fence_after();
                    QA_TR(1, t, 4 + 2 * h);
#pragma unroll
                    for (int w = 0; w < 2; ++w) {
                        const uint64_t dv = umma_desc_sw128(sVT + s * 16384 + w * 8192 + h * 4096);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            tc_mma_f16_ts(tO(h), tSP(s, h) + (uint32_t)(w * 32 + k * 8), dv + (uint64_t)(2 * k), id_o,
                                          (w | k) != 0 ? 1u : 0u);
                    }
                    tc_commit(o_full(h));
                }
                tc_commit(vt_free(s));
            };
            for (int t = 0; t < n_my; ++t) {
                issue_s(t);
                if (t > 0) issue_pv(t - 1);
            }
            issue_pv(n_my - 1);
        }
    } else if (warp < 11) {
        // ======================================= T warps =======================================
        const int lg = warp & 3, hq = (warp - 3) >> 2;           // TMEM lane group, half (0/1) = head of the output job
        const int r = lg * 32 + lane;                            // tile row (= TMEM lane) of this thread
        const int wdw = r >> 6, j64 = r & 63;
        const int nW = (p.H >> 3) * (p.W >> 3), wpr = p.W >> 3;
        const int kc = j64 >> 3;
        int mt = first_mt;
        for (int t = 0; t < n_my; ++t, mt += mt_step) {
            const int s = t & 1;
            // ---- A. drain the QKV accumulator.  Q' and K' are single buffered: free once S(t-1) has
            //         retired.  v stays packed in registers until V^T[s] is free.
            if (warp == 3) QA_TR(2, t, 0);
            mbar_wait(qkv_full, t & 1);
            if (warp == 3) QA_TR(2, t, 1);
            if (t >= 1) mbar_wait(s_full((t - 1) & 1), ((t - 1) >> 1) & 1);
            tc_fence_after();
            if (warp == 3) QA_TR(2, t, 2);
            uint32_t pv[2][8];
            const uint32_t tq = tQKV + ((uint32_t)(lg * 32) << 16) + (uint32_t)(hq * 32);
            uint32_t va[16], vb[16];
            tc_ld16_nowait(tq, va);
            tc_wait_ld16(va);
#pragma unroll
            for (int jj = 0; jj < 6; ++jj) {
                const int sec = jj >> 1;                         // 0 = q, 1 = k, 2 = v
                const int cs = hq * 2 + (jj & 1);                // 16-column group inside the section
                const int c = sec * 4 + cs;                      // 16-column sub-chunk of the accumulator
                uint32_t (&v)[16] = (jj & 1) ? vb : va;
                uint32_t (&vn)[16] = (jj & 1) ? va : vb;
                // the next sub-chunk is in flight while this one is converted
                if (jj < 5) tc_ld16_nowait(tq + (uint32_t)(((jj + 1) >> 1) * 64 + ((jj + 1) & 1) * 16), vn);
                const float4* bp = reinterpret_cast<const float4*>(sbias + c * 16);
                uint32_t pk[8];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 bb = bp[j];
                    pk[2 * j] = packf<SRK_BF16>(__uint_as_float(v[4 * j]) + bb.x, __uint_as_float(v[4 * j + 1]) + bb.y);
                    pk[2 * j + 1] = packf<SRK_BF16>(__uint_as_float(v[4 * j + 2]) + bb.z, __uint_as_float(v[4 * j + 3]) + bb.w);
                }
                const int h = cs >> 1, dh = cs & 1;              // head, half of the head dim
                if (sec == 0) {
                    // Q'_h row r: this window's 32 columns hold q, the other window's 32 columns are zero
                    unsigned char* row = sm + QA_Q_OFF + h * 16384 + r * 128;
                    const int ch = wdw * 4 + dh * 2, cz = (wdw ^ 1) * 4 + dh * 2;
                    *reinterpret_cast<uint4*>(row + (((ch) ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    *reinterpret_cast<uint4*>(row + (((ch + 1) ^ (r & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                    *reinterpret_cast<uint4*>(row + (((cz) ^ (r & 7)) << 4)) = make_uint4(0, 0, 0, 0);
                    *reinterpret_cast<uint4*>(row + (((cz + 1) ^ (r & 7)) << 4)) = make_uint4(0, 0, 0, 0);
                } else if (sec == 1) {
                    // K'_h row j (key index inside the window): columns 32*window + d
                    unsigned char* row = sm + QA_K_OFF + h * 8192 + j64 * 128;
                    const int ch = wdw * 4 + dh * 2;
                    *reinterpret_cast<uint4*>(row + (((ch) ^ (j64 & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    *reinterpret_cast<uint4*>(row + (((ch + 1) ^ (j64 & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e) pv[jj & 1][e] = pk[e];
                }
                if (jj < 5) tc_wait_ld16(vn);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(qkv_empty);
            if (warp == 3) QA_TR(2, t, 3);
            // ---- B. V^T[s] was last read by the P V chain of tile t-2: wait for it, pull its O row out of TMEM
            if (t >= 2) mbar_wait(vt_free(s), ((t - 2) >> 1) & 1);
            if (warp == 3) QA_TR(2, t, 4);
            // ---- C. labels and V^T of set s: element (key j, d) -> row d (0..63 = head*32 + d), key column j
            if (hq == 0) {
                const int win = (mt * 2 + wdw) % nW;
                const int wi = win / wpr, wj = win - wi * wpr;
                const bool masked = p.shift > 0 && (wi == (p.H >> 3) - 1 || wj == wpr - 1);
                slab[(s * 2 + wdw) * 64 + j64] = (unsigned char)(masked ? win_pos_label(win, j64, p.H, p.W, p.shift) : 0);
            }
            unsigned char* vt = sm + QA_VT_OFF + s * 16384 + wdw * 8192 + (r & 7) * 2;
#pragma unroll
            for (int e2 = 0; e2 < 2; ++e2) {
                const int cs = hq * 2 + e2;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int d = cs * 16 + i;
                    const uint16_t val = (uint16_t)((i & 1) ? (pv[e2][i >> 1] >> 16) : (pv[e2][i >> 1] & 0xffffu));
                    *reinterpret_cast<uint16_t*>(vt + d * 128 + ((kc ^ (d & 7)) << 4)) = val;
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(stg_full(s));
            if (warp == 3) QA_TR(2, t, 5);
        }
    } else {
        // ======================================= softmax warps =======================================
        // two warps per (TMEM lane group, head): each thread owns one row and 32 of its 64 keys; the
        // row maximum is exchanged through shared memory with a 64-thread named barrier
        const int idx = warp - 11;
        const int lg = warp & 3, h = (idx >> 2) & 1, half = idx >> 3;
        const int r = lg * 32 + lane;
        const int wdw = r >> 6, i = r & 63;                      // window and query position of the row
        const int nW = (p.H >> 3) * (p.W >> 3), wpr = p.W >> 3;
        const float sc2 = p.scale * LOG2E;
        // bias(i, j) = bp[-15*(j>>3) - (j&7)], this thread's keys are j = 32*half + jj
        const float* bp = stab + h * 225 + ((i >> 3) + 7 - 4 * half) * 15 + ((i & 7) + 7);
        const int bar_id = 1 + lg * 2 + h;
        float* mine = smax + ((lg * 2 + h) * 2 + half) * 32 + lane;
        const float* theirs = smax + ((lg * 2 + h) * 2 + (half ^ 1)) * 32 + lane;
        auto output = [&](int t, int mtile) {                  // O_h row of tile t, this thread's 16 of the 32 columns
            mbar_wait(o_full(h), t & 1);
            tc_fence_after();
            uint32_t ov[16];
            tc_ld16_nowait(tO(h) + ((uint32_t)(lg * 32) << 16) + (uint32_t)(half * 16), ov);
            tc_wait_ld16(ov);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(o_empty(h));
            const float* ps = ssum + (((t & 1) * 2 + h) * 2) * 128 + r;
            const float inv = rcp_approx(ps[0] + ps[128]);
            const long long m = (long long)mtile * 128 + r;
            if (m < p.M) {
                uint16_t* dst = p.out + (size_t)m * p.ldo + pair * 64 + h * 32 + half * 16;
#pragma unroll
                for (int j = 0; j < 16; j += 8)
                    *reinterpret_cast<uint4*>(dst + j) = make_uint4(
                        packf<SRK_BF16>(__uint_as_float(ov[j]) * inv, __uint_as_float(ov[j + 1]) * inv),
                        packf<SRK_BF16>(__uint_as_float(ov[j + 2]) * inv, __uint_as_float(ov[j + 3]) * inv),
                        packf<SRK_BF16>(__uint_as_float(ov[j + 4]) * inv, __uint_as_float(ov[j + 5]) * inv),
                        packf<SRK_BF16>(__uint_as_float(ov[j + 6]) * inv, __uint_as_float(ov[j + 7]) * inv));
            }
        };
        int mt = first_mt;
        for (int t = 0; t < n_my; ++t, mt += mt_step) {
            const int s = t & 1;
            const int win = (mt * 2 + wdw) % nW;
            const int wi = win / wpr, wj = win - wi * wpr;
            const bool masked = p.shift > 0 && (wi == (p.H >> 3) - 1 || wj == wpr - 1);
            const uint32_t trow = tSP(s, h) + ((uint32_t)(lg * 32) << 16);
            if (warp == 11) QA_TR(3, t, 0);
            mbar_wait(s_full(s), (t >> 1) & 1);
            tc_fence_after();
            if (warp == 11) QA_TR(3, t, 1);
            float sv[32];
            {
                uint32_t v[32];
                tc_ld32(trow + (uint32_t)(half * 32), v);
#pragma unroll
                for (int j = 0; j < 32; ++j) sv[j] = __uint_as_float(v[j]);
            }
            // scores in the log2 domain: (s * scale + bias) * log2(e); 8 bias loads in flight per batch
#pragma unroll
            for (int jh = 0; jh < 4; ++jh) {
                float b[8];
#pragma unroll
                for (int jw = 0; jw < 8; ++jw) b[jw] = bp[-15 * jh - jw];
#pragma unroll
                for (int jw = 0; jw < 8; ++jw) sv[jh * 8 + jw] = fmaf(sv[jh * 8 + jw], sc2, b[jw]);
            }
            if (masked) {
                const unsigned char* lw = slab + (s * 2 + wdw) * 64;
                const unsigned char li = lw[i];
#pragma unroll
                for (int j = 0; j < 32; ++j) if (lw[half * 32 + j] != li) sv[j] += -100.f * LOG2E;
            }
            float mx0 = -3.0e38f, mx1 = -3.0e38f, mx2 = -3.0e38f, mx3 = -3.0e38f;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                mx0 = fmaxf(mx0, sv[j]); mx1 = fmaxf(mx1, sv[j + 1]); mx2 = fmaxf(mx2, sv[j + 2]); mx3 = fmaxf(mx3, sv[j + 3]);
            }
            float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
            asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");     // partner has read the previous tile's maximum
            *mine = mx;
            asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
            if (warp == 11) QA_TR(3, t, 2);
            mx = fmaxf(mx, *theirs);
            float sum0 = 0.f, sum1 = 0.f;
            uint32_t pk[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const float p0 = ex2_approx(sv[2 * e] - mx);
                const float p1 = ex2_approx(sv[2 * e + 1] - mx);
                sum0 += p0; sum1 += p1;
                pk[e] = packf<SRK_BF16>(p0, p1);
            }
            // P'_h row: probabilities in the key range of this row's window, zeros in the other window's range
            uint32_t zz[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) zz[e] = 0u;
            tc_st16_nowait(trow + (uint32_t)(wdw * 32 + half * 16), pk);
            tc_st16_nowait(trow + (uint32_t)((wdw ^ 1) * 32 + half * 16), zz);
            ssum[((s * 2 + h) * 2 + half) * 128 + r] = sum0 + sum1;
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full(s, h));
            if (warp == 11) QA_TR(3, t, 3);
            if (t > 0) output(t - 1, mt - mt_step);
            if (warp == 11) QA_TR(3, t, 4);             // its P V chain ran under this tile's softmax
        }
        if (n_my > 0) output(n_my - 1, first_mt + (n_my - 1) * mt_step);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// host entry used by gemm_tcgen05() for the fused call when K == 192 (SRK_ATTN_TC5=0 selects the
// mma.sync attention epilogue of gemm_tc5.cu instead)
int qkv_attention_tcgen05(const srk_gemm_args* a, cudaStream_t st) {
    QaP p{};
    p.M = a->M; p.m_tiles = ceil_div(a->M, 128); p.nH = a->attn_heads; p.n_pairs = a->attn_heads / 2;
    p.H = a->H; p.W = a->W; p.shift = a->attn_shift; p.scale = a->attn_scale;
    p.bias = a->bias; p.table = a->attn_table; p.out = (uint16_t*)a->out16; p.ldo = a->ld16;
    SRK_REQUIRE(a->K == 192, "qkv_attention(tcgen05): built for a padded embedding of 192 (K=%d)", a->K);
    CUtensorMap ma, mb;
    {
        cuuint64_t dims[2] = {(cuuint64_t)a->K, (cuuint64_t)a->M};
        cuuint64_t strides[1] = {(cuuint64_t)a->lda * 2};
        cuuint32_t box[2] = {64, 128};
        if (int rc = encode_map(&ma, SRK_BF16, 2, a->A, dims, strides, box)) return rc;
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)a->K, (cuuint64_t)a->N};
        cuuint64_t strides[1] = {(cuuint64_t)a->K * 2};
        cuuint32_t box[2] = {64, 64};
        if (int rc = encode_map(&mb, SRK_BF16, 2, a->Wt, dims, strides, box)) return rc;
    }
    static bool attr[64] = {};
    if (first_use_on_device(attr)) SRK_CUDA(cudaFuncSetAttribute(qkv_attn_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)QA_SMEM));
    int grid = (num_sms() / p.n_pairs) * p.n_pairs;
    if (grid > p.m_tiles * p.n_pairs) grid = p.m_tiles * p.n_pairs;
    static int trace = -1;
    if (trace < 0) { const char* e = getenv("SRK_QA_TRACE"); trace = e ? atoi(e) : 0; }
    if (trace > 0) {                                         // debugging aid: per-tile time stamps of CTA 0
        --trace;
        long long* d; const size_t nb = 4 * 16 * 8 * sizeof(long long);
        SRK_CUDA(cudaMalloc(&d, nb)); SRK_CUDA(cudaMemset(d, 0, nb));
        p.trace = d;
        qkv_attn_tc5_kernel<<<grid, QA_THREADS, QA_SMEM, st>>>(ma, mb, p);
        SRK_CUDA(cudaDeviceSynchronize());
        static long long hbuf[4 * 16 * 8];
        SRK_CUDA(cudaMemcpy(hbuf, d, nb, cudaMemcpyDeviceToHost)); cudaFree(d);
        if (trace == 0) {
            const long long t0 = hbuf[0];
            const char* names[4] = {"qkvmma", "attmma", "Twarp ", "smax  "};
            for (int t = 0; t < 16; ++t)
                for (int r = 0; r < 4; ++r) {
                    fprintf(stderr, "tile %2d %s", t, names[r]);
                    for (int e = 0; e < 8; ++e) { long long v = hbuf[(r * 16 + t) * 8 + e]; fprintf(stderr, " %7lld", v ? v - t0 : -1); }
                    fprintf(stderr, "\n");
                }
        }
        return 0;
    }
    qkv_attn_tc5_kernel<<<grid, QA_THREADS, QA_SMEM, st>>>(ma, mb, p);
    SRK_LAUNCH_CHECK("qkv_attn_tc5_kernel");
    return 0;
}

}  // namespace srk
