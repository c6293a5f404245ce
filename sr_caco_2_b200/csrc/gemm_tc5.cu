// tcgen05 GEMM engine (the product path): persistent, warp-specialised, TMA-fed, TMEM
// accumulators, fused epilogues.
//
//   D[128 x BN] (fp32, TMEM) += A[128 x 64] (smem, 128B swizzle, K-major) * B[BN x 64]^T (smem)
//
//   warp 0      : TMA producer.  A tiles come either from a 2-D tensor map over the (M, K) row
//                 matrix, or -- implicit-GEMM 3x3 convolution -- from a 4-D tensor map over the
//                 NHWC image: for tap (dy,dx) the box (64 ch, BW, BH, 1) is fetched at
//                 (c0, x0+dx-1, y0+dy-1, b); TMA's out-of-bounds zero fill IS the conv padding
//                 and the halo, no im2col buffer exists anywhere.  Weights: 2-D map over (N, K).
//   warp 1      : allocates TMEM (2 accumulator stages), issues tcgen05.mma (one lane), commits
//                 to the smem "empty" barriers and to the TMEM "full" barrier.
//   warps 2..9  : epilogue (two warps per TMEM lane group).  Phase T: tcgen05.ld drains the
//                 accumulator (thread = row) into a padded fp32 staging tile in shared memory
//                 and releases the TMEM stage.  Phase R: the same warps walk the staged rows
//                 with lane = column pair, so every global access is a coalesced 128/256 B row
//                 segment: bias, activation, residual (fp32 stream, window_reverse + roll row
//                 map), fp32 / 16-bit stores, PixelShuffle(2) / pixelshuffle-direct addressing
//                 and, optionally, LayerNorm of the finished row (warp-shuffle statistics).
//   The epilogue of tile i overlaps the MMAs of tile i+1 (two TMEM stages).
#include "tc5_ptx.cuh"
#include "attention.cuh"

namespace srk {

constexpr int TBM = 128, TBK = 64;
constexpr int TC_THREADS = 320;           // 10 warps: TMA, MMA, 8 epilogue
constexpr int TC_SMEM_TOTAL = 227 * 1024;

struct TcParams {
    GemmP g;
    int n_tiles, m_tiles, nkb;
    int conv_bw, conv_bh, conv_tx, conv_ty;     // conv-mode M tiling (box BW x BH = 128 pixels)
    // shared-memory plan (host computed): B-stationary keeps the CTA's whole weight tile
    // (nkb x BN x 64) resident and streams only A through `stages` 16 KB slots
    int bstat, stages, stage_bytes, bres_bytes;
    // conv halo mode (64-output-channel convs): ONE TMA box (64 ch, 8 + 2r, 16 + 2r) per (tile, channel block) feeds
    // all kt * kt taps as shifted-window A descriptors into the swizzled halo tile; `stages` / `stage_bytes` then
    // describe the ring of weight k-blocks (none when the weights are resident).  These convs are bound by the TMA
    // request rate of the per-tap boxes (~5 cycles per 128 B pixel row, 9 or 25 boxes per channel block), not by bytes.
    int halo, halo_r, halo_box_bytes, halo_stride, a_stages;
};

#ifndef SRK_LN_ROW_BATCH
#define SRK_LN_ROW_BATCH 2                                 // E_RES_LN epilogue: rows processed together (measured: 2 is best, 4 spills)
#endif
#ifndef SRK_FP32_EPI_WARPS
#define SRK_FP32_EPI_WARPS 8                               // epilogue warps of the fp32 (residual / LayerNorm) path: 8 or 16.  Measured: 16 is 6 % slower -- warps are
                                                           // allocated in groups of 4, so 18 warps count as 20 and cap the kernel at 96 registers (spills)
#endif

template <int BN, int EPI>
struct TcCfg {
    static constexpr bool kStage16 = EPI == E_O16 || EPI == E_PIXSHUF || EPI == E_ATTN;   // 16-bit staging (math in phase T)
    static constexpr int EPI_WARPS = (kStage16 || SRK_FP32_EPI_WARPS == 16) ? 16 : 8;   // warps per TMEM lane group: 4 or 2
    // 16-bit epilogues whose staging fits twice run as TWO warp groups that take alternate tiles, so
    // the TMEM drain of one tile (64 B/clk port) overlaps the global stores of the other
    static constexpr int kGroups = (EPI == E_ATTN || (EPI == E_O16 && BN <= 192)) ? 2 : 1;
    static constexpr int THREADS = 64 + 32 * EPI_WARPS;
    static constexpr int A_BYTES = TBM * TBK * 2;
    static constexpr int B_BYTES = BN * TBK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int SROW = BN + 4;                       // fp32 staging row stride (floats)
    static constexpr int SROW16 = BN * 2 + 16;                // 16-bit staging row stride (bytes)
    static constexpr int STAGING_BYTES = kStage16 ? kGroups * TBM * SROW16 : TBM * SROW * 4;
    // barriers + (16-bit epilogues) 2 slots per warp group of [bias row | row map of the 128 tile rows],
    // (fp32 epilogues) per-lane-group scratch
    static constexpr int AUX_BYTES = EPI == E_ATTN ? 4096 : (kStage16 ? 256 + kGroups * 2 * (BN + 128) * 4 : 256 + 8 * (BN + 32) * 4);
    static constexpr int AVAIL = TC_SMEM_TOTAL - STAGING_BYTES - 1024 - AUX_BYTES;   // operand bytes available
    static constexpr int MAX_STAGES = 8;
    static constexpr int TMEM_COLS = (2 * BN <= 128) ? 128 : (2 * BN <= 256 ? 256 : 512);
    static_assert(AVAIL / STAGE_BYTES >= 3, "not enough shared memory for the operand pipeline");
    static_assert(BN % 64 == 0 && (BN <= 192 || (BN == 256 && kStage16)), "BN must be 64, 128, 192 (or 256 with 16-bit staging)");
};

// row (0..127) of M tile `mt` -> GEMM row index m (or -1 when the row is outside the problem)
__device__ __forceinline__ int tile_row_to_m(const TcParams& p, int mt, int r) {
    const GemmP& g = p.g;
    if (g.a_mode == SRK_A_ROWS) {
        const int m = mt * TBM + r;
        return m < g.M ? m : -1;
    }
    const int per_img = p.conv_tx * p.conv_ty;
    const int b = mt / per_img, t = mt - b * per_img;
    const int ty = t / p.conv_tx, tx = t - ty * p.conv_tx;
    const int py = r / p.conv_bw, px = r - py * p.conv_bw;
    const int y = ty * p.conv_bh + py, x = tx * p.conv_bw + px;
    if (y >= g.H || x >= g.W) return -1;
    return (b * g.H + y) * g.W + x;
}

template <int BN, int EPI, int ACT, int DT>
__global__ void __launch_bounds__((TcCfg<BN, EPI>::THREADS), 1)
gemm_tc5_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                const TcParams p) {
    using Cfg = TcCfg<BN, EPI>;
    extern __shared__ unsigned char tc_smem_raw[];
    const uint32_t raw = smem_u32(tc_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;                     // SW128 tiles need 1024 B alignment
    // [resident B (B-stationary only)] [operand stages] [epilogue staging] [barriers]
    const uint32_t halo_base = base + p.bres_bytes;                   // halo ring (conv halo mode only)
    const uint32_t stages_base = halo_base + p.a_stages * p.halo_stride;
    const uint32_t stg_base = stages_base + p.stages * p.stage_bytes;
    const uint32_t bars = stg_base + Cfg::STAGING_BYTES;    // full[8] empty[8] tfull[2] tempty[2] bfull tmem_ptr
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (Cfg::MAX_STAGES + s); };
    auto tfull_bar = [&](int s) { return bars + 8u * (2 * Cfg::MAX_STAGES + s); };
    auto tempty_bar = [&](int s) { return bars + 8u * (2 * Cfg::MAX_STAGES + 2 + s); };
    const uint32_t bfull_bar = bars + 8u * (2 * Cfg::MAX_STAGES + 4);
    const uint32_t tmem_slot = bars + 8u * (2 * Cfg::MAX_STAGES + 5);
    auto afull_bar = [&](int s) { return bars + 8u * (2 * Cfg::MAX_STAGES + 6 + s); };      // halo ring (<= 4 stages)
    auto aempty_bar = [&](int s) { return bars + 8u * (2 * Cfg::MAX_STAGES + 10 + s); };
    const int NS = p.stages;
    // B-stationary: CTA c owns n-tile c % n_tiles and M tiles c / n_tiles, + gridDim / n_tiles, ...
    // (the host makes gridDim a multiple of n_tiles), so "tile += gridDim.x" keeps nt fixed.
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(tc_smem_raw + (tmem_slot - raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const GemmP& g = p.g;
    const int total_tiles = p.m_tiles * p.n_tiles;

    if (threadIdx.x == 0) {
        for (int s = 0; s < Cfg::MAX_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        mbar_init(bfull_bar, 1);
        for (int s = 0; s < 4; ++s) { mbar_init(afull_bar(s), 1); mbar_init(aempty_bar(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), Cfg::EPI_WARPS / Cfg::kGroups); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                     "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    // resident weight tile, loaded once
    if (warp == 0 && lane == 0 && p.bstat && (int)blockIdx.x < total_tiles) {
        const int ntb = blockIdx.x % p.n_tiles;
        mbar_expect_tx(bfull_bar, (uint32_t)p.bres_bytes);
        for (int kb = 0; kb < p.nkb; ++kb) {
            if (EPI == E_ATTN) {
                // head pair ntb: rows [q | k | v] x (2 heads x 32) gathered from the [which][head][32] weight
#pragma unroll
                for (int w = 0; w < 3; ++w)
                    tma_load_2d(base + kb * Cfg::B_BYTES + w * 8192, &map_b, bfull_bar, kb * TBK,
                                w * g.attn_heads * 32 + ntb * 64);
            } else {
                tma_load_2d(base + kb * Cfg::B_BYTES, &map_b, bfull_bar, kb * TBK, ntb * BN);
            }
        }
    }

    if (warp == 0) {
        // ================================ TMA producer ================================
        {   // the whole warp runs the loop (uniform control flow, address arithmetic on the uniform datapath); one elected
            // lane arms the barriers and issues the TMA loads
            int stage = 0, phase = 0, hstage = 0, hphase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int mt = tile / p.n_tiles, nt = tile - mt * p.n_tiles;
                int cb_, cx = 0, cy = 0;
                if (g.a_mode == SRK_A_CONV3X3) {
                    const int per_img = p.conv_tx * p.conv_ty;
                    cb_ = mt / per_img;
                    const int t = mt - cb_ * per_img;
                    cy = (t / p.conv_tx) * p.conv_bh;
                    cx = (t - (t / p.conv_tx) * p.conv_tx) * p.conv_bw;
                } else {
                    cb_ = 0;
                }
                if (p.halo) {
                    // channel-block major: one halo box, then (unless resident) the kt * kt weight k-blocks that use it
                    const int taps = g.kt * g.kt;
                    for (int cblk = 0; cblk < g.cpb; ++cblk) {
                        mbar_wait(aempty_bar(hstage), hphase ^ 1);
                        if (elect_one()) {
                            mbar_expect_tx(afull_bar(hstage), (uint32_t)p.halo_box_bytes);
                            tma_load_4d(halo_base + hstage * p.halo_stride, &map_a, afull_bar(hstage), cblk * 64,
                                        cx - p.halo_r, cy - p.halo_r, cb_);
                        }
                        __syncwarp();
                        if (++hstage == p.a_stages) { hstage = 0; hphase ^= 1; }
                        if (!p.bstat) {
                            for (int tap = 0; tap < taps; ++tap) {
                                mbar_wait(empty_bar(stage), phase ^ 1);
                                if (elect_one()) {
                                    mbar_expect_tx(full_bar(stage), (uint32_t)Cfg::B_BYTES);
                                    tma_load_2d(stages_base + stage * p.stage_bytes, &map_b, full_bar(stage),
                                                (tap * g.cpb + cblk) * TBK, nt * BN);
                                }
                                __syncwarp();
                                if (++stage == NS) { stage = 0; phase ^= 1; }
                            }
                        }
                    }
                    continue;
                }
                // (tap, channel block) of k-block kb = tap * cpb + cblk are carried as counters: this single thread is the
                // kernel's critical path for the convs (two integer divisions per k-block were ~2/3 of its instructions)
                int cblk = 0, dx = -(g.kt >> 1), dy = -(g.kt >> 1);
                for (int kb = 0; kb < p.nkb; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    const uint32_t sa = stages_base + stage * p.stage_bytes, sb = sa + Cfg::A_BYTES;
                    if (elect_one()) {
                        mbar_expect_tx(full_bar(stage), (uint32_t)p.stage_bytes);
                        if (g.a_mode == SRK_A_CONV3X3) tma_load_4d(sa, &map_a, full_bar(stage), cblk * 64, cx + dx, cy + dy, cb_);
                        else tma_load_2d(sa, &map_a, full_bar(stage), kb * TBK, mt * TBM);
                        if (!p.bstat) tma_load_2d(sb, &map_b, full_bar(stage), kb * TBK, nt * BN);
                    }
                    __syncwarp();
                    if (g.a_mode == SRK_A_CONV3X3 && ++cblk == g.cpb) {
                        cblk = 0;
                        if (++dx > (g.kt >> 1)) { dx = -(g.kt >> 1); ++dy; }
                    }
                    if (++stage == NS) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ================================
        {   // warp-uniform like the producer: descriptors stay in uniform registers, one elected lane issues MMAs and commits
            const uint32_t idesc = umma_idesc(g.dtype == SRK_BF16 ? 1 : 0, TBM, BN);
            int stage = 0, phase = 0, it = 0, hstage = 0, hphase = 0;
            if (p.bstat && (int)blockIdx.x < total_tiles) { mbar_wait(bfull_bar, 0); tc_fence_after(); }
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
                const int as = it & 1, aphase = (it >> 1) & 1;
                mbar_wait(tempty_bar(as), aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
                if (p.halo) {
                    const int taps = g.kt * g.kt, hw = 8 + 2 * p.halo_r;          // halo width in pixels (= 128 B rows)
                    for (int cblk = 0; cblk < g.cpb; ++cblk) {
                        mbar_wait(afull_bar(hstage), hphase);
                        tc_fence_after();
                        const uint32_t ha = halo_base + hstage * p.halo_stride;
                        int ty = 0, tx = 0;
                        for (int tap = 0; tap < taps; ++tap) {
                            const int kb = tap * g.cpb + cblk;
                            uint32_t sb = base + kb * Cfg::B_BYTES;
                            if (!p.bstat) {
                                mbar_wait(full_bar(stage), phase);
                                tc_fence_after();
                                sb = stages_base + stage * p.stage_bytes;
                            }
                            // A = the 8 x 16 pixel window of the halo tile shifted by the tap: 16 core groups (image
                            // rows) hw * 128 B apart, starting at halo pixel (tap / kt, tap % kt).  Legal because the
                            // 128 B swizzle is a function of the absolute shared-memory address (1024 B aligned tile).
                            const uint64_t da = umma_desc_sw128_sbo(ha + (uint32_t)(ty * hw + tx) * 128u, (uint32_t)hw * 128u);
                            if (++tx == g.kt) { tx = 0; ++ty; }
                            const uint64_t db = umma_desc_sw128(sb);
                            if (elect_one()) {
#pragma unroll
                                for (int k = 0; k < TBK / 16; ++k)
                                    tc_mma_f16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (cblk | tap | k) != 0 ? 1u : 0u);
                                if (!p.bstat) tc_commit(empty_bar(stage));
                                if (tap == taps - 1) {
                                    tc_commit(aempty_bar(hstage));
                                    if (cblk == g.cpb - 1) tc_commit(tfull_bar(as));
                                }
                            }
                            __syncwarp();
                            if (!p.bstat && ++stage == NS) { stage = 0; phase ^= 1; }
                        }
                        if (++hstage == p.a_stages) { hstage = 0; hphase ^= 1; }
                    }
                    continue;
                }
                for (int kb = 0; kb < p.nkb; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t sa = stages_base + stage * p.stage_bytes;
                    const uint32_t sb = p.bstat ? base + kb * Cfg::B_BYTES : sa + Cfg::A_BYTES;
                    const uint64_t da = umma_desc_sw128(sa), db = umma_desc_sw128(sb);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < TBK / 16; ++k)    // +32 B per UMMA_K inside the swizzle atom
                            tc_mma_f16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                                       (kb | k) != 0 ? 1u : 0u);
                        tc_commit(empty_bar(stage));           // frees the smem slot when the MMAs retire
                        if (kb == p.nkb - 1) tc_commit(tfull_bar(as));   // accumulator complete
                    }
                    __syncwarp();
                    if (++stage == NS) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ================================ epilogue ================================
        if constexpr (EPI == E_ATTN) {
            // ---- fused window attention: the CTA's tile is 2 windows x (q|k|v of one head pair).
            //      The 16 epilogue warps form two groups of 8 that take alternate tiles (group g uses
            //      TMEM stage g and staging buffer g), so the drain / attention / store phases of two
            //      tiles interleave.  Phase T stages the tile as bf16 rows (the attention kernel's smem
            //      layout); the 4 warps of a window then run its 8 (head, strip) units and only O is stored.
            const int ew = warp - 2, lg = warp & 3;
            const int grp = ew >> 3, hq = (ew >> 2) & 1;
            const int window = lg >> 1, strip = (lg & 1) * 2 + hq;
            const int tid_w = (lg & 1) * 64 + hq * 32 + lane;             // 0..127 inside the (group, window) warps
            unsigned char* stg_all = tc_smem_raw + (stg_base - raw) + (size_t)grp * TBM * Cfg::SROW16;
            unsigned char* wrows = stg_all + (size_t)(window * 64) * Cfg::SROW16;
            const uint32_t wrows_s = stg_base + (uint32_t)(grp * TBM + window * 64) * Cfg::SROW16;
            float* sbias = reinterpret_cast<float*>(tc_smem_raw + (bars + 256 - raw));     // [192] local column order
            float* stab = sbias + BN;                                                     // [2][225]
            int* slab = reinterpret_cast<int*>(stab + 2 * 225) + (grp * 2 + window) * 64;  // [grp][window][64]
            const int pair = blockIdx.x % p.n_tiles;
            const int nH = g.attn_heads;
            const bool has_bias = g.bias != nullptr;
            {
                const int et = threadIdx.x - 64;                                          // 0..511
                if (et < BN && has_bias) sbias[et] = __ldg(g.bias + (et >> 6) * nH * 32 + pair * 64 + (et & 63));
                if (et < 450) stab[et] = __ldg(g.attn_table + pair * 450 + et);
                asm volatile("bar.sync 9, 512;" ::: "memory");
            }
            const int nW = (g.H >> 3) * (g.W >> 3), wpr = g.W >> 3;
            constexpr int SC = BN / 16;
            const int bar_id = 5 + grp * 2 + window;
            for (int tile = blockIdx.x + grp * gridDim.x, it = grp; tile < total_tiles; tile += 2 * gridDim.x, it += 2) {
                const int mt = tile / p.n_tiles;
                const int as = grp, aphase = (it >> 1) & 1;
                const int wg = mt * 2 + window;                         // global window index
                const bool wvalid = (long long)wg * 64 < g.M;
                const int win = wg % nW;
                const int wi = win / wpr, wj = win - wi * wpr;
                const bool masked = g.attn_shift > 0 && (wi == (g.H >> 3) - 1 || wj == wpr - 1);
                if (tid_w < 64) slab[tid_w] = masked ? win_pos_label(win, tid_w, g.H, g.W, g.attn_shift) : 0;
                mbar_wait(tfull_bar(as), aphase);
                tc_fence_after();
                {
                    const uint32_t t_row = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(as * BN);
                    unsigned char* srow = stg_all + (size_t)(lg * 32 + lane) * Cfg::SROW16;
                    auto convert16 = [&](const uint32_t* v, int c) {      // 16 accumulator columns -> 32 B of the staged row
                        uint32_t pk[8];
                        if (has_bias) {
                            const float4* bp = reinterpret_cast<const float4*>(sbias + c * 16);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float4 bb = bp[j];
                                pk[2 * j] = packf<SRK_BF16>(__uint_as_float(v[4 * j]) + bb.x, __uint_as_float(v[4 * j + 1]) + bb.y);
                                pk[2 * j + 1] = packf<SRK_BF16>(__uint_as_float(v[4 * j + 2]) + bb.z, __uint_as_float(v[4 * j + 3]) + bb.w);
                            }
                        } else {                                  // bias folded into the GEMM (pad columns of A hold 1)
#pragma unroll
                            for (int j = 0; j < 8; ++j) pk[j] = packf<SRK_BF16>(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
                        }
                        *reinterpret_cast<uint4*>(srow + c * 32) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        *reinterpret_cast<uint4*>(srow + c * 32 + 16) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                    };
#pragma unroll
                    for (int jj = 0; jj < SC / 4; ++jj) {          // 32-column TMEM loads, interleaved between the two warps
                        const int c2 = hq + 2 * jj;
                        uint32_t v[32];
                        tc_ld32(t_row + c2 * 32, v);
                        convert16(v, 2 * c2);
                        convert16(v + 16, 2 * c2 + 1);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty_bar(as));
                asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");       // window staged (+ labels)
                if (wvalid) {
#pragma unroll 1
                    for (int hl = 0; hl < 2; ++hl)
                        attn_unit<32>(wrows_s, wrows, Cfg::SROW16, strip, hl * 32, 64 + hl * 32, 128 + hl * 32,
                                      stab + hl * 225, slab, masked, g.attn_scale, lane);
                }
                asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");       // every unit wrote its O
                // attention output of the window: 64 rows x (2 heads x 32) = the first 128 B of each staged row
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int idx = tid_w + 128 * jj;
                    const int row = idx >> 3, col = idx & 7;
                    const long long m = (long long)wg * 64 + row;
                    if (m < g.M) {
                        const uint4 val = *reinterpret_cast<const uint4*>(wrows + (size_t)row * Cfg::SROW16 + col * 16);
                        *reinterpret_cast<uint4*>(g.out16 + (size_t)m * g.ld16 + pair * 64 + col * 8) = val;
                    }
                }
                asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");       // staging free for this group's next tile
            }
        } else if constexpr (Cfg::kStage16) {
            // ---- 16-bit outputs: bias + activation + pack in phase T (thread = row), then a pure
            //      coalesced 16 B/lane copy of the staged rows in phase R ----
            // G warp groups take alternate tiles (group g: TMEM stage g, staging buffer g).  Inside a
            // group WPL warps share a TMEM lane group: warp q drains 16-column sub-chunks q, q+WPL, ...
            // in phase T and copies rows q*RPW .. of the lane group in phase R.
            constexpr int G = Cfg::kGroups;
            constexpr int WPL = 4 / G;
            constexpr int RPW = 32 / WPL;
            const int ew = warp - 2, lg = warp & 3;
            const int grp = G == 2 ? (ew >> 3) : 0;
            const int q = (ew >> 2) & (WPL - 1);
            unsigned char* stg16 = tc_smem_raw + (stg_base - raw) + (size_t)(grp * TBM + lg * 32) * Cfg::SROW16;
            // [bias row | row map of the tile's 128 rows], two slots per group (late readers of the previous tile)
            float* aux0 = reinterpret_cast<float*>(tc_smem_raw + (bars + 256 - raw)) + grp * 2 * (BN + 128);
            constexpr int SC = BN / 16;                            // 16-column sub-chunks
            constexpr int CPR = BN / 8;                            // 16 B chunks per row
            const int bar_id = 1 + grp;
            constexpr int bar_n = 32 * 4 * WPL;                    // the group's warps
            static_assert(SC % WPL == 0 && (RPW * CPR) % 32 == 0, "tile shape does not split over the warps");
            for (int tile = blockIdx.x + grp * gridDim.x, it = grp; tile < total_tiles; tile += G * gridDim.x, it += G) {
                const int mt = tile / p.n_tiles, nt = tile - mt * p.n_tiles;
                const int as = it & 1, aphase = (it >> 1) & 1;
                const int n0 = nt * BN;
                float* sbias = aux0 + ((it / G) & 1) * (BN + 128);
                int* srowm = reinterpret_cast<int*>(sbias + BN) + lg * 32;
                if (q == 0) {
                    if (lg == 0) {
#pragma unroll
                        for (int j = 0; j < BN / 32; ++j)      // the packed GELU takes the halved bias
                            sbias[lane + 32 * j] = __ldg(g.bias + n0 + lane + 32 * j) * (ACT == SRK_ACT_GELU ? 0.5f : 1.f);
                    }
                    int mm = tile_row_to_m(p, mt, lg * 32 + lane);
                    if (EPI == E_PIXSHUF && mm >= 0) {
                        // PixelShuffle(2): store the index of output pixel (b, 2y, 2x) instead of m, so the
                        // per-chunk address needs no integer division
                        const int bi = mm / g.T, rem = mm - bi * g.T;
                        const int y = rem / g.W, x = rem - y * g.W;
                        mm = (bi * 2 * g.H + 2 * y) * (2 * g.W) + 2 * x;
                    }
                    srowm[lane] = mm;
                }
                asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(bar_n) : "memory");
                mbar_wait(tfull_bar(as), aphase);
                tc_fence_after();
                {
                    const uint32_t t_row = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(as * BN);
                    unsigned char* srow = stg16 + (size_t)lane * Cfg::SROW16;
                    // convert 16 accumulator columns (+ bias, activation) and stage them as 32 B of the row
                    auto convert16 = [&](const uint32_t* v, int c) {
                        const float4* bp = reinterpret_cast<const float4*>(sbias + c * 16);
                        uint32_t pk[8];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4 bb = bp[j];
                            if constexpr (ACT == SRK_ACT_GELU) {
                                uint32_t a0, a1;
                                unpack64(gelu2(pack64(v[4 * j], v[4 * j + 1]), pack64(__float_as_uint(bb.x), __float_as_uint(bb.y))), a0, a1);
                                pk[2 * j] = packf<DT>(__uint_as_float(a0), __uint_as_float(a1));
                                unpack64(gelu2(pack64(v[4 * j + 2], v[4 * j + 3]), pack64(__float_as_uint(bb.z), __float_as_uint(bb.w))), a0, a1);
                                pk[2 * j + 1] = packf<DT>(__uint_as_float(a0), __uint_as_float(a1));
                            } else {
                                pk[2 * j] = packf<DT>(actf<ACT>(__uint_as_float(v[4 * j]) + bb.x),
                                                      actf<ACT>(__uint_as_float(v[4 * j + 1]) + bb.y));
                                pk[2 * j + 1] = packf<DT>(actf<ACT>(__uint_as_float(v[4 * j + 2]) + bb.z),
                                                          actf<ACT>(__uint_as_float(v[4 * j + 3]) + bb.w));
                            }
                        }
                        *reinterpret_cast<uint4*>(srow + c * 32) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        *reinterpret_cast<uint4*>(srow + c * 32 + 16) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                    };
                    constexpr bool kX32 = (BN / 32) % WPL == 0;    // 32-column TMEM loads when they split evenly
                    if constexpr (kX32) {
#pragma unroll
                        for (int jj = 0; jj < (BN / 32) / WPL; ++jj) {
                            const int c2 = q + WPL * jj;               // 32-column chunk
                            uint32_t v[32];
                            tc_ld32(t_row + c2 * 32, v);
                            convert16(v, 2 * c2);
                            convert16(v + 16, 2 * c2 + 1);
                        }
                    } else {
#pragma unroll
                        for (int jj = 0; jj < SC / WPL; ++jj) {
                            const int c = q + WPL * jj;
                            uint32_t v[16];
                            tc_ld16_nowait(t_row + c * 16, v);
                            tc_wait_ld16(v);
                            convert16(v, c);
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty_bar(as));
                asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(bar_n) : "memory");
                // phase R: this warp's RPW rows, CPR 16 B chunks each
                const unsigned char* sbase = stg16 + (size_t)(q * RPW) * Cfg::SROW16;
#pragma unroll
                for (int jj = 0; jj < (RPW * CPR) / 32; ++jj) {
                    const int idx = lane + 32 * jj;
                    const int row = idx / CPR, col = idx - row * CPR;
                    const int m = srowm[q * RPW + row];
                    const uint4 val = *reinterpret_cast<const uint4*>(sbase + (size_t)row * Cfg::SROW16 + col * 16);
                    if (m >= 0) {
                        uint16_t* dst;
                        if (EPI == E_PIXSHUF) {
                            // column n -> sub-pixel group n / (N/4), channel n % (N/4); N/4 is a multiple of 64
                            const int cq = g.N >> 2, n = n0 + col * 8;
                            const int grp4 = n / cq, ch = n - grp4 * cq;
                            dst = g.out16 + ((size_t)m + (grp4 >> 1) * (2 * g.W) + (grp4 & 1)) * g.ld16 + ch;
                        } else {
                            dst = g.out16 + (size_t)m * g.ld16 + n0 + col * 8;
                        }
                        *reinterpret_cast<uint4*>(dst) = val;
                    }
                }
                // (the bar.sync at the top of the group's next tile also protects the staging rows)
            }
        } else {
        constexpr int WPL = Cfg::EPI_WARPS / 4;                // warps per TMEM lane group (2 or 4)
        constexpr int RPW = 32 / WPL;                          // rows per warp in phase R (16 or 8)
        const int ew = warp - 2;                               // 0 .. EPI_WARPS-1
        const int lg = warp & 3;                               // TMEM lane group this warp may access
        const int half = ew >> 2;                              // which column chunks in phase T, which rows in phase R
        float* stg = reinterpret_cast<float*>(tc_smem_raw + (stg_base - raw)) + (size_t)(lg * 32) * Cfg::SROW;
        constexpr int NP = BN / 64;                            // column pairs per lane
        constexpr bool kRes = EPI == E_RES_LN || EPI == E_RES;
        constexpr bool kPrefetch = kRes && NP <= 3;            // residual rows held in registers (<= 96)
        const bool has_res = EPI == E_GENERIC ? (g.res != nullptr) : kRes;
        const bool has_ln = EPI == E_GENERIC ? (g.ln_g != nullptr) : (EPI == E_RES_LN);
        const int act = g.act;
        const float inv_c = 1.f / (float)(g.ln_C > 0 ? g.ln_C : 1);
        int it = 0;
        float2 lng[NP], lnb[NP];
        if (has_ln) {
#pragma unroll
            for (int k = 0; k < NP; ++k) {
                const int n = 64 * k + 2 * lane;
                lng[k] = n < g.ln_C ? __ldg(reinterpret_cast<const float2*>(g.ln_g + n)) : make_float2(0.f, 0.f);
                // pad columns: beta doubles as the constant written there (1.0 in ln_C, ln_C + 1 when ln_pad_one)
                lnb[k] = n < g.ln_C ? __ldg(reinterpret_cast<const float2*>(g.ln_b + n))
                                    : ((g.ln_pad_one && n == g.ln_C) ? make_float2(1.f, 1.f) : make_float2(0.f, 0.f));
            }
        }
        // row bookkeeping of one tile: lane i (< RPW) describes row i of this warp's phase-R rows
        // image epilogue (pixelshuffle-direct): lane i also carries the row's position in the cropped image -- y * s, x * s
        // and the image base -- so that the per-element addressing is adds and compares (it was three integer divisions
        // per element: the folded tail conv spent most of its time there)
        const bool has_img = EPI == E_GENERIC && g.img != nullptr;
        int my_iy = 0, my_ix = 0, my_ib = 0;
        auto img_of = [&](int m_, int& iy_, int& ix_, int& ib_) {
            iy_ = 0; ix_ = 0; ib_ = 0;
            if (has_img && m_ >= 0) {
                const int bi = m_ / g.T, rem = m_ - bi * g.T;
                const int y = rem / g.W, x = rem - y * g.W;
                iy_ = y * g.img_s; ix_ = x * g.img_s; ib_ = bi * g.img_hc * g.img_wc;
            }
        };
        auto rows_of = [&](int tile_, int& m_, int& r32_, int& r16_) {
            m_ = -1; r32_ = 0; r16_ = 0;
            if (tile_ >= total_tiles) return;
            const int mt_ = tile_ / p.n_tiles;
            m_ = lane < RPW ? tile_row_to_m(p, mt_, lg * 32 + half * RPW + lane) : -1;
            if (m_ >= 0) {
                r32_ = (has_res || g.out32 || has_ln) ? row32_of(g, m_) : m_;
                r16_ = m_;
                if (has_ln) {
                    r16_ = r32_;
                    if (g.ln_win_shift >= 0) {
                        const int bi = r32_ / g.T;
                        r16_ = bi * g.T + token_to_win_pos(r32_ - bi * g.T, g.H, g.W, g.ln_win_shift);
                    }
                }
            }
        };
        // residual rows live in registers and are software-pipelined one tile ahead: as soon as
        // row i of the current tile is finished, row i of the NEXT tile is requested into the
        // same registers, so the HBM latency hides behind the rest of phase R + the next phase T
        float2 resv[kPrefetch ? RPW : 1][NP];
        int my_m, my_r32, my_r16;
        rows_of(blockIdx.x, my_m, my_r32, my_r16);
        img_of(my_m, my_iy, my_ix, my_ib);
        if (kPrefetch) {
            const int n0f = (blockIdx.x % p.n_tiles) * BN;
#pragma unroll
            for (int i = 0; i < RPW; ++i) {
                const int m = __shfl_sync(0xffffffffu, my_m, i);
                const int r32 = __shfl_sync(0xffffffffu, my_r32, i);
                const float* rr = g.res + (size_t)r32 * g.ld32 + n0f + 2 * lane;
#pragma unroll
                for (int k = 0; k < NP; ++k)
                    resv[kPrefetch ? i : 0][k] = m >= 0 ? __ldg(reinterpret_cast<const float2*>(rr + 64 * k)) : make_float2(0.f, 0.f);
            }
        }
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int nt = tile % p.n_tiles;
            const int as = it & 1, aphase = (it >> 1) & 1;
            const int n0 = nt * BN;
            int nx_m, nx_r32, nx_r16;
            rows_of(tile + gridDim.x, nx_m, nx_r32, nx_r16);
            // this lane's two columns of each 64-column group as sub-pixel (i, j) of the s x s block (-1: pad column)
            int ci[NP][2], cj[NP][2];
            if (has_img) {
#pragma unroll
                for (int k = 0; k < NP; ++k)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int n = n0 + 64 * k + 2 * lane + e;
                        ci[k][e] = n < g.img_s * g.img_s ? n / g.img_s : -1;
                        cj[k][e] = n - ci[k][e] * g.img_s;
                    }
            }
            const int n0x = ((tile + (int)gridDim.x) % p.n_tiles) * BN;
            float2 bia[NP];
#pragma unroll
            for (int k = 0; k < NP; ++k) bia[k] = __ldg(reinterpret_cast<const float2*>(g.bias + n0 + 64 * k + 2 * lane));

            // ---- phase T: TMEM -> staging (thread = row, this warp's column chunks) ----
            mbar_wait(tfull_bar(as), aphase);
            tc_fence_after();
            {
                const uint32_t t_row = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(as * BN);
                float* srow = stg + (size_t)lane * Cfg::SROW;
                constexpr int SC = BN / 16;                    // 16-column sub-chunks, interleaved between the WPL warps
                if constexpr (WPL == 2) {
                    constexpr int CH = BN / 32;                // 32-column chunks, interleaved between the two warps
#pragma unroll 1
                    for (int c = half; c < CH; c += 2) {       // (measured: x32 loads beat pipelined x16 loads here)
                        uint32_t v[32];
                        tc_ld32(t_row + c * 32, v);
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            *reinterpret_cast<uint4*>(srow + c * 32 + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    }
                } else {
#pragma unroll 1
                    for (int c = half; c < SC; c += WPL) {
                        uint32_t v[16];
                        tc_ld16_nowait(t_row + c * 16, v);
                        tc_wait_ld16(v);
#pragma unroll
                        for (int j = 0; j < 16; j += 4)
                            *reinterpret_cast<uint4*>(srow + c * 16 + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(as));        // accumulator stage drained
            asm volatile("bar.sync %0, %1;" ::"r"(1 + lg), "n"(32 * WPL) : "memory");   // the lane group is staged

            // ---- phase R: staged rows -> global (lane = column pair) ----
            auto process_row = [&](const int i, float2 (&rslot)[NP]) {
                const int m = __shfl_sync(0xffffffffu, my_m, i);
                const int r32 = __shfl_sync(0xffffffffu, my_r32, i);
                const int r16 = __shfl_sync(0xffffffffu, my_r16, i);
                const bool valid = m >= 0;                     // predicate the stores, keep one basic block
                const float* srow = stg + (size_t)(half * RPW + i) * Cfg::SROW;
                float2 v[NP];
#pragma unroll
                for (int k = 0; k < NP; ++k) {
                    v[k] = *reinterpret_cast<const float2*>(srow + 64 * k + 2 * lane);
                    if (EPI == E_GENERIC) {
                        v[k].x = apply_act(v[k].x + bia[k].x, act);
                        v[k].y = apply_act(v[k].y + bia[k].y, act);
                    } else {
                        v[k].x = actf<ACT>(v[k].x + bia[k].x);
                        v[k].y = actf<ACT>(v[k].y + bia[k].y);
                    }
                }
                if (has_res) {
                    const float* rr = g.res + (size_t)r32 * g.ld32 + n0 + 2 * lane;
#pragma unroll
                    for (int k = 0; k < NP; ++k) {
                        const float2 t = kPrefetch ? rslot[k] : __ldg(reinterpret_cast<const float2*>(rr + 64 * k));
                        v[k].x = v[k].x * g.res_scale + t.x;
                        v[k].y = v[k].y * g.res_scale + t.y;
                    }
                    if (kPrefetch) {                           // request row i of the next tile (unconditionally, see process_rows)
                        const int rx = __shfl_sync(0xffffffffu, nx_r32, i);
                        const float* rn = g.res + (size_t)rx * g.ld32 + n0x + 2 * lane;
#pragma unroll
                        for (int k = 0; k < NP; ++k)
                            asm volatile("ld.global.v2.f32 {%0, %1}, [%2];" : "=f"(rslot[k].x), "=f"(rslot[k].y) : "l"(rn + 64 * k));
                    }
                }
                if ((EPI == E_GENERIC || kRes) && g.out32 && valid) {
                    float* oo = g.out32 + (size_t)r32 * g.ld32 + n0 + 2 * lane;
#pragma unroll
                    for (int k = 0; k < NP; ++k) *reinterpret_cast<float2*>(oo + 64 * k) = v[k];
                }
                if (has_ln) {
                    float sm = 0.f;
#pragma unroll
                    for (int k = 0; k < NP; ++k) sm += v[k].x + v[k].y;     // pad columns are exactly 0
                    const float mean = warp_sum(sm) * inv_c;
                    float q = 0.f;
#pragma unroll
                    for (int k = 0; k < NP; ++k) {
                        if (64 * k + 2 * lane < g.ln_C) {
                            const float a0 = v[k].x - mean, a1 = v[k].y - mean;
                            q += a0 * a0 + a1 * a1;
                        }
                    }
                    const float rstd = rsqrtf(warp_sum(q) * inv_c + 1e-5f);
                    uint16_t* oo = g.out16 + (size_t)r16 * g.ld16 + 2 * lane;
#pragma unroll
                    for (int k = 0; k < NP; ++k) {
                        const bool in = 64 * k + 2 * lane < g.ln_C;
                        const float y0 = in ? (v[k].x - mean) * rstd * lng[k].x + lnb[k].x : lnb[k].x;
                        const float y1 = in ? (v[k].y - mean) * rstd * lng[k].y + lnb[k].y : lnb[k].y;
                        if (valid)
                            *reinterpret_cast<uint32_t*>(oo + 64 * k) =
                                EPI == E_GENERIC ? pack2(y0, y1, g.out16_dtype) : packf<DT>(y0, y1);
                    }
                } else if (!valid) {
                } else if (EPI == E_PIXSHUF || (EPI == E_GENERIC && g.out16 && g.out16_mode == SRK_O16_PIXSHUF2)) {
                    // PixelShuffle(2): a 64-column group never straddles a sub-pixel (N/4 % 64 == 0)
#pragma unroll
                    for (int k = 0; k < NP; ++k)
                        *reinterpret_cast<uint32_t*>(g.out16 + off16_of(g, m, n0 + 64 * k) + 2 * lane) =
                            EPI == E_GENERIC ? pack2(v[k].x, v[k].y, g.out16_dtype) : packf<DT>(v[k].x, v[k].y);
                } else if (EPI == E_O16 || g.out16) {
                    uint16_t* oo = g.out16 + (size_t)m * g.ld16 + n0 + 2 * lane;
#pragma unroll
                    for (int k = 0; k < NP; ++k)
                        *reinterpret_cast<uint32_t*>(oo + 64 * k) =
                            EPI == E_GENERIC ? pack2(v[k].x, v[k].y, g.out16_dtype) : packf<DT>(v[k].x, v[k].y);
                }
                if (EPI == E_GENERIC && has_img) {
                    const int iy = __shfl_sync(0xffffffffu, my_iy, i), ix = __shfl_sync(0xffffffffu, my_ix, i);
                    const int ib = __shfl_sync(0xffffffffu, my_ib, i);
                    if (valid) {
#pragma unroll
                        for (int k = 0; k < NP; ++k) {
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const int oy = iy + ci[k][e], ox = ix + cj[k][e];
                                if (ci[k][e] >= 0 && oy < g.img_hc && ox < g.img_wc)
                                    g.img[(size_t)ib + (size_t)oy * g.img_wc + ox] = (e ? v[k].y : v[k].x) * g.img_scale;
                            }
                        }
                    }
                }
            };
            // E_RES_LN: RB rows per step, their LayerNorm butterflies interleaved stage by stage (the
            // 2 x 5 dependent shuffles of one row are the longest latency chain of phase R)
            constexpr int RB = SRK_LN_ROW_BATCH;                   // rows whose butterflies are interleaved
            auto process_rows = [&](const int i0) {
                int m[RB], r32[RB], r16[RB];
                float2 v[RB][NP];
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    m[r] = __shfl_sync(0xffffffffu, my_m, i0 + r);
                    r32[r] = __shfl_sync(0xffffffffu, my_r32, i0 + r);
                    r16[r] = __shfl_sync(0xffffffffu, my_r16, i0 + r);
                    const float* srow = stg + (size_t)(half * RPW + i0 + r) * Cfg::SROW;
                    float2 (&rs)[NP] = resv[kPrefetch ? i0 + r : 0];
#pragma unroll
                    for (int k = 0; k < NP; ++k) {
                        v[r][k] = *reinterpret_cast<const float2*>(srow + 64 * k + 2 * lane);
                        v[r][k].x = actf<ACT>(v[r][k].x + bia[k].x) * g.res_scale + rs[k].x;
                        v[r][k].y = actf<ACT>(v[r][k].y + bia[k].y) * g.res_scale + rs[k].y;
                    }
                    // request row i of the next tile into the same registers
                    // (rows outside the problem carry r32 = 0: an unconditional load keeps the request in flight in the slot's
                    //  own registers; a predicated load + select makes the compiler wait for the data on the spot)
                    const int rx = __shfl_sync(0xffffffffu, nx_r32, i0 + r);
                    const float* rn = g.res + (size_t)rx * g.ld32 + n0x + 2 * lane;
#pragma unroll
                    for (int k = 0; k < NP; ++k)
                        asm volatile("ld.global.v2.f32 {%0, %1}, [%2];" : "=f"(rs[k].x), "=f"(rs[k].y) : "l"(rn + 64 * k));
                    if (m[r] >= 0) {
                        float* oo = g.out32 + (size_t)r32[r] * g.ld32 + n0 + 2 * lane;
#pragma unroll
                        for (int k = 0; k < NP; ++k) *reinterpret_cast<float2*>(oo + 64 * k) = v[r][k];
                    }
                }
                float sm[RB], mean[RB], q[RB];
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    sm[r] = 0.f;
#pragma unroll
                    for (int k = 0; k < NP; ++k) sm[r] += v[r][k].x + v[r][k].y;     // pad columns are exactly 0
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                    for (int r = 0; r < RB; ++r) sm[r] += __shfl_xor_sync(0xffffffffu, sm[r], o);
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    mean[r] = sm[r] * inv_c; q[r] = 0.f;
#pragma unroll
                    for (int k = 0; k < NP; ++k) {
                        if (64 * k + 2 * lane < g.ln_C) {
                            const float a0 = v[r][k].x - mean[r], a1 = v[r][k].y - mean[r];
                            q[r] += a0 * a0 + a1 * a1;
                        }
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                    for (int r = 0; r < RB; ++r) q[r] += __shfl_xor_sync(0xffffffffu, q[r], o);
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    const float rstd = rsqrtf(q[r] * inv_c + 1e-5f);
                    uint16_t* oo = g.out16 + (size_t)r16[r] * g.ld16 + 2 * lane;
#pragma unroll
                    for (int k = 0; k < NP; ++k) {
                        const bool in = 64 * k + 2 * lane < g.ln_C;
                        const float y0 = in ? (v[r][k].x - mean[r]) * rstd * lng[k].x + lnb[k].x : lnb[k].x;
                        const float y1 = in ? (v[r][k].y - mean[r]) * rstd * lng[k].y + lnb[k].y : lnb[k].y;
                        if (m[r] >= 0) *reinterpret_cast<uint32_t*>(oo + 64 * k) = packf<DT>(y0, y1);
                    }
                }
            };
            if constexpr (kPrefetch && EPI == E_RES_LN) {
#pragma unroll
                for (int i = 0; i < RPW; i += RB) process_rows(i);
            } else if constexpr (kPrefetch) {
#pragma unroll
                for (int i = 0; i < RPW; ++i) process_row(i, resv[i]);
            } else {
#pragma unroll 2
                for (int i = 0; i < RPW; ++i) process_row(i, resv[0]);
            }
            asm volatile("bar.sync %0, %1;" ::"r"(1 + lg), "n"(32 * WPL) : "memory");   // staging free for the next tile
            my_m = nx_m; my_r32 = nx_r32; my_r16 = nx_r16;
            img_of(my_m, my_iy, my_ix, my_ib);
        }
        }  // !kStage16
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
    }
}

// ---- host side -----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

int encode_map(CUtensorMap* map, int dtype, int rank, const void* ptr, const cuuint64_t* dims,
                      const cuuint64_t* strides_bytes, const cuuint32_t* box) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return fail(SRK_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult rc = enc(map, dtype == SRK_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                      (cuuint32_t)rank, const_cast<void*>(ptr), dims, strides_bytes, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) return fail(SRK_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)rc);
    return 0;
}

int num_sms() {
    static int n[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    int v = dev < 64 ? n[dev] : 0;
    if (!v) {
        v = 148;
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        if (dev < 64) n[dev] = v;
    }
    return v;
}

template <int BN, int EPI, int ACT, int DT>
static int launch_tc5(const CUtensorMap& ma, const CUtensorMap& mb, const TcParams& p_in, cudaStream_t st) {
    using Cfg = TcCfg<BN, EPI>;
    static bool attr_set[64] = {};
    if (first_use_on_device(attr_set)) SRK_CUDA(cudaFuncSetAttribute(gemm_tc5_kernel<BN, EPI, ACT, DT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      TC_SMEM_TOTAL));
    TcParams p = p_in;
    const int total = p.m_tiles * p.n_tiles;
    int grid = total < num_sms() ? total : num_sms();
    // shared-memory plan: keep the weight tile resident when it fits and every CTA sees >= 2 M tiles
    const int bres = p.nkb * Cfg::B_BYTES;
    size_t halo_bytes = 0;
    if (p.halo) {
        // conv halo mode: [resident B | -] [halo ring] [ring of weight k-blocks]
        p.bstat = (bres + 3 * p.halo_stride <= Cfg::AVAIL) && (num_sms() >= p.n_tiles) && (p.m_tiles * p.n_tiles >= 2 * num_sms());
        if (p.bstat) {
            grid = (num_sms() / p.n_tiles) * p.n_tiles;
            if (grid > total) grid = total;
            p.bres_bytes = bres;
            p.stages = 0; p.stage_bytes = Cfg::B_BYTES;
            p.a_stages = (Cfg::AVAIL - bres) / p.halo_stride;
            if (p.a_stages > 4) p.a_stages = 4;
        } else {
            p.bres_bytes = 0;
            p.stage_bytes = Cfg::B_BYTES;
            p.a_stages = p.g.cpb >= 2 ? 3 : 2;                      // halo ring first, the rest goes to the weight ring
            p.stages = (Cfg::AVAIL - p.a_stages * p.halo_stride) / p.stage_bytes;
            if (p.stages > Cfg::MAX_STAGES) p.stages = Cfg::MAX_STAGES;
            if (p.stages < 2) return fail(SRK_ERR_UNSUPPORTED, "gemm(tcgen05): conv halo mode does not fit shared memory");
        }
        halo_bytes = (size_t)p.a_stages * p.halo_stride;
    } else {
        p.a_stages = 0; p.halo_stride = 0;
        p.bstat = (bres + 3 * Cfg::A_BYTES <= Cfg::AVAIL) && (num_sms() >= p.n_tiles) &&
                  (EPI == E_ATTN || p.m_tiles * p.n_tiles >= 2 * num_sms());
        if (EPI == E_ATTN && !p.bstat) return fail(SRK_ERR_UNSUPPORTED, "gemm(tcgen05): fused attention needs the resident-weight mode");
        if (p.bstat) {
            grid = (num_sms() / p.n_tiles) * p.n_tiles;             // multiple of n_tiles: nt is fixed per CTA
            if (grid > total) grid = total;                         // total is a multiple of n_tiles too
            p.bres_bytes = bres;
            p.stage_bytes = Cfg::A_BYTES;
        } else {
            p.bres_bytes = 0;
            p.stage_bytes = Cfg::STAGE_BYTES;
        }
        p.stages = (Cfg::AVAIL - p.bres_bytes) / p.stage_bytes;
        if (p.stages > Cfg::MAX_STAGES) p.stages = Cfg::MAX_STAGES;
    }
    const size_t smem = (size_t)p.bres_bytes + halo_bytes + (size_t)p.stages * p.stage_bytes + Cfg::STAGING_BYTES + 1024 + Cfg::AUX_BYTES;
    gemm_tc5_kernel<BN, EPI, ACT, DT><<<grid, Cfg::THREADS, smem, st>>>(ma, mb, p);
    SRK_LAUNCH_CHECK("gemm_tc5_kernel");
    return 0;
}

// picks the specialised epilogue when the call matches one, else the generic kernel
template <int BN>
static int dispatch_tc5(const srk_gemm_args* a, const CUtensorMap& ma, const CUtensorMap& mb, const TcParams& p,
                        cudaStream_t st) {
    const bool res = a->res != nullptr, o32 = a->out32 != nullptr, o16 = a->out16 != nullptr;
    const bool ln = a->ln_g != nullptr, img = a->img != nullptr;
    const bool ps = o16 && a->out16_mode == SRK_O16_PIXSHUF2;
    const int dt = a->out16_dtype, act = a->act;
    if (!img && !res && !o32 && o16 && !ln && !ps) {
        if (act == SRK_ACT_NONE && dt == SRK_BF16) return launch_tc5<BN, E_O16, SRK_ACT_NONE, SRK_BF16>(ma, mb, p, st);
        if (act == SRK_ACT_NONE && dt == SRK_FP16) return launch_tc5<BN, E_O16, SRK_ACT_NONE, SRK_FP16>(ma, mb, p, st);
        if (act == SRK_ACT_GELU && dt == SRK_BF16) return launch_tc5<BN, E_O16, SRK_ACT_GELU, SRK_BF16>(ma, mb, p, st);
        if (act == SRK_ACT_LRELU && dt == SRK_FP16) return launch_tc5<BN, E_O16, SRK_ACT_LRELU, SRK_FP16>(ma, mb, p, st);
        if (act == SRK_ACT_RELU && dt == SRK_FP16) return launch_tc5<BN, E_O16, SRK_ACT_RELU, SRK_FP16>(ma, mb, p, st);
        if (act == SRK_ACT_LRELU02 && dt == SRK_FP16) return launch_tc5<BN, E_O16, SRK_ACT_LRELU02, SRK_FP16>(ma, mb, p, st);
    }
    if (!img && res && ln && act == SRK_ACT_NONE && dt == SRK_BF16)
        return launch_tc5<BN, E_RES_LN, SRK_ACT_NONE, SRK_BF16>(ma, mb, p, st);
    if (!img && res && ln && act == SRK_ACT_NONE && dt == SRK_FP16)
        return launch_tc5<BN, E_RES_LN, SRK_ACT_NONE, SRK_FP16>(ma, mb, p, st);
    if (!img && res && !ln && !ps && act == SRK_ACT_NONE && (!o16 || dt == SRK_FP16))
        return launch_tc5<BN, E_RES, SRK_ACT_NONE, SRK_FP16>(ma, mb, p, st);
    if (!img && !res && !o32 && ps && act == SRK_ACT_NONE && dt == SRK_FP16)
        return launch_tc5<BN, E_PIXSHUF, SRK_ACT_NONE, SRK_FP16>(ma, mb, p, st);
    if (!img && !res && !o32 && ps && act == SRK_ACT_LRELU02 && dt == SRK_FP16)
        return launch_tc5<BN, E_PIXSHUF, SRK_ACT_LRELU02, SRK_FP16>(ma, mb, p, st);
    return launch_tc5<BN, E_GENERIC, 0, 0>(ma, mb, p, st);
}

bool g_conv_halo = true;     // srk_gemm_conv_halo(): tests compare the halo-tile convs with the per-tap-box convs

int gemm_tcgen05(const srk_gemm_args* a, cudaStream_t st) {
    SRK_REQUIRE(a->N % 64 == 0, "gemm(tcgen05): N=%d must be a multiple of 64", a->N);
    SRK_REQUIRE(((uintptr_t)a->A & 15) == 0 && ((uintptr_t)a->Wt & 15) == 0, "gemm(tcgen05): operands must be 16 B aligned");
    TcParams p{};
    p.g = make_gemm_params(a);
    const bool ps256 = a->out16 && a->out16_mode == SRK_O16_PIXSHUF2 && !a->res && !a->out32 && !a->img && !a->ln_g &&
                       (a->act == SRK_ACT_NONE || a->act == SRK_ACT_LRELU02) && a->out16_dtype == SRK_FP16 && a->N % 256 == 0;
    const int BN = ps256 ? 256 : (a->N % 192 == 0 ? 192 : (a->N % 128 == 0 ? 128 : 64));
    if (a->ln_g) {
        SRK_REQUIRE(a->N == BN, "gemm(tcgen05): fused LayerNorm needs the whole row in one tile (N=%d)", a->N);
        SRK_REQUIRE(!a->ln_pad_one || (a->ln_C % 2 == 0 && a->ln_C + 2 <= a->N), "gemm(tcgen05): ln_pad_one needs two pad columns");
        SRK_REQUIRE(a->out16 && a->out16_mode == SRK_O16_ROWS && a->ln_b && a->ln_C > 0 && a->ln_C <= a->N,
                    "gemm(tcgen05): bad fused LayerNorm arguments");
    }
    if (a->out16 && a->out16_mode == SRK_O16_PIXSHUF2)
        SRK_REQUIRE((a->N / 4) % 64 == 0, "gemm(tcgen05): pixel-shuffle output needs N/4 %% 64 == 0");
    if (a->res || a->out32) SRK_REQUIRE(a->ld32 % 4 == 0, "gemm(tcgen05): ld32 %% 4");
    p.n_tiles = a->N / BN;
    p.nkb = a->K / TBK;
    const bool attn = a->attn_table != nullptr;
    if (attn) SRK_REQUIRE(BN == 192 && a->N == p.n_tiles * 192, "gemm(tcgen05): fused attention needs N == pairs * 192");
    CUtensorMap ma, mb;
    if (a->a_mode == SRK_A_CONV3X3 && BN == 64 && a->lda == 64 && g_conv_halo) {
        // halo mode (64 -> 64 channel convs: one channel block, measured 34 vs 38 us for 3x3, 56 vs 58 us for 5x5; with three channel
        // blocks, 192 -> 64, the per-tap boxes win 50 vs 52 us): 8 x 16 pixel tiles; ONE box (64 ch, 8 + 2r, 16 + 2r) serves every tap
        const int r = a->conv_k == 5 ? 2 : 1;
        p.halo = 1; p.halo_r = r;
        p.conv_bw = 8; p.conv_bh = 16;
        p.halo_box_bytes = (8 + 2 * r) * (16 + 2 * r) * 128;
        p.halo_stride = (int)align_up((size_t)p.halo_box_bytes, 1024);
        p.conv_tx = (a->W + 7) / 8;
        p.conv_ty = (a->H + 15) / 16;
        p.m_tiles = a->nB * p.conv_tx * p.conv_ty;
        cuuint64_t dims[4] = {(cuuint64_t)a->lda, (cuuint64_t)a->W, (cuuint64_t)a->H, (cuuint64_t)a->nB};
        cuuint64_t strides[3] = {(cuuint64_t)a->lda * 2, (cuuint64_t)a->W * a->lda * 2, (cuuint64_t)a->H * a->W * a->lda * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)(8 + 2 * r), (cuuint32_t)(16 + 2 * r), 1};
        if (int rc = encode_map(&ma, a->dtype, 4, a->A, dims, strides, box)) return rc;
    } else if (a->a_mode == SRK_A_CONV3X3) {
        // pick the 128-pixel box (BW x BH) that wastes the fewest out-of-image pixels
        long best = -1;
        for (int bw = 8; bw <= 128; bw *= 2) {
            const int bh = 128 / bw;
            const long cover = (long)((a->W + bw - 1) / bw) * bw * (long)((a->H + bh - 1) / bh) * bh;
            if (best < 0 || cover < best) { best = cover; p.conv_bw = bw; p.conv_bh = bh; }
        }
        p.conv_tx = (a->W + p.conv_bw - 1) / p.conv_bw;
        p.conv_ty = (a->H + p.conv_bh - 1) / p.conv_bh;
        p.m_tiles = a->nB * p.conv_tx * p.conv_ty;
        cuuint64_t dims[4] = {(cuuint64_t)a->lda, (cuuint64_t)a->W, (cuuint64_t)a->H, (cuuint64_t)a->nB};
        cuuint64_t strides[3] = {(cuuint64_t)a->lda * 2, (cuuint64_t)a->W * a->lda * 2, (cuuint64_t)a->H * a->W * a->lda * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)p.conv_bw, (cuuint32_t)p.conv_bh, 1};
        if (int rc = encode_map(&ma, a->dtype, 4, a->A, dims, strides, box)) return rc;
    } else {
        p.m_tiles = ceil_div(a->M, TBM);
        cuuint64_t dims[2] = {(cuuint64_t)a->K, (cuuint64_t)a->M};
        cuuint64_t strides[1] = {(cuuint64_t)a->lda * 2};
        cuuint32_t box[2] = {64, 128};
        if (int rc = encode_map(&ma, a->dtype, 2, a->A, dims, strides, box)) return rc;
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)a->K, (cuuint64_t)a->N};
        cuuint64_t strides[1] = {(cuuint64_t)a->K * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)(attn ? 64 : BN)};
        if (int rc = encode_map(&mb, a->dtype, 2, a->Wt, dims, strides, box)) return rc;
    }
    if (attn) return launch_tc5<192, E_ATTN, SRK_ACT_NONE, SRK_BF16>(ma, mb, p, st);
    switch (BN) {
        case 256: return a->act == SRK_ACT_LRELU02 ? launch_tc5<256, E_PIXSHUF, SRK_ACT_LRELU02, SRK_FP16>(ma, mb, p, st)
                                                   : launch_tc5<256, E_PIXSHUF, SRK_ACT_NONE, SRK_FP16>(ma, mb, p, st);
        case 192: return dispatch_tc5<192>(a, ma, mb, p, st);
        case 128: return dispatch_tc5<128>(a, ma, mb, p, st);
        default: return dispatch_tc5<64>(a, ma, mb, p, st);
    }
}

}  // namespace srk

extern "C" int srk_gemm_conv_halo(int on) { srk::g_conv_halo = on != 0; return 0; }
