// Warp-level 8x8-window attention on a window staged in shared memory (shared by the stand-alone
// attention kernel and by the fused qkv-GEMM + attention epilogue of the tcgen05 engine).
// Restates WindowAttention.forward (dlib/models/network_swinir.py:150-176) and the shift mask of
// calculate_mask (:260-285).
#pragma once
#include "common.cuh"

namespace srk {

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0,
                                         uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
        "{%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src));
}

__device__ __forceinline__ uint32_t packbf(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}


// One (16-query strip, head) unit of 8x8-window attention on a window staged in shared memory.
//   rows_s / rows : shared-memory address / pointer of the window's 64 staged rows (stride RS bytes)
//   qc, kc, vc    : element column of this head's q / k / v inside a row
//   tb            : this head's 225-entry relative-position-bias table (shared memory)
//   lab, masked   : region labels of the window's 64 positions (shift mask), and whether any differ
// S / P / O live in registers (mma.sync m16n8k16, bf16 in, fp32 accumulate); O (normalised) is
// written over the strip's own q slot, or (AO_OUT) straight into a 128-row K-major 128B-swizzled UMMA
// operand tile `ao` (k-blocks of 16 KB: row r, column head * DP + d), rows ao_row0 + window position.
template <int DP, bool AO_OUT = false, typename LT = int>
__device__ __forceinline__ void attn_unit(uint32_t rows_s, unsigned char* rows, int RS, int strip, int qc, int kc,
                                          int vc, const float* tb, const LT* lab, bool masked, float scale, int lane,
                                          unsigned char* ao = nullptr, int ao_row0 = 0, int head = 0) {
    const int g = lane >> 2, t = lane & 3;
    const int r0 = strip * 16;
    const float LOG2E = 1.4426950408889634f;
    const uint32_t a_addr = rows_s + (r0 + (lane & 7) + ((lane >> 3) & 1) * 8) * RS + (lane >> 4) * 16;
    const uint32_t k_addr = rows_s + ((lane & 7) + (lane >> 4) * 8) * RS + ((lane >> 3) & 1) * 16;
    const uint32_t v_addr = rows_s + ((lane & 7) + ((lane >> 3) & 1) * 8) * RS + (lane >> 4) * 16;
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
    // ---- S = Q K^T ----
#pragma unroll
    for (int ks = 0; ks < DP / 16; ++ks) {
        uint32_t a[4];
        ldsm_x4(a, a_addr + (qc + ks * 16) * 2);
#pragma unroll
        for (int np = 0; np < 4; ++np) {
            uint32_t b[4];
            ldsm_x4(b, k_addr + np * 16 * RS + (kc + ks * 16) * 2);
            mma_bf16(s[2 * np], a, b[0], b[1]);
            mma_bf16(s[2 * np + 1], a, b[2], b[3]);
        }
    }
    // ---- + bias + mask, softmax ----
    // rel_pos_index(i, j) with i = r0 + g (+8), j = nt*8 + 2t + e collapses to
    //   base - 15*nt - e   (+15 for the second row): compile-time offsets from one pointer
    const float* bp = tb + (2 * strip + 7) * 15 + (g - 2 * t + 7);
    const int i0 = r0 + g, i1 = r0 + g + 8;
    float m0 = -3.0e38f, m1 = -3.0e38f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            s[nt][e] = fmaf(s[nt][e], scale, bp[-15 * nt - e]);
            s[nt][2 + e] = fmaf(s[nt][2 + e], scale, bp[15 - 15 * nt - e]);
        }
    }
    if (masked) {                                              // only the last window row / column
        const int l0 = lab[i0], l1 = lab[i1];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int lj = lab[nt * 8 + 2 * t + e];
                if (l0 != lj) s[nt][e] += -100.f;
                if (l1 != lj) s[nt][2 + e] += -100.f;
            }
        }
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        m0 = fmaxf(m0, fmaxf(s[nt][0], s[nt][1]));
        m1 = fmaxf(m1, fmaxf(s[nt][2], s[nt][3]));
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    const float mb0 = m0 * LOG2E, mb1 = m1 * LOG2E;
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const float p0 = ex2_approx(fmaf(s[nt][e], LOG2E, -mb0));
            const float p1 = ex2_approx(fmaf(s[nt][2 + e], LOG2E, -mb1));
            s[nt][e] = p0; s[nt][2 + e] = p1;
            sum0 += p0; sum1 += p1;
        }
    }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
    const float inv0 = rcp_approx(sum0), inv1 = rcp_approx(sum1);
    // ---- O = P V  (P in [0,1] is rounded to bf16 un-normalised; 1/sum is applied to O) ----
    float o[DP / 8][4];
#pragma unroll
    for (int i = 0; i < DP / 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        uint32_t a[4];
        a[0] = packbf(s[2 * kk][0], s[2 * kk][1]);
        a[1] = packbf(s[2 * kk][2], s[2 * kk][3]);
        a[2] = packbf(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        a[3] = packbf(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
        for (int nd = 0; nd < DP / 16; ++nd) {
            uint32_t b[4];
            ldsm_x4_trans(b, v_addr + kk * 16 * RS + (vc + nd * 16) * 2);
            mma_bf16(o[2 * nd], a, b[0], b[1]);
            mma_bf16(o[2 * nd + 1], a, b[2], b[3]);
        }
    }
    if constexpr (AO_OUT) {
        static_assert(DP == 32, "the operand-tile output is built for a padded head dim of 32");
        // column head * 32 + nt * 8 + 2 t  ->  k-block head / 2, 16 B chunk (head & 1) * 4 + nt, byte 4 t inside it
        const int ra = ao_row0 + r0 + g, rb = ra + 8;
        unsigned char* pa = ao + (head >> 1) * 16384 + ra * 128 + 4 * t;
        unsigned char* pb = ao + (head >> 1) * 16384 + rb * 128 + 4 * t;
#pragma unroll
        for (int nt = 0; nt < DP / 8; ++nt) {
            const int ch = (head & 1) * 4 + nt;
            *reinterpret_cast<uint32_t*>(pa + ((ch ^ (ra & 7)) << 4)) = packbf(o[nt][0] * inv0, o[nt][1] * inv0);
            *reinterpret_cast<uint32_t*>(pb + ((ch ^ (rb & 7)) << 4)) = packbf(o[nt][2] * inv1, o[nt][3] * inv1);
        }
        return;
    }
    // ---- O overwrites this strip's own q slot (only this warp reads it, and it is done) ----
    __syncwarp();
#pragma unroll
    for (int nt = 0; nt < DP / 8; ++nt) {
        const int col = qc + nt * 8 + 2 * t;
        *reinterpret_cast<uint32_t*>(rows + (size_t)(r0 + g) * RS + col * 2) = packbf(o[nt][0] * inv0, o[nt][1] * inv0);
        *reinterpret_cast<uint32_t*>(rows + (size_t)(r0 + g + 8) * RS + col * 2) = packbf(o[nt][2] * inv1, o[nt][3] * inv1);
    }
}

}  // namespace srk
