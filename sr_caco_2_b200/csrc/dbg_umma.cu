// Experiment (not part of libsrk.so: scripts/dbg_umma.py builds and runs it): can a tcgen05.mma A operand be a SHIFTED window of a larger
// 128B-swizzled tile?  A: (R, 64) fp16 rows loaded by ONE TMA box into a 1024 B aligned buffer; the MMA reads
// 128 rows as 16 core groups of 8 rows: group g starts at row  shift + g * pitch  (SBO = pitch * 128 B).
// D[r][n] = sum_k A[shift + (r / 8) * pitch + r % 8][k] * Bm[n][k]   is compared on the host.
#include "tc5_ptx.cuh"

namespace srk {

__global__ void __launch_bounds__(128, 1)
dbg_umma_shift_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, int R,
                      int shift, int pitch, int base_offset, float* __restrict__ D) {
    extern __shared__ unsigned char dbg_raw[];
    const uint32_t raw = smem_u32(dbg_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t sA = base, sB = base + 256 * 128, bars = sB + 64 * 128;
    const uint32_t full = bars, done = bars + 8, slot = bars + 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(full, 1); mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(dbg_raw + (slot - raw));
    if (threadIdx.x == 0) {
        mbar_expect_tx(full, (uint32_t)(R * 128 + 64 * 128));
        tma_load_2d(sA, &map_a, full, 0, 0);
        tma_load_2d(sB, &map_b, full, 0, 0);
        mbar_wait(full, 0);
        tc_fence_after();
        uint64_t da = 0;
        const uint32_t a0 = sA + (uint32_t)shift * 128u;
        da |= (uint64_t)((a0 >> 4) & 0x3FFF);
        da |= (uint64_t)1 << 16;
        da |= (uint64_t)((pitch * 128) >> 4) << 32;
        da |= (uint64_t)1 << 46;
        da |= (uint64_t)(base_offset & 7) << 49;
        da |= (uint64_t)2 << 61;
        const uint64_t db = umma_desc_sw128(sB);
        const uint32_t idesc = umma_idesc(0, 128, 64);
        for (int k = 0; k < 4; ++k) tc_mma_f16(tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, k ? 1u : 0u);
        tc_commit(done);
    }
    mbar_wait(done, 0);
    tc_fence_after();
    uint32_t v[32];
    for (int c = 0; c < 2; ++c) {
        tc_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 32), v);
        for (int j = 0; j < 32; ++j) D[(size_t)(warp * 32 + lane) * 64 + c * 32 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

}  // namespace srk

using namespace srk;

// experiment entry (scripts/dbg_umma.py)
extern "C" int srk_dbg_umma_shift(const void* A, int R, const void* Bm, int shift, int pitch, int base_offset, float* D,
                                  void* stream) {
    SRK_REQUIRE(A && Bm && D && R > 0 && R <= 256, "dbg_umma: bad arguments");
    CUtensorMap ma, mb;
    {
        cuuint64_t dims[2] = {64, (cuuint64_t)R};
        cuuint64_t strides[1] = {128};
        cuuint32_t box[2] = {64, (cuuint32_t)R};
        if (int rc = encode_map(&ma, SRK_FP16, 2, A, dims, strides, box)) return rc;
    }
    {
        cuuint64_t dims[2] = {64, 64};
        cuuint64_t strides[1] = {128};
        cuuint32_t box[2] = {64, 64};
        if (int rc = encode_map(&mb, SRK_FP16, 2, Bm, dims, strides, box)) return rc;
    }
    const size_t smem = 256 * 128 + 64 * 128 + 64 + 1024;
    SRK_CUDA(cudaFuncSetAttribute(dbg_umma_shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dbg_umma_shift_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(ma, mb, R, shift, pitch, base_offset, D);
    SRK_LAUNCH_CHECK("dbg_umma_shift_kernel");
    return 0;
}
