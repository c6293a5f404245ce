// Microbenchmark (not part of libsrk.so): TMEM read throughput of tcgen05.ld for 4 / 8 / 16 warps per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu && ./tmem_bw
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(int iters, long long* cyc, uint32_t* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = slot + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    long long t0 = clock64();
    if (MODE == 0) {            // x32 load, wait after each
        for (int i = 0; i < iters; ++i) {
            uint32_t v[32];
            ld32(tm + (uint32_t)((i * 32) & 511), v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) acc ^= v[j];
        }
    } else if (MODE == 1) {     // 4 x32 loads in flight, one wait
        for (int i = 0; i < iters; i += 4) {
            uint32_t v0[32], v1[32], v2[32], v3[32];
            ld32(tm + 0, v0); ld32(tm + 32, v1); ld32(tm + 64, v2); ld32(tm + 96, v3);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) acc ^= v0[j] ^ v1[j] ^ v2[j] ^ v3[j];
        }
    } else {                    // x16 loads, wait after each
        for (int i = 0; i < iters; ++i) {
            uint32_t v[16];
            ld16(tm + (uint32_t)((i * 16) & 511), v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 16; ++j) acc ^= v[j];
        }
    }
    long long t1 = clock64();
    if (threadIdx.x % 32 == 0) cyc[blockIdx.x * 16 + warp] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
}
int main() {
    long long* cyc; uint32_t* sink;
    cudaMalloc(&cyc, 148 * 16 * 8); cudaMalloc(&sink, 148 * 512 * 4);
    const int iters = 4096;
    for (int mode = 0; mode < 3; ++mode)
        for (int warps = 4; warps <= 16; warps *= 2) {
            cudaMemset(cyc, 0, 148 * 16 * 8);
            if (mode == 0) k<0><<<148, warps * 32>>>(iters, cyc, sink);
            else if (mode == 1) k<1><<<148, warps * 32>>>(iters, cyc, sink);
            else k<2><<<148, warps * 32>>>(iters, cyc, sink);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[16];
            cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            long long mx = 0; for (int w = 0; w < warps; ++w) mx = h[w] > mx ? h[w] : mx;
            const double bytes = (double)warps * iters * 32 * (mode == 2 ? 16 : 32) * 4;
            printf("mode %d (%s) warps %2d: %lld cyc, %.1f B/clk/SM, %.1f clk per load-instr per warp  [%s]\n", mode,
                   mode == 0 ? "x32 wait each" : mode == 1 ? "4 x x32 in flight" : "x16 wait each", warps, mx, bytes / mx, (double)mx / iters,
                   cudaGetErrorString(e));
        }
    return 0;
}
