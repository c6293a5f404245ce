#include "gemm_common.cuh"
namespace srk {
int gemm_tcgen05(const srk_gemm_args* g, cudaStream_t st) {
    return fail(SRK_ERR_UNSUPPORTED, "tcgen05 engine not built yet");
}
}
