// Error plumbing, device checks and the integer index-map export of libsrk.
#include "common.cuh"
#include <string.h>
#include <vector>

namespace srk {
thread_local char g_err[512] = "";
thread_local long long g_launches = 0;
static int g_engine = SRK_ENGINE_TCGEN05;
thread_local bool g_prof_on = false;
struct ProfRec { int family; cudaEvent_t a, b; };
static thread_local std::vector<ProfRec> g_prof;
void prof_begin(int family, cudaStream_t st) {
    ProfRec r{family, nullptr, nullptr};
    cudaEventCreate(&r.a); cudaEventCreate(&r.b);
    cudaEventRecord(r.a, st);
    g_prof.push_back(r);
}
void prof_end(cudaStream_t st) { if (!g_prof.empty()) cudaEventRecord(g_prof.back().b, st); }

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

__global__ void index_map_kernel(int kind, int H, int W, int shift, int a, int32_t* out, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (kind == 0) {
        out[i] = win_pos_to_token((int)i, H, W, shift);
    } else if (kind == 1) {
        const int win = (int)(i >> 12), qi = (int)((i >> 6) & 63), kj = (int)(i & 63);
        out[i] = shift > 0 &&
                 win_pos_label(win, qi, H, W, shift) != win_pos_label(win, kj, H, W, shift);
    } else if (kind == 2) {
        out[i] = rel_pos_index((int)(i >> 6), (int)(i & 63));
    } else {
        // pixel shuffle: output element (c, oy, ox) of (C, H*r, W*r) <- input (c*r*r + (oy%r)*r + ox%r, oy/r, ox/r)
        const int r = shift, Wo = W * r, Ho = H * r;
        const int ox = (int)(i % Wo), oy = (int)((i / Wo) % Ho), c = (int)(i / ((long long)Wo * Ho));
        const int ch = c * r * r + (oy % r) * r + (ox % r);
        out[i] = (ch * H + oy / r) * W + ox / r;
    }
}
}  // namespace srk

using namespace srk;

extern "C" const char* srk_last_error(void) { return g_err; }
extern "C" int srk_version(void) { return SRK_VERSION; }
extern "C" long long srk_launch_count(int reset) {
    long long v = g_launches;
    if (reset) g_launches = 0;
    return v;
}
extern "C" int srk_profile(int enable) { g_prof_on = enable != 0; return 0; }
extern "C" int srk_profile_read(double* ms, long long* calls, int reset) {
    SRK_REQUIRE(ms && calls, "profile_read: null pointer");
    for (int i = 0; i < SRK_PROF_N; ++i) { ms[i] = 0.0; calls[i] = 0; }
    for (auto& r : g_prof) {
        SRK_CUDA(cudaEventSynchronize(r.b));
        float t = 0.f;
        SRK_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
        if (r.family >= 0 && r.family < SRK_PROF_N) { ms[r.family] += t; calls[r.family]++; }
    }
    if (reset) {
        for (auto& r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
        g_prof.clear();
    }
    return 0;
}
extern "C" int srk_set_engine(int engine) {
    if (engine != SRK_ENGINE_TCGEN05 && engine != SRK_ENGINE_MMA_SYNC)
        return fail(SRK_ERR_INVALID, "unknown engine %d", engine);
    g_engine = engine;
    return 0;
}
extern "C" int srk_get_engine(void) { return g_engine; }

extern "C" int srk_check_device(int device) {
    cudaDeviceProp prop;
    SRK_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(SRK_ERR_ARCH, "device %d is sm_%d%d; libsrk is built for sm_100a only", device,
                    prop.major, prop.minor);
    return 0;
}

extern "C" int srk_index_map(int kind, int H, int W, int shift, int a, int32_t* out, void* stream) {
    SRK_REQUIRE(out != nullptr, "index_map: null output");
    long long n = 0;
    if (kind == 0 || kind == 1) {
        SRK_REQUIRE(H > 0 && W > 0 && H % 8 == 0 && W % 8 == 0, "index_map: H,W must be multiples of 8");
        SRK_REQUIRE(shift == 0 || shift == 4, "index_map: shift must be 0 or 4");
        n = (long long)H * W * (kind == 1 ? 64 : 1);
    } else if (kind == 2) {
        n = 64 * 64;
    } else if (kind == 3) {
        SRK_REQUIRE(a > 0 && H > 0 && W > 0 && shift > 0, "index_map: bad pixel-shuffle shape");
        n = (long long)a * H * shift * W * shift;
    } else {
        return fail(SRK_ERR_INVALID, "index_map: unknown kind %d", kind);
    }
    index_map_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(kind, H, W, shift, a, out, n);
    SRK_LAUNCH_CHECK("index_map_kernel");
    return 0;
}

extern "C" int srk_gemm(const srk_gemm_args* g, void* stream) {
    if (int rc = validate_gemm(g)) return rc;
    ProfScope ps(g->res && g->ln_g && g->a_mode == SRK_A_ROWS ? SRK_PROF_GEMM_RES_LN : SRK_PROF_GEMM, stream);
    if (g_engine == SRK_ENGINE_MMA_SYNC) return gemm_mma_sync(g, (cudaStream_t)stream);
    return gemm_tcgen05(g, (cudaStream_t)stream);
}
