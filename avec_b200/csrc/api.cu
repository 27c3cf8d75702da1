// Library-level entry points: status strings, diagnostics and the avec_gemm dispatcher.
#include "common.cuh"
#include <atomic>
#include <cstdlib>

static thread_local int g_last_cuda_error = 0;
// process-wide: the autograd backward of the host mirror runs on PyTorch's backward thread, the forward on the caller's
static std::atomic<long long> g_launches{0};

void avec_set_last_cuda_error(int e) { g_last_cuda_error = e; }
void avec_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// ---- programmatic dependent launch policy (see common.cuh avec_launch_pdl) ---------------------------------------------------
static int g_pdl_enabled = 1;
static cudaStream_t g_pdl_excluded[8];
static int g_pdl_nexcluded = 0;
extern "C" void avec_set_pdl(int enabled) { g_pdl_enabled = enabled; }
extern "C" void avec_pdl_exclude_stream(avec_stream_t stream, int enabled) {
    if (!enabled) { g_pdl_nexcluded = 0; return; }
    cudaStream_t st = as_stream(stream);
    for (int i = 0; i < g_pdl_nexcluded; ++i) if (g_pdl_excluded[i] == st) return;
    if (g_pdl_nexcluded < 8) g_pdl_excluded[g_pdl_nexcluded++] = st;
}
bool avec_pdl_for_stream(cudaStream_t st) {
    static int env = -1;
    if (env < 0) { const char* e = getenv("AVEC_PDL"); env = e ? atoi(e) : 1; }
    if (!env || !g_pdl_enabled) return false;
    for (int i = 0; i < g_pdl_nexcluded; ++i) if (g_pdl_excluded[i] == st) return false;
    return true;
}
// AVEC_PDL=1: the tcgen05 GEMM only; 2 (default): also the small latency-bound kernels around it
bool avec_pdl_small_kernels() {
    static int env = -1;
    if (env < 0) { const char* e = getenv("AVEC_PDL"); env = e ? atoi(e) : 2; }
    return env >= 2;
}

int avec_gemm_simt(const avec_gemm_args* a, cudaStream_t st);
int avec_gemm_tc(const avec_gemm_args* a, cudaStream_t st);       // gemm_tc.cu
bool avec_gemm_tc_supported(const avec_gemm_args* a);             // gemm_tc.cu

extern "C" const char* avec_strerror(int status) {
    switch (status) {
    case AVEC_OK: return "ok";
    case AVEC_ERR_INVALID: return "invalid argument or unsupported shape";
    case AVEC_ERR_LAUNCH: return "CUDA launch failed (see avec_last_cuda_error)";
    case AVEC_ERR_UNSUPPORTED: return "combination not implemented";
    case AVEC_ERR_DRIVER: return "CUDA driver entry point unavailable or tensor-map encode failed";
    default: return "unknown avec status";
    }
}
extern "C" int avec_last_cuda_error(void) { return g_last_cuda_error; }
extern "C" int avec_version(void) { return 100; }
extern "C" long long avec_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
extern "C" void avec_reset_launch_count(void) { g_launches.store(0, std::memory_order_relaxed); }

extern "C" int avec_gemm(const avec_gemm_args* a, avec_stream_t stream) {
    AVEC_CHECK_ARG(a && a->A && a->B && a->out && a->M > 0 && a->N > 0 && a->K > 0);
    AVEC_CHECK_ARG(a->epi >= AVEC_EPI_LINEAR && a->epi <= AVEC_EPI_RELU);
    AVEC_CHECK_ARG(a->epi != AVEC_EPI_ACCUM || a->out_dtype == AVEC_F32);
    AVEC_CHECK_ARG((a->epi != AVEC_EPI_RESIDUAL && a->epi != AVEC_EPI_DSWISH) || a->aux);
    cudaStream_t st = as_stream(stream);
    const bool drop = a->drop_p > 0.0f;
    AVEC_CHECK_ARG(!drop || (a->drop_rng && a->drop_p < 1.0f && a->epi != AVEC_EPI_ACCUM && a->epi != AVEC_EPI_RELU && !a->colstats));
    if (drop && (a->impl == AVEC_IMPL_SIMT || !avec_gemm_tc_supported(a))) return AVEC_ERR_UNSUPPORTED;   // caller: separate avec_dropout
    if (a->impl == AVEC_IMPL_SIMT) return avec_gemm_simt(a, st);
    if (a->impl == AVEC_IMPL_TCGEN05) {
        if (!avec_gemm_tc_supported(a)) return AVEC_ERR_UNSUPPORTED;
        return avec_gemm_tc(a, st);
    }
    if (avec_gemm_tc_supported(a)) return avec_gemm_tc(a, st);
    return avec_gemm_simt(a, st);
}
