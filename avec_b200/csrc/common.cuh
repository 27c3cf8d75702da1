// Shared device/host helpers for the avec_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <string.h>
#include "../../include/avec_b200.h"

#define AVEC_CHECK_ARG(cond) do { if (!(cond)) return AVEC_ERR_INVALID; } while (0)
#define AVEC_LAUNCH_CHECK() do { avec_count_launch(); cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) { avec_set_last_cuda_error((int)e__); return AVEC_ERR_LAUNCH; } } while (0)

void avec_set_last_cuda_error(int e);
void avec_count_launch();

bool avec_pdl_for_stream(cudaStream_t st);     // api.cu
bool avec_pdl_small_kernels();                  // api.cu

static inline cudaStream_t as_stream(avec_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline long long cdivll(long long a, long long b) { return (a + b - 1) / b; }

typedef __nv_bfloat16 bf16;

// Launch `kernel` as a programmatic dependent of the kernel in front of it on the stream when the policy allows (`force`, or a
// grid of at most two waves of threads: parked CTAs of a large grid would take registers and thread slots
// from kernels of other streams), else the ordinary way.  EVERY kernel launched through this helper executes pdl_wait() before its
// first global-memory access - that is what keeps the stream's ordering intact.
template <typename... KArgs, typename... Args>
static inline void avec_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool force, Args&&... args) {
    const unsigned long long threads = (unsigned long long)grid.x * grid.y * grid.z * block.x * block.y * block.z;
    if (avec_pdl_for_stream(st) && (force || (avec_pdl_small_kernels() && threads <= 148ull * 4096ull))) {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = grid;
        cfg.blockDim = block;
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
    } else {
        kernel<<<grid, block, smem, st>>>(static_cast<KArgs>(args)...);
    }
}

// BatchNorm column statistics are accumulated into one of AVEC_STATS_REPLICAS copies (selected by CTA index) to spread the
// same-address L2 atomics; avec_bn_finalize sums the copies.
#define AVEC_STATS_REPLICAS 32

__device__ __forceinline__ float ldf(const float* p) { return *p; }
__device__ __forceinline__ float ldf(const bf16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void stf(float* p, float v) { *p = v; }
__device__ __forceinline__ void stf(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

// runtime-dtype element access (used by epilogues where templating every combination is not worth it)
__device__ __forceinline__ float ld_any(const void* p, int dtype, size_t idx) {
    return dtype == AVEC_F32 ? reinterpret_cast<const float*>(p)[idx]
                             : __bfloat162float(reinterpret_cast<const bf16*>(p)[idx]);
}
__device__ __forceinline__ void st_any(void* p, int dtype, size_t idx, float v) {
    if (dtype == AVEC_F32) reinterpret_cast<float*>(p)[idx] = v;
    else reinterpret_cast<bf16*>(p)[idx] = __float2bfloat16_rn(v);
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }
__device__ __forceinline__ float swishf_(float x) { return x * sigmoidf_(x); }
// d/dx [x*sigmoid(x)] = s * (1 + x*(1-s))
__device__ __forceinline__ float dswishf_(float x) { float s = sigmoidf_(x); return s * (1.0f + x * (1.0f - s)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---- V consecutive elements <-> fp32 registers (V = 4 or 8; pointers must be V*sizeof(T)-aligned) -----------------
template <int V> __device__ __forceinline__ void load_vec(const float* p, float (&v)[V]) {
#pragma unroll
    for (int i = 0; i < V / 4; ++i) { float4 t = reinterpret_cast<const float4*>(p)[i]; v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w; }
}
template <int V> __device__ __forceinline__ void load_vec(const bf16* p, float (&v)[V]) {
    uint32_t w[V / 2];
    if (V == 8) { uint4 t = *reinterpret_cast<const uint4*>(p); w[0] = t.x; w[1] = t.y; w[V / 2 - 2] = t.z; w[V / 2 - 1] = t.w; }
    else { uint2 t = *reinterpret_cast<const uint2*>(p); w[0] = t.x; w[1] = t.y; }
#pragma unroll
    for (int i = 0; i < V / 2; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u); }
}
template <int V> __device__ __forceinline__ void store_vec(float* p, const float (&v)[V]) {
#pragma unroll
    for (int i = 0; i < V / 4; ++i) reinterpret_cast<float4*>(p)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
template <int V> __device__ __forceinline__ void store_vec(bf16* p, const float (&v)[V]) {
    uint32_t w[V / 2];
#pragma unroll
    for (int i = 0; i < V / 2; ++i) { __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]); w[i] = *reinterpret_cast<uint32_t*>(&t); }
    if (V == 8) *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[V / 2 - 2], w[V / 2 - 1]);
    else *reinterpret_cast<uint2*>(p) = make_uint2(w[0], w[1]);
}

// dispatch a runtime dtype code onto a template parameter
// dispatch dtype and vector width (V = 8 when C % 8 == 0, else 4; C % 4 == 0 is required by the vectorised kernels)
#define AVEC_DISPATCH_DTYPE_VEC(code, C_, T, V, ...)                                       \
    do {                                                                                    \
        if ((C_) % 4 != 0) return AVEC_ERR_INVALID;                                         \
        if ((code) == AVEC_F32) { using T = float;                                          \
            if ((C_) % 8 == 0) { constexpr int V = 8; __VA_ARGS__; } else { constexpr int V = 4; __VA_ARGS__; } } \
        else if ((code) == AVEC_BF16) { using T = bf16;                                     \
            if ((C_) % 8 == 0) { constexpr int V = 8; __VA_ARGS__; } else { constexpr int V = 4; __VA_ARGS__; } } \
        else return AVEC_ERR_INVALID;                                                       \
    } while (0)

#define AVEC_DISPATCH_DTYPE(code, T, ...)                                   \
    do {                                                                    \
        if ((code) == AVEC_F32) { using T = float; __VA_ARGS__; }           \
        else if ((code) == AVEC_BF16) { using T = bf16; __VA_ARGS__; }      \
        else return AVEC_ERR_INVALID;                                       \
    } while (0)

// ---- counter-based random bits (Philox4x32-10, Salmon et al. SC'11): dropout masks, SpecAugment / video-augmentation draws ----
// ---- programmatic dependent launch (sm_90+) ------------------------------------------------------------------------------
// pdl_trigger(): the next kernel of the stream, IF it was launched with cudaLaunchAttributeProgrammaticStreamSerialization, may be
// scheduled once every CTA of this grid has got here (its CTAs then run their prologue and park in pdl_wait()).  Without that
// attribute on the next launch the instruction does nothing.  pdl_wait(): returns once the preceding grid has completed and its
// memory operations are visible (immediately, when the kernel was launched the ordinary way).
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
    }
    return c;
}
// 16 random bits of element (row, col): 16-bit lane (col & 7) of Philox(ctr = (row, col >> 3, site, step), key = seed)
__device__ __forceinline__ uint4 dropout_bits(const unsigned long long* rng, uint32_t site, uint32_t row, uint32_t grp) {
    const unsigned long long seed = rng[0], step = rng[1];
    return philox4x32_10(make_uint4(row, grp, site, (uint32_t)step), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}
__device__ __forceinline__ uint32_t bits16(const uint4& r, int j) {   // j = 0..7
    const uint32_t w = (j >> 1) == 0 ? r.x : (j >> 1) == 1 ? r.y : (j >> 1) == 2 ? r.z : r.w;
    return (j & 1) ? (w >> 16) : (w & 0xFFFFu);
}

// ---- GEMM epilogue shared by the SIMT kernel and the tcgen05 kernel ------------------------------
struct EpiParams {
    int M, N;
    int kind;
    float alpha;
    const float* bias;
    void* out; int out_dtype; long long ldo;
    void* out2; int out2_dtype; long long ldo2;
    const void* aux; int aux_dtype; long long ldaux;
    float* colstats;
    // nn.Dropout fused behind the GEMM (tcgen05 fast epilogue only): the value the epilogue kind produces BEFORE the residual
    // add is multiplied by keep(row, col) * drop_scale, with exactly the mask avec_dropout draws for (rng, site)
    const unsigned long long* drop_rng; uint32_t drop_site, drop_thresh; float drop_scale;
};

static inline EpiParams make_epi(const avec_gemm_args* a) {
    EpiParams p;
    p.M = a->M; p.N = a->N; p.kind = a->epi; p.alpha = a->alpha; p.bias = a->bias;
    p.out = a->out; p.out_dtype = a->out_dtype; p.ldo = a->ldo;
    p.out2 = a->out2; p.out2_dtype = a->out2_dtype; p.ldo2 = a->ldo2;
    p.aux = a->aux; p.aux_dtype = a->aux_dtype; p.ldaux = a->ldaux;
    p.colstats = a->colstats;
    p.drop_rng = (a->drop_p > 0.0f) ? a->drop_rng : nullptr;
    p.drop_site = (uint32_t)a->drop_site;
    p.drop_thresh = (uint32_t)(a->drop_p * 65536.0f + 0.5f);
    p.drop_scale = a->drop_p > 0.0f ? 1.0f / (1.0f - a->drop_p) : 1.0f;
    return p;
}

// Apply the epilogue to one accumulator element (acc already holds the K-sum).  Returns v = acc + bias so that the
// caller can accumulate BatchNorm column statistics from exactly what was normalised.
__device__ __forceinline__ float epilogue_elem(const EpiParams& p, int row, int c, float acc) {
    float v = acc + (p.bias ? p.bias[c] : 0.0f);
    const size_t o = (size_t)row * p.ldo + c;
    switch (p.kind) {
    case AVEC_EPI_LINEAR:
        st_any(p.out, p.out_dtype, o, p.alpha * v);
        break;
    case AVEC_EPI_SWISH:
        if (p.out2) st_any(p.out2, p.out2_dtype, (size_t)row * p.ldo2 + c, v);
        st_any(p.out, p.out_dtype, o, swishf_(v));
        break;
    case AVEC_EPI_RESIDUAL:
        st_any(p.out, p.out_dtype, o, ld_any(p.aux, p.aux_dtype, (size_t)row * p.ldaux + c) + p.alpha * v);
        break;
    case AVEC_EPI_DSWISH:
        st_any(p.out, p.out_dtype, o, p.alpha * v * dswishf_(ld_any(p.aux, p.aux_dtype, (size_t)row * p.ldaux + c)));
        break;
    case AVEC_EPI_ACCUM:
        atomicAdd(reinterpret_cast<float*>(p.out) + o, p.alpha * v);
        break;
    case AVEC_EPI_RELU: {
        float r = p.alpha * v;
        if (p.aux) r += ld_any(p.aux, p.aux_dtype, (size_t)row * p.ldaux + c);
        st_any(p.out, p.out_dtype, o, fmaxf(r, 0.0f));
        break;
    }
    default: break;
    }
    return v;
}

// ---- convolution geometry helpers ------------------------------------------------------------------
struct ConvGeom {
    int N, Ti, Hi, Wi, C;
    int To, Ho, Wo, Co;
    int KT, KH, KW;
    int st, sh, sw;
    int pt, ph, pw;
};
static inline ConvGeom make_geom(const avec_conv_geom& g) {
    ConvGeom c;
    c.N = g.N; c.Ti = g.Ti; c.Hi = g.Hi; c.Wi = g.Wi; c.C = g.C;
    c.To = g.To; c.Ho = g.Ho; c.Wo = g.Wo; c.Co = g.Co;
    c.KT = g.KT; c.KH = g.KH; c.KW = g.KW;
    c.st = g.st; c.sh = g.sh; c.sw = g.sw;
    c.pt = g.pt; c.ph = g.ph; c.pw = g.pw;
    return c;
}
