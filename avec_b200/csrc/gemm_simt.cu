// SIMT (CUDA-core FMA, fp32 accumulate) GEMM / implicit-GEMM convolution.
// This is the *parity-mode* engine: fp32 or bf16 operands, exact fp32 accumulation, every addressing mode of
// avec_gemm (plain strided, conv fwd / dgrad / wgrad gathers) and every epilogue.  The production bf16 path for the
// big contractions is the tcgen05 kernel in gemm_tc.cu; avec_gemm() (api.cu) dispatches between the two.
#include "common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

template <typename T>
struct PlainLoader {
    const T* p; long long sr, sk; int R, K;
    __device__ __forceinline__ float operator()(int r, int k) const {
        return (r < R && k < K) ? ldf(p + (long long)r * sr + (long long)k * sk) : 0.0f;
    }
};

// A(m = output site, k = (tap, ci)) = X[n, to*st+kt-pt, ho*sh+kh-ph, wo*sw+kw-pw, ci]
template <typename T>
struct ConvFwdLoader {
    const T* x; ConvGeom g; int R, K;
    __device__ __forceinline__ float operator()(int m, int k) const {
        if (m >= R || k >= K) return 0.0f;
        int ci = k % g.C, tap = k / g.C;
        int kw = tap % g.KW; tap /= g.KW; int kh = tap % g.KH; int kt = tap / g.KH;
        int wo = m % g.Wo; int t = m / g.Wo; int ho = t % g.Ho; t /= g.Ho; int to = t % g.To; int n = t / g.To;
        int ti = to * g.st + kt - g.pt, hi = ho * g.sh + kh - g.ph, wi = wo * g.sw + kw - g.pw;
        if ((unsigned)ti >= (unsigned)g.Ti || (unsigned)hi >= (unsigned)g.Hi || (unsigned)wi >= (unsigned)g.Wi) return 0.0f;
        return ldf(x + ((((long long)n * g.Ti + ti) * g.Hi + hi) * g.Wi + wi) * g.C + ci);
    }
};

// A(m = input site, k = (tap, co)) = dY[n, (ti+pt-kt)/st, (hi+ph-kh)/sh, (wi+pw-kw)/sw, co]   (when divisible, in range)
template <typename T>
struct ConvDgradLoader {
    const T* dy; ConvGeom g; int R, K;
    __device__ __forceinline__ float operator()(int m, int k) const {
        if (m >= R || k >= K) return 0.0f;
        int co = k % g.Co, tap = k / g.Co;
        int kw = tap % g.KW; tap /= g.KW; int kh = tap % g.KH; int kt = tap / g.KH;
        int wi = m % g.Wi; int t = m / g.Wi; int hi = t % g.Hi; t /= g.Hi; int ti = t % g.Ti; int n = t / g.Ti;
        int a = ti + g.pt - kt, b = hi + g.ph - kh, c = wi + g.pw - kw;
        if (a < 0 || b < 0 || c < 0 || a % g.st || b % g.sh || c % g.sw) return 0.0f;
        int to = a / g.st, ho = b / g.sh, wo = c / g.sw;
        if (to >= g.To || ho >= g.Ho || wo >= g.Wo) return 0.0f;
        return ldf(dy + ((((long long)n * g.To + to) * g.Ho + ho) * g.Wo + wo) * g.Co + co);
    }
};
// B(n = (tap, ci), k = output site) = X[shift(site, tap)][ci]
template <typename T>
struct ConvWgradXLoader {
    const T* x; ConvGeom g; int R, K;
    __device__ __forceinline__ float operator()(int nn, int k) const {
        if (nn >= R || k >= K) return 0.0f;
        int ci = nn % g.C, tap = nn / g.C;
        int kw = tap % g.KW; tap /= g.KW; int kh = tap % g.KH; int kt = tap / g.KH;
        int wo = k % g.Wo; int t = k / g.Wo; int ho = t % g.Ho; t /= g.Ho; int to = t % g.To; int n = t / g.To;
        int ti = to * g.st + kt - g.pt, hi = ho * g.sh + kh - g.ph, wi = wo * g.sw + kw - g.pw;
        if ((unsigned)ti >= (unsigned)g.Ti || (unsigned)hi >= (unsigned)g.Hi || (unsigned)wi >= (unsigned)g.Wi) return 0.0f;
        return ldf(x + ((((long long)n * g.Ti + ti) * g.Hi + hi) * g.Wi + wi) * g.C + ci);
    }
};

template <typename LA, typename LB, bool A_KFAST, bool B_KFAST>
__global__ void __launch_bounds__(NT) gemm_simt_kernel(LA la, LB lb, int K, int k_per_split, EpiParams ep) {
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    __shared__ float cs[2][BN];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int kbeg = blockIdx.z * k_per_split;
    const int kend = min(K, kbeg + k_per_split);
    const int tx = tid % 16, ty = tid / 16;  // thread computes rows ty*4..+3, cols tx*4..+3
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

    for (int k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll
        for (int i = 0; i < (BM * BK) / NT; ++i) {
            int idx = tid + i * NT;
            int r, k;
            if (A_KFAST) { k = idx % BK; r = idx / BK; } else { r = idx % BM; k = idx / BM; }
            int kk = k0 + k;
            As[k][r] = kk < kend ? la(m0 + r, kk) : 0.0f;
        }
#pragma unroll
        for (int i = 0; i < (BN * BK) / NT; ++i) {
            int idx = tid + i * NT;
            int r, k;
            if (B_KFAST) { k = idx % BK; r = idx / BK; } else { r = idx % BN; k = idx / BN; }
            int kk = k0 + k;
            Bs[k][r] = kk < kend ? lb(n0 + r, kk) : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

    float s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int row = m0 + ty * 4 + i;
        if (row >= ep.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int c = n0 + tx * 4 + j;
            if (c >= ep.N) continue;
            // split-K slices other than the first must not add the bias again
            float v;
            if (blockIdx.z > 0) { EpiParams e2 = ep; e2.bias = nullptr; v = epilogue_elem(e2, row, c, acc[i][j]); }
            else v = epilogue_elem(ep, row, c, acc[i][j]);
            s1[j] += v; s2[j] += v * v;
        }
    }
    if (ep.colstats) {
        if (tid < BN) { cs[0][tid] = 0.0f; cs[1][tid] = 0.0f; }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 4; ++j) { atomicAdd(&cs[0][tx * 4 + j], s1[j]); atomicAdd(&cs[1][tx * 4 + j], s2[j]); }
        __syncthreads();
        if (tid < BN && n0 + tid < ep.N) {
            float* dst = ep.colstats + (size_t)(blockIdx.x % AVEC_STATS_REPLICAS) * 2 * ep.N;
            atomicAdd(dst + n0 + tid, cs[0][tid]);
            atomicAdd(dst + ep.N + n0 + tid, cs[1][tid]);
        }
    }
}

template <typename LA, typename LB, bool AK, bool BK_>
int launch(const LA& la, const LB& lb, const avec_gemm_args* a, cudaStream_t st) {
    int split = (a->epi == AVEC_EPI_ACCUM && a->split_k > 1) ? a->split_k : 1;
    int kps = cdiv(cdiv(a->K, split), BK) * BK;
    split = cdiv(a->K, kps);
    dim3 grid(cdiv(a->M, BM), cdiv(a->N, BN), split);
    if (grid.y > 65535u || grid.z > 65535u) return AVEC_ERR_INVALID;
    gemm_simt_kernel<LA, LB, AK, BK_><<<grid, NT, 0, st>>>(la, lb, a->K, kps, make_epi(a));
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}

template <typename T>
int dispatch(const avec_gemm_args* a, cudaStream_t st) {
    const T* A = reinterpret_cast<const T*>(a->A);
    const T* B = reinterpret_cast<const T*>(a->B);
    ConvGeom g = make_geom(a->g);
    const int taps = g.KT * g.KH * g.KW;
    switch (a->mode) {
    case AVEC_GEMM_PLAIN: {
        PlainLoader<T> la{A, a->sam, a->sak, a->M, a->K};
        PlainLoader<T> lb{B, a->sbn, a->sbk, a->N, a->K};
        bool ak = a->sak == 1 || a->sam != 1, bk = a->sbk == 1 || a->sbn != 1;
        if (ak && bk) return launch<PlainLoader<T>, PlainLoader<T>, true, true>(la, lb, a, st);
        if (ak && !bk) return launch<PlainLoader<T>, PlainLoader<T>, true, false>(la, lb, a, st);
        if (!ak && bk) return launch<PlainLoader<T>, PlainLoader<T>, false, true>(la, lb, a, st);
        return launch<PlainLoader<T>, PlainLoader<T>, false, false>(la, lb, a, st);
    }
    case AVEC_GEMM_CONV_FWD: {
        AVEC_CHECK_ARG(a->K == taps * g.C && a->N == g.Co && (long long)a->M == (long long)g.N * g.To * g.Ho * g.Wo);
        ConvFwdLoader<T> la{A, g, a->M, a->K};
        PlainLoader<T> lb{B, (long long)a->K, 1, a->N, a->K};
        return launch<ConvFwdLoader<T>, PlainLoader<T>, true, true>(la, lb, a, st);
    }
    case AVEC_GEMM_CONV_DGRAD: {
        AVEC_CHECK_ARG(a->K == taps * g.Co && a->N == g.C && (long long)a->M == (long long)g.N * g.Ti * g.Hi * g.Wi);
        ConvDgradLoader<T> la{A, g, a->M, a->K};
        PlainLoader<T> lb{B, (long long)a->K, 1, a->N, a->K};
        return launch<ConvDgradLoader<T>, PlainLoader<T>, true, true>(la, lb, a, st);
    }
    case AVEC_GEMM_CONV_WGRAD: {
        AVEC_CHECK_ARG(a->M == g.Co && a->N == taps * g.C && (long long)a->K == (long long)g.N * g.To * g.Ho * g.Wo);
        PlainLoader<T> la{A, 1, (long long)g.Co, a->M, a->K};
        ConvWgradXLoader<T> lb{B, g, a->N, a->K};
        return launch<PlainLoader<T>, ConvWgradXLoader<T>, false, false>(la, lb, a, st);
    }
    default: return AVEC_ERR_INVALID;
    }
}

}  // namespace

int avec_gemm_simt(const avec_gemm_args* a, cudaStream_t st) {
    AVEC_DISPATCH_DTYPE(a->ab_dtype, T, return dispatch<T>(a, st));
}
