// Relative-position multi-head self-attention core, forward and backward (one CTA per (batch item, head)).
// Semantics follow RelPos1dMultiHeadAttention.forwardQKV + rel_to_abs (reference nnet/attentions.py:258-323):
//   S[i,j] = (q_i.k_j + q_i.e_{T-1+j-i}) / sqrt(d) + (masked ? -1e9 : 0),  P = softmax_j S,  o_i = sum_j P_ij v_j
// K, V, (Q, dO) and E tiles of one head are staged in shared memory as fp32; scores never leave the SM in the forward
// except for the saved probabilities the backward consumes.
#include "common.cuh"

namespace {

constexpr int ATT_THREADS = 256;
constexpr int ATT_WARPS = ATT_THREADS / 32;
constexpr int MAX_KPL = 10;  // keys per lane  -> T <= 320
constexpr int MAX_CPL = 5;   // head channels per lane -> d <= 160

template <typename T>
__global__ void __launch_bounds__(ATT_THREADS) relpos_attn_fwd_kernel(
    const T* __restrict__ qkv, const T* __restrict__ e, const int* __restrict__ klen, int qlen, T* __restrict__ o,
    float* __restrict__ probs, int Tn, int H, int d) {
    extern __shared__ float sm[];
    const int ds = d + 1;
    float* Ks = sm;                       // [Tn][ds]
    float* Vs = Ks + (size_t)Tn * ds;     // [Tn][ds]
    float* Es = Vs + (size_t)Tn * ds;     // [2Tn-1][ds]
    float* qs = Es + (size_t)(2 * Tn - 1) * ds;  // [ATT_WARPS][ds]
    float* ps = qs + ATT_WARPS * ds;      // [ATT_WARPS][Tn]
    const int b = blockIdx.x / H, h = blockIdx.x % H;
    const int D = H * d;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const T* base = qkv + (size_t)b * Tn * 3 * D + h * d;
    for (int idx = tid; idx < Tn * d; idx += ATT_THREADS) {
        int j = idx / d, c = idx % d;
        Ks[j * ds + c] = ldf(base + (size_t)j * 3 * D + D + c);
        Vs[j * ds + c] = ldf(base + (size_t)j * 3 * D + 2 * D + c);
    }
    for (int idx = tid; idx < (2 * Tn - 1) * d; idx += ATT_THREADS) {
        int r = idx / d, c = idx % d;
        Es[r * ds + c] = ldf(e + (size_t)r * D + h * d + c);
    }
    __syncthreads();
    const int kl = klen ? klen[b] : Tn;
    const float scale = rsqrtf((float)d);
    float* q = qs + warp * ds;
    float* p = ps + warp * Tn;
    for (int i = warp; i < Tn; i += ATT_WARPS) {
        for (int c = lane; c < d; c += 32) q[c] = ldf(base + (size_t)i * 3 * D + c);
        __syncwarp();
        float s[MAX_KPL];
        float mx = -INFINITY;
#pragma unroll
        for (int u = 0; u < MAX_KPL; ++u) {
            int j = lane + u * 32;
            s[u] = -INFINITY;
            if (j < Tn) {
                const float* kr = Ks + j * ds;
                const float* er = Es + (Tn - 1 + j - i) * ds;
                float acc = 0.0f;
                for (int c = 0; c < d; ++c) acc = fmaf(q[c], kr[c] + er[c], acc);
                acc *= scale;
                if (j >= kl || i >= qlen) acc += -1e9f;
                s[u] = acc;
                mx = fmaxf(mx, acc);
            }
        }
        mx = warp_max(mx);
        float sum = 0.0f;
#pragma unroll
        for (int u = 0; u < MAX_KPL; ++u) {
            int j = lane + u * 32;
            if (j < Tn) { s[u] = __expf(s[u] - mx); sum += s[u]; }
        }
        sum = warp_sum(sum);
        const float inv = 1.0f / sum;
        float* prow = probs + (((size_t)b * H + h) * Tn + i) * Tn;
#pragma unroll
        for (int u = 0; u < MAX_KPL; ++u) {
            int j = lane + u * 32;
            if (j < Tn) { float pv = s[u] * inv; p[j] = pv; prow[j] = pv; }
        }
        __syncwarp();
        float acc[MAX_CPL];
#pragma unroll
        for (int u = 0; u < MAX_CPL; ++u) acc[u] = 0.0f;
        for (int j = 0; j < Tn; ++j) {
            float pj = p[j];
            const float* vr = Vs + j * ds;
#pragma unroll
            for (int u = 0; u < MAX_CPL; ++u) { int c = lane + u * 32; if (c < d) acc[u] = fmaf(pj, vr[c], acc[u]); }
        }
#pragma unroll
        for (int u = 0; u < MAX_CPL; ++u) { int c = lane + u * 32; if (c < d) stf(o + ((size_t)b * Tn + i) * D + h * d + c, acc[u]); }
        __syncwarp();
    }
}

// Backward.  Phase B: dV_j = sum_i P_ij dO_i.  Phase A: dP, delta, dS (written to ds_ws), dQ.  Phase C: dK_j, dE_r.
template <typename T>
__global__ void __launch_bounds__(ATT_THREADS) relpos_attn_bwd_kernel(
    const T* __restrict__ d_o, const T* __restrict__ qkv, const T* __restrict__ e, const float* __restrict__ probs,
    float* __restrict__ ds_ws, T* __restrict__ dqkv, float* __restrict__ de, int Tn, int H, int d) {
    extern __shared__ float sm[];
    const int ds = d + 1;
    float* Qs = sm;
    float* Ks = Qs + (size_t)Tn * ds;
    float* Vs = Ks + (size_t)Tn * ds;
    float* Os = Vs + (size_t)Tn * ds;  // dO
    float* ps = Os + (size_t)Tn * ds;  // [ATT_WARPS][Tn] scratch row
    const int b = blockIdx.x / H, h = blockIdx.x % H;
    const int D = H * d;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const T* base = qkv + (size_t)b * Tn * 3 * D + h * d;
    T* dbase = dqkv + (size_t)b * Tn * 3 * D + h * d;
    for (int idx = tid; idx < Tn * d; idx += ATT_THREADS) {
        int j = idx / d, c = idx % d;
        Qs[j * ds + c] = ldf(base + (size_t)j * 3 * D + c);
        Ks[j * ds + c] = ldf(base + (size_t)j * 3 * D + D + c);
        Vs[j * ds + c] = ldf(base + (size_t)j * 3 * D + 2 * D + c);
        Os[j * ds + c] = ldf(d_o + ((size_t)b * Tn + j) * D + h * d + c);
    }
    __syncthreads();
    const float scale = rsqrtf((float)d);
    const float* P = probs + ((size_t)b * H + h) * Tn * Tn;
    float* dS = ds_ws + ((size_t)b * H + h) * Tn * Tn;
    float* p = ps + warp * Tn;

    // ---- phase B: dV_j = sum_i P[i][j] * dO_i  (warp per key j, lane owns channels)
    for (int j = warp; j < Tn; j += ATT_WARPS) {
        float acc[MAX_CPL];
#pragma unroll
        for (int u = 0; u < MAX_CPL; ++u) acc[u] = 0.0f;
        for (int i0 = 0; i0 < Tn; i0 += 32) {
            int i = i0 + lane;
            float pij = i < Tn ? P[(size_t)i * Tn + j] : 0.0f;
            int cnt = min(32, Tn - i0);
            for (int ii = 0; ii < cnt; ++ii) {
                float pv = __shfl_sync(0xffffffffu, pij, ii);
                const float* orow = Os + (i0 + ii) * ds;
#pragma unroll
                for (int u = 0; u < MAX_CPL; ++u) { int c = lane + u * 32; if (c < d) acc[u] = fmaf(pv, orow[c], acc[u]); }
            }
        }
#pragma unroll
        for (int u = 0; u < MAX_CPL; ++u) { int c = lane + u * 32; if (c < d) stf(dbase + (size_t)j * 3 * D + 2 * D + c, acc[u]); }
    }

    // ---- phase A: per query row i: dP_ij = dO_i.V_j, delta, dS_ij, dQ_i
    for (int i = warp; i < Tn; i += ATT_WARPS) {
        const float* orow = Os + i * ds;
        float dp[MAX_KPL], pr[MAX_KPL];
        float delta = 0.0f;
#pragma unroll
        for (int u = 0; u < MAX_KPL; ++u) {
            int j = lane + u * 32;
            dp[u] = 0.0f; pr[u] = 0.0f;
            if (j < Tn) {
                const float* vr = Vs + j * ds;
                float acc = 0.0f;
                for (int c = 0; c < d; ++c) acc = fmaf(orow[c], vr[c], acc);
                dp[u] = acc;
                pr[u] = P[(size_t)i * Tn + j];
                delta = fmaf(pr[u], acc, delta);
            }
        }
        delta = warp_sum(delta);
#pragma unroll
        for (int u = 0; u < MAX_KPL; ++u) {
            int j = lane + u * 32;
            if (j < Tn) { float v = pr[u] * (dp[u] - delta) * scale; p[j] = v; dS[(size_t)i * Tn + j] = v; }
        }
        __syncwarp();
        float acc[MAX_CPL];
#pragma unroll
        for (int u = 0; u < MAX_CPL; ++u) acc[u] = 0.0f;
        for (int j = 0; j < Tn; ++j) {
            float sv = p[j];
            const float* kr = Ks + j * ds;
            const T* er = e + (size_t)(Tn - 1 + j - i) * D + h * d;
#pragma unroll
            for (int u = 0; u < MAX_CPL; ++u) { int c = lane + u * 32; if (c < d) acc[u] = fmaf(sv, kr[c] + ldf(er + c), acc[u]); }
        }
#pragma unroll
        for (int u = 0; u < MAX_CPL; ++u) { int c = lane + u * 32; if (c < d) stf(dbase + (size_t)i * 3 * D + c, acc[u]); }
        __syncwarp();
    }
    __syncthreads();

    // ---- phase C1: dK_j = sum_i dS[i][j] * Q_i
    for (int j = warp; j < Tn; j += ATT_WARPS) {
        float acc[MAX_CPL];
#pragma unroll
        for (int u = 0; u < MAX_CPL; ++u) acc[u] = 0.0f;
        for (int i0 = 0; i0 < Tn; i0 += 32) {
            int i = i0 + lane;
            float sij = i < Tn ? dS[(size_t)i * Tn + j] : 0.0f;
            int cnt = min(32, Tn - i0);
            for (int ii = 0; ii < cnt; ++ii) {
                float sv = __shfl_sync(0xffffffffu, sij, ii);
                const float* qrow = Qs + (i0 + ii) * ds;
#pragma unroll
                for (int u = 0; u < MAX_CPL; ++u) { int c = lane + u * 32; if (c < d) acc[u] = fmaf(sv, qrow[c], acc[u]); }
            }
        }
#pragma unroll
        for (int u = 0; u < MAX_CPL; ++u) { int c = lane + u * 32; if (c < d) stf(dbase + (size_t)j * 3 * D + D + c, acc[u]); }
    }
    // ---- phase C2: dE_r = sum_{i, j = r-(Tn-1)+i in [0,Tn)} dS[i][j] * Q_i   (summed over the batch: atomics)
    for (int r = warp; r < 2 * Tn - 1; r += ATT_WARPS) {
        int ilo = max(0, Tn - 1 - r), ihi = min(Tn - 1, 2 * Tn - 2 - r);  // j = r-(Tn-1)+i
        float acc[MAX_CPL];
#pragma unroll
        for (int u = 0; u < MAX_CPL; ++u) acc[u] = 0.0f;
        for (int i0 = ilo; i0 <= ihi; i0 += 32) {
            int i = i0 + lane;
            float sij = i <= ihi ? dS[(size_t)i * Tn + (r - (Tn - 1) + i)] : 0.0f;
            int cnt = min(32, ihi - i0 + 1);
            for (int ii = 0; ii < cnt; ++ii) {
                float sv = __shfl_sync(0xffffffffu, sij, ii);
                const float* qrow = Qs + (i0 + ii) * ds;
#pragma unroll
                for (int u = 0; u < MAX_CPL; ++u) { int c = lane + u * 32; if (c < d) acc[u] = fmaf(sv, qrow[c], acc[u]); }
            }
        }
#pragma unroll
        for (int u = 0; u < MAX_CPL; ++u) { int c = lane + u * 32; if (c < d) atomicAdd(de + (size_t)r * D + h * d + c, acc[u]); }
    }
}

size_t fwd_smem(int T, int d) { return ((size_t)(4 * T - 1) * (d + 1) + ATT_WARPS * (d + 1) + ATT_WARPS * T) * sizeof(float); }
size_t bwd_smem(int T, int d) { return ((size_t)4 * T * (d + 1) + ATT_WARPS * T) * sizeof(float); }

}  // namespace

extern "C" int avec_relpos_attn_fwd(const void* qkv, const void* e, const int* klen, int qlen, void* o, float* probs, int B,
                                    int T, int H, int d, int dtype, avec_stream_t stream) {
    AVEC_CHECK_ARG(qkv && e && o && probs && B > 0 && T > 0 && H > 0 && d > 0);
    AVEC_CHECK_ARG(T <= 32 * MAX_KPL && d <= 32 * MAX_CPL);
    size_t smem = fwd_smem(T, d);
    if (smem > 227 * 1024) return AVEC_ERR_UNSUPPORTED;
    AVEC_DISPATCH_DTYPE(dtype, Tt, {
        auto kfn = relpos_attn_fwd_kernel<Tt>;
        if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return AVEC_ERR_LAUNCH;
        kfn<<<B * H, ATT_THREADS, smem, as_stream(stream)>>>((const Tt*)qkv, (const Tt*)e, klen, qlen, (Tt*)o, probs, T, H, d);
    });
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}

extern "C" int avec_relpos_attn_bwd(const void* d_o, const void* qkv, const void* e, const float* probs, float* ds_ws,
                                    void* dqkv, float* de, int B, int T, int H, int d, int dtype, avec_stream_t stream) {
    AVEC_CHECK_ARG(d_o && qkv && e && probs && ds_ws && dqkv && de && B > 0 && T > 0);
    AVEC_CHECK_ARG(T <= 32 * MAX_KPL && d <= 32 * MAX_CPL);
    size_t smem = bwd_smem(T, d);
    if (smem > 227 * 1024) return AVEC_ERR_UNSUPPORTED;
    AVEC_DISPATCH_DTYPE(dtype, Tt, {
        auto kfn = relpos_attn_bwd_kernel<Tt>;
        if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return AVEC_ERR_LAUNCH;
        kfn<<<B * H, ATT_THREADS, smem, as_stream(stream)>>>((const Tt*)d_o, (const Tt*)qkv, (const Tt*)e, probs, ds_ws, (Tt*)dqkv, de, T, H, d);
    });
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}
