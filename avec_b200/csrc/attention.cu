// Relative-position multi-head self-attention core, forward and backward (one CTA per (batch item, head)).
// Semantics follow RelPos1dMultiHeadAttention.forwardQKV + rel_to_abs (reference nnet/attentions.py:258-323):
//   S[i,j] = (q_i.k_j + q_i.e_{T-1+j-i}) / sqrt(d) + (masked ? -1e9 : 0),  P = softmax_j S,  o_i = sum_j P_ij v_j
// K, V, (Q, dO) and E tiles of one head are staged in shared memory as fp32; scores never leave the SM in the forward
// except for the saved probabilities the backward consumes.
#include "attention_common.cuh"

namespace {

template <typename T>
__global__ void __launch_bounds__(ATT_THREADS) relpos_attn_fwd_kernel(
    const T* __restrict__ qkv, const T* __restrict__ e, const int* __restrict__ klen, int qlen, T* __restrict__ o,
    float* __restrict__ probs, int Tn, int H, int d, int G, int D1, int Tf, const float* __restrict__ ub, const float* __restrict__ vb) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    const int ds = att_ds<T>(d), qds = d + 1;
    T* Ks = reinterpret_cast<T*>(sm_raw);  // [Tn][ds]
    T* Vs = Ks + (size_t)Tn * ds;          // [Tn][ds]
    T* Es = Vs + (size_t)Tn * ds;          // [2Tn-1][ds]
    float* qs = reinterpret_cast<float*>(Es + (size_t)(2 * Tn - 1) * ds);  // [ATT_WARPS][2][qds]  (q + u | q + v), fp32
    float* ps = qs + ATT_WARPS * 2 * qds;  // [ATT_WARPS][Tn]
    const int b = blockIdx.x / H, h = blockIdx.x % H;
    const int D = H * d;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const T* qkv_b = qkv + (size_t)b * Tf * 3 * D1;
    for (int idx = tid; idx < Tn * d; idx += ATT_THREADS) {
        int j = idx / d, c = idx % d;
        stf(Ks + j * ds + c, fetch_qkv(qkv_b, j, h * d + c, 1, G, D1, Tf));
        stf(Vs + j * ds + c, fetch_qkv(qkv_b, j, h * d + c, 2, G, D1, Tf));
    }
    for (int idx = tid; idx < (2 * Tn - 1) * d; idx += ATT_THREADS) {
        int r = idx / d, c = idx % d;
        Es[r * ds + c] = e[(size_t)r * D + h * d + c];
    }
    __syncthreads();
    const int kl = klen ? klen[b] : Tn;
    const float scale = rsqrtf((float)d);
    float* q = qs + warp * 2 * qds;  // q + u (content term)
    float* qv = q + qds;             // q + v (position term)
    float* p = ps + warp * Tn;
    for (int i = warp; i < Tn; i += ATT_WARPS) {
        for (int c = lane; c < d; c += 32) {
            const float qq = fetch_qkv(qkv_b, i, h * d + c, 0, G, D1, Tf);
            q[c] = qq + (ub ? ub[(h * d + c) % D1] : 0.0f);
            qv[c] = qq + (vb ? vb[(h * d + c) % D1] : 0.0f);
        }
        __syncwarp();
        float s[MAX_KPL];
        float mx = -INFINITY;
#pragma unroll
        for (int u = 0; u < MAX_KPL; ++u) {
            int j = lane + u * 32;
            s[u] = -INFINITY;
            if (j < Tn) {
                const T* kr = Ks + j * ds;
                const T* er = Es + (Tn - 1 + j - i) * ds;
                float acc = 0.0f;
                for (int c = 0; c < d; ++c) acc = fmaf(q[c], ldf(kr + c), fmaf(qv[c], ldf(er + c), acc));
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
            const T* vr = Vs + j * ds;
#pragma unroll
            for (int u = 0; u < MAX_CPL; ++u) { int c = lane + u * 32; if (c < d) acc[u] = fmaf(pj, ldf(vr + c), acc[u]); }
        }
#pragma unroll
        for (int u = 0; u < MAX_CPL; ++u) {
            int c = lane + u * 32;
            if (c < d) {
                const int ee = h * d + c, fi = ee / D1, frame = i * G + fi;
                if (frame < Tf) stf(o + ((size_t)b * Tf + frame) * D1 + (ee - fi * D1), acc[u]);
            }
        }
        __syncwarp();
    }
}

// Backward.  Phase B: dV_j = sum_i P_ij dO_i.  Phase A: dP, delta, dS (written to ds_ws), dQ.  Phase C: dK_j, dE_r.
template <typename T>
__global__ void __launch_bounds__(ATT_THREADS) relpos_attn_bwd_kernel(
    const T* __restrict__ d_o, const T* __restrict__ qkv, const T* __restrict__ e, const float* __restrict__ probs,
    float* __restrict__ ds_ws, T* __restrict__ dqkv, float* __restrict__ de, int Tn, int H, int d, int G, int D1, int Tf,
    const float* __restrict__ ub, const float* __restrict__ vb, float* __restrict__ dub, float* __restrict__ dvb) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    const int ds = att_ds<T>(d);
    T* Qs = reinterpret_cast<T*>(sm_raw);
    T* Ks = Qs + (size_t)Tn * ds;
    T* Vs = Ks + (size_t)Tn * ds;
    T* Os = Vs + (size_t)Tn * ds;  // dO
    float* ps = reinterpret_cast<float*>(Os + (size_t)Tn * ds);  // [ATT_WARPS][Tn] scratch row
    const int b = blockIdx.x / H, h = blockIdx.x % H;
    const int D = H * d;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const T* qkv_b = qkv + (size_t)b * Tf * 3 * D1;
    T* dqkv_b = dqkv + (size_t)b * Tf * 3 * D1;
    for (int idx = tid; idx < Tn * d; idx += ATT_THREADS) {
        int j = idx / d, c = idx % d;
        const int ee = h * d + c, fi = ee / D1, frame = j * G + fi;
        stf(Qs + j * ds + c, fetch_qkv(qkv_b, j, ee, 0, G, D1, Tf));
        stf(Ks + j * ds + c, fetch_qkv(qkv_b, j, ee, 1, G, D1, Tf));
        stf(Vs + j * ds + c, fetch_qkv(qkv_b, j, ee, 2, G, D1, Tf));
        stf(Os + j * ds + c, frame < Tf ? ldf(d_o + ((size_t)b * Tf + frame) * D1 + (ee - fi * D1)) : 0.0f);
    }
    // gradient element (token, channel c of this head) -> [frames, 3 * D1] matrix; padded frames are dropped
    auto store_grad = [&](int tok, int c, int which, float val) {
        const int ee = h * d + c, fi = ee / D1, frame = tok * G + fi;
        if (frame < Tf) stf(dqkv_b + (size_t)frame * 3 * D1 + which * D1 + (ee - fi * D1), val);
    };
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
                const T* orow = Os + (i0 + ii) * ds;
#pragma unroll
                for (int u = 0; u < MAX_CPL; ++u) { int c = lane + u * 32; if (c < d) acc[u] = fmaf(pv, ldf(orow + c), acc[u]); }
            }
        }
#pragma unroll
        for (int u = 0; u < MAX_CPL; ++u) { int c = lane + u * 32; if (c < d) store_grad(j, c, 2, acc[u]); }
    }

    // per-lane u / v bias values of this head's channels, and their gradient accumulators (summed over this warp's rows)
    float ubv[MAX_CPL], vbv[MAX_CPL], du_acc[MAX_CPL], dv_acc[MAX_CPL];
#pragma unroll
    for (int u = 0; u < MAX_CPL; ++u) {
        int c = lane + u * 32;
        ubv[u] = (ub && c < d) ? ub[(h * d + c) % D1] : 0.0f;
        vbv[u] = (vb && c < d) ? vb[(h * d + c) % D1] : 0.0f;
        du_acc[u] = 0.0f; dv_acc[u] = 0.0f;
    }

    // ---- phase A: per query row i: dP_ij = dO_i.V_j, delta, dS_ij, dQ_i
    for (int i = warp; i < Tn; i += ATT_WARPS) {
        const T* orow = Os + i * ds;
        float dp[MAX_KPL], pr[MAX_KPL];
        float delta = 0.0f;
#pragma unroll
        for (int u = 0; u < MAX_KPL; ++u) {
            int j = lane + u * 32;
            dp[u] = 0.0f; pr[u] = 0.0f;
            if (j < Tn) {
                const T* vr = Vs + j * ds;
                float acc = 0.0f;
                for (int c = 0; c < d; ++c) acc = fmaf(ldf(orow + c), ldf(vr + c), acc);
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
        float acc[MAX_CPL], acce[MAX_CPL];
#pragma unroll
        for (int u = 0; u < MAX_CPL; ++u) { acc[u] = 0.0f; acce[u] = 0.0f; }
        for (int j = 0; j < Tn; ++j) {
            float sv = p[j];
            const T* kr = Ks + j * ds;
            const T* er = e + (size_t)(Tn - 1 + j - i) * D + h * d;
#pragma unroll
            for (int u = 0; u < MAX_CPL; ++u) {
                int c = lane + u * 32;
                if (c < d) { acc[u] = fmaf(sv, ldf(kr + c), acc[u]); acce[u] = fmaf(sv, ldf(er + c), acce[u]); }
            }
        }
#pragma unroll
        for (int u = 0; u < MAX_CPL; ++u) {
            int c = lane + u * 32;
            if (c < d) {
                store_grad(i, c, 0, acc[u] + acce[u]);
                // u / v biases are added to every (also zero-padded) query row: their gradients take the two parts separately
                du_acc[u] += acc[u];
                dv_acc[u] += acce[u];
            }
        }
        __syncwarp();
    }
#pragma unroll
    for (int u = 0; u < MAX_CPL; ++u) {
        int c = lane + u * 32;
        if (c < d) {
            if (dub) atomicAdd(dub + (h * d + c) % D1, du_acc[u]);
            if (dvb) atomicAdd(dvb + (h * d + c) % D1, dv_acc[u]);
        }
    }
    __syncthreads();

    // ---- phase C1: dK_j = sum_i dS[i][j] * (Q_i + u)
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
                const T* qrow = Qs + (i0 + ii) * ds;
#pragma unroll
                for (int u = 0; u < MAX_CPL; ++u) { int c = lane + u * 32; if (c < d) acc[u] = fmaf(sv, ldf(qrow + c) + ubv[u], acc[u]); }
            }
        }
#pragma unroll
        for (int u = 0; u < MAX_CPL; ++u) { int c = lane + u * 32; if (c < d) store_grad(j, c, 1, acc[u]); }
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
                const T* qrow = Qs + (i0 + ii) * ds;
#pragma unroll
                for (int u = 0; u < MAX_CPL; ++u) { int c = lane + u * 32; if (c < d) acc[u] = fmaf(sv, ldf(qrow + c) + vbv[u], acc[u]); }
            }
        }
#pragma unroll
        for (int u = 0; u < MAX_CPL; ++u) { int c = lane + u * 32; if (c < d) atomicAdd(de + (size_t)r * D + h * d + c, acc[u]); }
    }
}

template <typename T> size_t fwd_smem(int Tn, int d) {
    return (size_t)(4 * Tn - 1) * att_ds<T>(d) * sizeof(T) + ((size_t)ATT_WARPS * 2 * (d + 1) + (size_t)ATT_WARPS * Tn) * sizeof(float);
}
template <typename T> size_t bwd_smem(int Tn, int d) {
    return (size_t)4 * Tn * att_ds<T>(d) * sizeof(T) + (size_t)ATT_WARPS * Tn * sizeof(float);
}

}  // namespace

// tensor-core kernels (attention_mma.cu); AVEC_ERR_UNSUPPORTED = shape outside their envelope
int avec_attn_mma_fwd(const void* qkv, const void* e, const int* klen, int qlen, void* o, float* probs, int B, int T, int H, int d, cudaStream_t st);
int avec_attn_mma_bwd(const void* d_o, const void* qkv, const void* e, const float* probs, void* dqkv, float* de, int B, int T, int H, int d,
                      cudaStream_t st);

// key-tiled kernels for sequences beyond the whole-head envelope (attention_long.cu)
int avec_attn_long_fwd(const void* qkv, const void* e, const int* klen, int qlen, void* o, float* probs, int B, int T, int H, int d, int G,
                       int Tf, const float* u, const float* v, int dtype, cudaStream_t st);
int avec_attn_long_bwd(const void* d_o, const void* qkv, const void* e, const float* probs, float* ds_ws, void* dqkv, float* de, int B, int T,
                       int H, int d, int G, int Tf, const float* u, const float* v, float* du, float* dv, int dtype, cudaStream_t st);

static int g_force_long = 0;
// whole-head kernels run ONE CTA per (item, head) (8 warps on an SM); from ~256 keys on the key-tiled kernels' T/32 x B x H grid wins
// (AO regular attention, 800 mel frames, B = 32: 58.8 vs 70.6 ms per training step, profiles/r01_ablation_sweep.md)
constexpr int LONG_FROM_T = 256;
extern "C" void avec_set_attention_long(int force) { g_force_long = force; }

extern "C" int avec_relpos_attn_fwd(const void* qkv, const void* e, const int* klen, int qlen, void* o, float* probs, int B,
                                    int T, int H, int d, int G, int Tf, const float* u, const float* v, int dtype, avec_stream_t stream) {
    AVEC_CHECK_ARG(qkv && e && o && probs && B > 0 && T > 0 && H > 0 && d > 0 && G >= 1 && (H * d) % G == 0);
    AVEC_CHECK_ARG(d <= 32 * MAX_CPL && Tf > (T - 1) * G && Tf <= T * G);
    if (dtype == AVEC_BF16 && G == 1 && !u && !v && Tf == T && !g_force_long) {
        const int rc = avec_attn_mma_fwd(qkv, e, klen, qlen, o, probs, B, T, H, d, as_stream(stream));
        if (rc != AVEC_ERR_UNSUPPORTED) return rc;
    }
    const int D1 = H * d / G;
    AVEC_DISPATCH_DTYPE(dtype, Tt, {
        const size_t smem = fwd_smem<Tt>(T, d);
        if (g_force_long || T > LONG_FROM_T || smem > 227 * 1024)
            return avec_attn_long_fwd(qkv, e, klen, qlen, o, probs, B, T, H, d, G, Tf, u, v, dtype, as_stream(stream));
        auto kfn = relpos_attn_fwd_kernel<Tt>;
        if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return AVEC_ERR_LAUNCH;
        kfn<<<B * H, ATT_THREADS, smem, as_stream(stream)>>>((const Tt*)qkv, (const Tt*)e, klen, qlen, (Tt*)o, probs, T, H, d, G, D1, Tf, u, v);
    });
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}

extern "C" int avec_relpos_attn_bwd(const void* d_o, const void* qkv, const void* e, const float* probs, float* ds_ws,
                                    void* dqkv, float* de, int B, int T, int H, int d, int G, int Tf, const float* u, const float* v,
                                    float* du, float* dv, int dtype, avec_stream_t stream) {
    AVEC_CHECK_ARG(d_o && qkv && e && probs && ds_ws && dqkv && de && B > 0 && T > 0 && G >= 1 && (H * d) % G == 0);
    AVEC_CHECK_ARG(d <= 32 * MAX_CPL && Tf > (T - 1) * G && Tf <= T * G);
    if (dtype == AVEC_BF16 && G == 1 && !u && !v && !du && !dv && Tf == T && !g_force_long) {
        const int rc = avec_attn_mma_bwd(d_o, qkv, e, probs, dqkv, de, B, T, H, d, as_stream(stream));
        if (rc != AVEC_ERR_UNSUPPORTED) return rc;
    }
    const int D1 = H * d / G;
    AVEC_DISPATCH_DTYPE(dtype, Tt, {
        const size_t smem = bwd_smem<Tt>(T, d);
        if (g_force_long || T > LONG_FROM_T || smem > 227 * 1024)
            return avec_attn_long_bwd(d_o, qkv, e, probs, ds_ws, dqkv, de, B, T, H, d, G, Tf, u, v, du, dv, dtype, as_stream(stream));
        auto kfn = relpos_attn_bwd_kernel<Tt>;
        if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return AVEC_ERR_LAUNCH;
        kfn<<<B * H, ATT_THREADS, smem, as_stream(stream)>>>((const Tt*)d_o, (const Tt*)qkv, (const Tt*)e, probs, ds_ws, (Tt*)dqkv, de, T, H, d,
                                                             G, D1, Tf, u, v, du, dv);
    });
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}
