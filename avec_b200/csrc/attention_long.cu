// Key-tiled relative-position attention for LONG sequences (BASELINE.json configs[4]: 800 / 1600 mel frames), forward and
// backward.  Same semantics and the same saved tensors as attention.cu (probs [B,H,T,T], dS workspace), but nothing of size
// O(T * d) is resident: a CTA owns 32 query rows (or 32 keys / 32 relative offsets in the backward) of one (item, head) and
// streams K / V / E (or Q / dO) through shared memory in tiles of 64 rows; only the 32 x T score rows stay in shared memory.
// Envelope: T <= ~1500 tokens (32 x T fp32 rows + tiles <= 227 KB), d <= 160.  Used when the whole-head kernels of attention.cu
// do not fit (T > 416 keys or tiles > 227 KB); CUDA-core arithmetic (the tensor-core kernels cover T <= 128).
#include "attention_common.cuh"

namespace {

constexpr int LQB = 32;                 // query rows (keys, offsets) per CTA
constexpr int LRPW = LQB / ATT_WARPS;   // rows per warp
constexpr int LTK = 64;                 // rows of a streamed tile

// stage `rows` rows of q / k / v (which = 0 / 1 / 2) of tokens tok0.. into tile[rows][ds] (zero rows beyond Tn)
template <typename T>
__device__ __forceinline__ void stage_qkv(T* tile, const T* qkv_b, int tok0, int rows, int which, int Tn, int h, int d, int ds, int G,
                                          int D1, int Tf) {
    for (int idx = threadIdx.x; idx < rows * d; idx += ATT_THREADS) {
        const int jj = idx / d, c = idx - jj * d, j = tok0 + jj;
        stf(tile + jj * ds + c, j < Tn ? fetch_qkv(qkv_b, j, h * d + c, which, G, D1, Tf) : 0.0f);
    }
}
// stage rows of dO ([frames, D1] matrix, grouped layout as q / k / v)
template <typename T>
__device__ __forceinline__ void stage_do(T* tile, const T* do_b, int tok0, int rows, int Tn, int h, int d, int ds, int G, int D1, int Tf) {
    for (int idx = threadIdx.x; idx < rows * d; idx += ATT_THREADS) {
        const int jj = idx / d, c = idx - jj * d, j = tok0 + jj;
        const int ee = h * d + c, fi = ee / D1, frame = j * G + fi;
        stf(tile + jj * ds + c, (j < Tn && frame < Tf) ? ldf(do_b + (size_t)frame * D1 + (ee - fi * D1)) : 0.0f);
    }
}
// stage the window of E rows a (query block i0, key tile j0) pair touches: local row rl <-> r = Tn-1 + j0 - (i0+LQB-1) + rl
template <typename T>
__device__ __forceinline__ void stage_e(T* tile, const T* e, int i0, int j0, int Tn, int h, int d, int ds, int D) {
    const int r0 = Tn - 1 + j0 - (i0 + LQB - 1);
    for (int idx = threadIdx.x; idx < (LTK + LQB - 1) * d; idx += ATT_THREADS) {
        const int rl = idx / d, c = idx - rl * d, r = r0 + rl;
        stf(tile + rl * ds + c, (r >= 0 && r < 2 * Tn - 1) ? ldf(e + (size_t)r * D + h * d + c) : 0.0f);
    }
}

template <typename T>
__global__ void __launch_bounds__(ATT_THREADS) attn_long_fwd_kernel(
    const T* __restrict__ qkv, const T* __restrict__ e, const int* __restrict__ klen, int qlen, T* __restrict__ o,
    float* __restrict__ probs, int Tn, int H, int d, int G, int D1, int Tf, const float* __restrict__ ub, const float* __restrict__ vb) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    const int ds = att_ds<T>(d), qds = d + 1;
    T* Kt = reinterpret_cast<T*>(sm_raw);                 // [LTK][ds]   K tile, later V tile
    T* Et = Kt + (size_t)LTK * ds;                        // [LTK+LQB-1][ds]
    float* qs = reinterpret_cast<float*>(Et + (size_t)(LTK + LQB - 1) * ds);   // [LQB][2][qds]  q + u | q + v
    float* ps = qs + (size_t)LQB * 2 * qds;               // [LQB][Tn] scores -> probabilities
    const int b = blockIdx.y / H, h = blockIdx.y % H, i0 = blockIdx.x * LQB;
    const int D = H * d;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const T* qkv_b = qkv + (size_t)b * Tf * 3 * D1;
    for (int idx = tid; idx < LQB * d; idx += ATT_THREADS) {
        const int rr = idx / d, c = idx - rr * d, i = i0 + rr;
        const float qq = i < Tn ? fetch_qkv(qkv_b, i, h * d + c, 0, G, D1, Tf) : 0.0f;
        qs[(size_t)(rr * 2) * qds + c] = qq + (ub ? ub[(h * d + c) % D1] : 0.0f);
        qs[(size_t)(rr * 2 + 1) * qds + c] = qq + (vb ? vb[(h * d + c) % D1] : 0.0f);
    }
    const int kl = klen ? klen[b] : Tn;
    const float scale = rsqrtf((float)d);
    // ---- scores
    for (int j0 = 0; j0 < Tn; j0 += LTK) {
        __syncthreads();
        stage_qkv(Kt, qkv_b, j0, LTK, 1, Tn, h, d, ds, G, D1, Tf);
        stage_e(Et, e, i0, j0, Tn, h, d, ds, D);
        __syncthreads();
        for (int rr = 0; rr < LRPW; ++rr) {
            const int row = warp * LRPW + rr, i = i0 + row;
            if (i >= Tn) continue;
            const float* q = qs + (size_t)(row * 2) * qds;
            const float* qv = q + qds;
#pragma unroll
            for (int u = 0; u < LTK / 32; ++u) {
                const int jj = lane + u * 32, j = j0 + jj;
                if (j < Tn) {
                    const T* kr = Kt + jj * ds;
                    const T* er = Et + (jj + LQB - 1 - row) * ds;
                    float acc = 0.0f;
                    for (int c = 0; c < d; ++c) acc = fmaf(q[c], ldf(kr + c), fmaf(qv[c], ldf(er + c), acc));
                    acc *= scale;
                    if (j >= kl || i >= qlen) acc += -1e9f;
                    ps[(size_t)row * Tn + j] = acc;
                }
            }
        }
    }
    __syncwarp();
    // ---- softmax (a warp only touches its own rows of ps)
    for (int rr = 0; rr < LRPW; ++rr) {
        const int row = warp * LRPW + rr, i = i0 + row;
        if (i >= Tn) continue;
        float* p = ps + (size_t)row * Tn;
        float mx = -INFINITY;
        for (int j = lane; j < Tn; j += 32) mx = fmaxf(mx, p[j]);
        mx = warp_max(mx);
        float sum = 0.0f;
        for (int j = lane; j < Tn; j += 32) { const float ev = __expf(p[j] - mx); p[j] = ev; sum += ev; }
        sum = warp_sum(sum);
        const float inv = 1.0f / sum;
        float* prow = probs + (((size_t)b * H + h) * Tn + i) * Tn;
        for (int j = lane; j < Tn; j += 32) { const float pv = p[j] * inv; p[j] = pv; prow[j] = pv; }
    }
    // ---- o = P V
    float acc[LRPW][MAX_CPL];
#pragma unroll
    for (int rr = 0; rr < LRPW; ++rr)
#pragma unroll
        for (int u = 0; u < MAX_CPL; ++u) acc[rr][u] = 0.0f;
    for (int j0 = 0; j0 < Tn; j0 += LTK) {
        __syncthreads();
        stage_qkv(Kt, qkv_b, j0, LTK, 2, Tn, h, d, ds, G, D1, Tf);
        __syncthreads();
        const int cnt = min(LTK, Tn - j0);
#pragma unroll
        for (int rr = 0; rr < LRPW; ++rr) {
            const int row = warp * LRPW + rr;
            if (i0 + row >= Tn) continue;
            const float* p = ps + (size_t)row * Tn + j0;
            for (int jj = 0; jj < cnt; ++jj) {
                const float pj = p[jj];
                const T* vr = Kt + jj * ds;
#pragma unroll
                for (int u = 0; u < MAX_CPL; ++u) { const int c = lane + u * 32; if (c < d) acc[rr][u] = fmaf(pj, ldf(vr + c), acc[rr][u]); }
            }
        }
    }
#pragma unroll
    for (int rr = 0; rr < LRPW; ++rr) {
        const int i = i0 + warp * LRPW + rr;
        if (i >= Tn) continue;
#pragma unroll
        for (int u = 0; u < MAX_CPL; ++u) {
            const int c = lane + u * 32;
            if (c < d) {
                const int ee = h * d + c, fi = ee / D1, frame = i * G + fi;
                if (frame < Tf) stf(o + ((size_t)b * Tf + frame) * D1 + (ee - fi * D1), acc[rr][u]);
            }
        }
    }
}

// Backward, kernel A (per 32 query rows): dP = dO V^T, delta, dS (kept in shared memory and written to ds_ws), dQ, du / dv.
template <typename T>
__global__ void __launch_bounds__(ATT_THREADS) attn_long_bwd_q_kernel(
    const T* __restrict__ d_o, const T* __restrict__ qkv, const T* __restrict__ e, const float* __restrict__ probs,
    float* __restrict__ ds_ws, T* __restrict__ dqkv, int Tn, int H, int d, int G, int D1, int Tf, float* __restrict__ dub,
    float* __restrict__ dvb) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    const int ds = att_ds<T>(d), qds = d + 1;
    T* Kt = reinterpret_cast<T*>(sm_raw);                 // [LTK][ds]   V tile, later K tile
    T* Et = Kt + (size_t)LTK * ds;                        // [LTK+LQB-1][ds]
    float* os = reinterpret_cast<float*>(Et + (size_t)(LTK + LQB - 1) * ds);   // [LQB][qds] dO rows
    float* ps = os + (size_t)LQB * qds;                   // [LQB][Tn] dP -> dS
    const int b = blockIdx.y / H, h = blockIdx.y % H, i0 = blockIdx.x * LQB;
    const int D = H * d;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const T* qkv_b = qkv + (size_t)b * Tf * 3 * D1;
    T* dqkv_b = dqkv + (size_t)b * Tf * 3 * D1;
    const T* do_b = d_o + (size_t)b * Tf * D1;
    for (int idx = tid; idx < LQB * d; idx += ATT_THREADS) {
        const int rr = idx / d, c = idx - rr * d, i = i0 + rr;
        const int ee = h * d + c, fi = ee / D1, frame = i * G + fi;
        os[(size_t)rr * qds + c] = (i < Tn && frame < Tf) ? ldf(do_b + (size_t)frame * D1 + (ee - fi * D1)) : 0.0f;
    }
    const float scale = rsqrtf((float)d);
    // ---- dP_ij = dO_i . V_j
    for (int j0 = 0; j0 < Tn; j0 += LTK) {
        __syncthreads();
        stage_qkv(Kt, qkv_b, j0, LTK, 2, Tn, h, d, ds, G, D1, Tf);
        __syncthreads();
        for (int rr = 0; rr < LRPW; ++rr) {
            const int row = warp * LRPW + rr;
            if (i0 + row >= Tn) continue;
            const float* orow = os + (size_t)row * qds;
#pragma unroll
            for (int u = 0; u < LTK / 32; ++u) {
                const int jj = lane + u * 32, j = j0 + jj;
                if (j < Tn) {
                    const T* vr = Kt + jj * ds;
                    float acc = 0.0f;
                    for (int c = 0; c < d; ++c) acc = fmaf(orow[c], ldf(vr + c), acc);
                    ps[(size_t)row * Tn + j] = acc;
                }
            }
        }
    }
    __syncwarp();
    // ---- delta_i = sum_j P_ij dP_ij;  dS_ij = P_ij (dP_ij - delta_i) / sqrt(d)
    for (int rr = 0; rr < LRPW; ++rr) {
        const int row = warp * LRPW + rr, i = i0 + row;
        if (i >= Tn) continue;
        float* p = ps + (size_t)row * Tn;
        const float* Prow = probs + (((size_t)b * H + h) * Tn + i) * Tn;
        float* dSrow = ds_ws + (((size_t)b * H + h) * Tn + i) * Tn;
        float delta = 0.0f;
        for (int j = lane; j < Tn; j += 32) delta = fmaf(Prow[j], p[j], delta);
        delta = warp_sum(delta);
        for (int j = lane; j < Tn; j += 32) { const float v = Prow[j] * (p[j] - delta) * scale; p[j] = v; dSrow[j] = v; }
    }
    // ---- dQ_i = sum_j dS_ij (K_j + E_{T-1+j-i})
    float acc[LRPW][MAX_CPL], acce[LRPW][MAX_CPL];
#pragma unroll
    for (int rr = 0; rr < LRPW; ++rr)
#pragma unroll
        for (int u = 0; u < MAX_CPL; ++u) { acc[rr][u] = 0.0f; acce[rr][u] = 0.0f; }
    for (int j0 = 0; j0 < Tn; j0 += LTK) {
        __syncthreads();
        stage_qkv(Kt, qkv_b, j0, LTK, 1, Tn, h, d, ds, G, D1, Tf);
        stage_e(Et, e, i0, j0, Tn, h, d, ds, D);
        __syncthreads();
        const int cnt = min(LTK, Tn - j0);
#pragma unroll
        for (int rr = 0; rr < LRPW; ++rr) {
            const int row = warp * LRPW + rr;
            if (i0 + row >= Tn) continue;
            const float* p = ps + (size_t)row * Tn + j0;
            for (int jj = 0; jj < cnt; ++jj) {
                const float sv = p[jj];
                const T* kr = Kt + jj * ds;
                const T* er = Et + (jj + LQB - 1 - row) * ds;
#pragma unroll
                for (int u = 0; u < MAX_CPL; ++u) {
                    const int c = lane + u * 32;
                    if (c < d) { acc[rr][u] = fmaf(sv, ldf(kr + c), acc[rr][u]); acce[rr][u] = fmaf(sv, ldf(er + c), acce[rr][u]); }
                }
            }
        }
    }
    float du_acc[MAX_CPL], dv_acc[MAX_CPL];
#pragma unroll
    for (int u = 0; u < MAX_CPL; ++u) { du_acc[u] = 0.0f; dv_acc[u] = 0.0f; }
#pragma unroll
    for (int rr = 0; rr < LRPW; ++rr) {
        const int i = i0 + warp * LRPW + rr;
        if (i >= Tn) continue;
#pragma unroll
        for (int u = 0; u < MAX_CPL; ++u) {
            const int c = lane + u * 32;
            if (c < d) {
                const int ee = h * d + c, fi = ee / D1, frame = i * G + fi;
                if (frame < Tf) stf(dqkv_b + (size_t)frame * 3 * D1 + (ee - fi * D1), acc[rr][u] + acce[rr][u]);
                du_acc[u] += acc[rr][u];
                dv_acc[u] += acce[rr][u];
            }
        }
    }
#pragma unroll
    for (int u = 0; u < MAX_CPL; ++u) {
        const int c = lane + u * 32;
        if (c < d) {
            if (dub) atomicAdd(dub + (h * d + c) % D1, du_acc[u]);
            if (dvb) atomicAdd(dvb + (h * d + c) % D1, dv_acc[u]);
        }
    }
}

// Backward, kernel B (per 32 keys): dV_j = sum_i P_ij dO_i,  dK_j = sum_i dS_ij (Q_i + u)
template <typename T>
__global__ void __launch_bounds__(ATT_THREADS) attn_long_bwd_kv_kernel(
    const T* __restrict__ d_o, const T* __restrict__ qkv, const float* __restrict__ probs, const float* __restrict__ ds_ws,
    T* __restrict__ dqkv, int Tn, int H, int d, int G, int D1, int Tf, const float* __restrict__ ub) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    const int ds = att_ds<T>(d);
    T* Ot = reinterpret_cast<T*>(sm_raw);   // [LTK][ds] dO tile
    T* Qt = Ot + (size_t)LTK * ds;          // [LTK][ds] Q tile
    const int b = blockIdx.y / H, h = blockIdx.y % H, jb = blockIdx.x * LQB;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const T* qkv_b = qkv + (size_t)b * Tf * 3 * D1;
    T* dqkv_b = dqkv + (size_t)b * Tf * 3 * D1;
    const T* do_b = d_o + (size_t)b * Tf * D1;
    const float* P = probs + ((size_t)b * H + h) * Tn * Tn;
    const float* dS = ds_ws + ((size_t)b * H + h) * Tn * Tn;
    float ubv[MAX_CPL], accv[LRPW][MAX_CPL], acck[LRPW][MAX_CPL];
#pragma unroll
    for (int u = 0; u < MAX_CPL; ++u) {
        const int c = lane + u * 32;
        ubv[u] = (ub && c < d) ? ub[(h * d + c) % D1] : 0.0f;
#pragma unroll
        for (int rr = 0; rr < LRPW; ++rr) { accv[rr][u] = 0.0f; acck[rr][u] = 0.0f; }
    }
    for (int i0 = 0; i0 < Tn; i0 += LTK) {
        __syncthreads();
        stage_do(Ot, do_b, i0, LTK, Tn, h, d, ds, G, D1, Tf);
        stage_qkv(Qt, qkv_b, i0, LTK, 0, Tn, h, d, ds, G, D1, Tf);
        __syncthreads();
#pragma unroll
        for (int rr = 0; rr < LRPW; ++rr) {
            const int j = jb + warp * LRPW + rr;
            if (j >= Tn) continue;
            for (int ii0 = 0; ii0 < LTK && i0 + ii0 < Tn; ii0 += 32) {
                const int i = i0 + ii0 + lane;
                const float pij = i < Tn ? P[(size_t)i * Tn + j] : 0.0f;
                const float sij = i < Tn ? dS[(size_t)i * Tn + j] : 0.0f;
                const int cnt = min(32, Tn - (i0 + ii0));
                for (int ii = 0; ii < cnt; ++ii) {
                    const float pv = __shfl_sync(0xffffffffu, pij, ii), sv = __shfl_sync(0xffffffffu, sij, ii);
                    const T* orow = Ot + (ii0 + ii) * ds;
                    const T* qrow = Qt + (ii0 + ii) * ds;
#pragma unroll
                    for (int u = 0; u < MAX_CPL; ++u) {
                        const int c = lane + u * 32;
                        if (c < d) {
                            accv[rr][u] = fmaf(pv, ldf(orow + c), accv[rr][u]);
                            acck[rr][u] = fmaf(sv, ldf(qrow + c) + ubv[u], acck[rr][u]);
                        }
                    }
                }
            }
        }
    }
#pragma unroll
    for (int rr = 0; rr < LRPW; ++rr) {
        const int j = jb + warp * LRPW + rr;
        if (j >= Tn) continue;
#pragma unroll
        for (int u = 0; u < MAX_CPL; ++u) {
            const int c = lane + u * 32;
            if (c < d) {
                const int ee = h * d + c, fi = ee / D1, frame = j * G + fi;
                if (frame < Tf) {
                    stf(dqkv_b + (size_t)frame * 3 * D1 + 2 * D1 + (ee - fi * D1), accv[rr][u]);
                    stf(dqkv_b + (size_t)frame * 3 * D1 + 1 * D1 + (ee - fi * D1), acck[rr][u]);
                }
            }
        }
    }
}

// Backward, kernel C (per 32 relative offsets r): dE_r += sum_{i, j = r-(T-1)+i in [0,T)} dS_ij (Q_i + v)   (atomics over b)
template <typename T>
__global__ void __launch_bounds__(ATT_THREADS) attn_long_bwd_e_kernel(
    const T* __restrict__ qkv, const float* __restrict__ ds_ws, float* __restrict__ de, int Tn, int H, int d, int G, int D1, int Tf,
    const float* __restrict__ vb) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    const int ds = att_ds<T>(d);
    T* Qt = reinterpret_cast<T*>(sm_raw);   // [LTK][ds]
    const int b = blockIdx.y / H, h = blockIdx.y % H, rb = blockIdx.x * LQB;
    const int D = H * d;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const T* qkv_b = qkv + (size_t)b * Tf * 3 * D1;
    const float* dS = ds_ws + ((size_t)b * H + h) * Tn * Tn;
    float vbv[MAX_CPL], acc[LRPW][MAX_CPL];
#pragma unroll
    for (int u = 0; u < MAX_CPL; ++u) {
        const int c = lane + u * 32;
        vbv[u] = (vb && c < d) ? vb[(h * d + c) % D1] : 0.0f;
#pragma unroll
        for (int rr = 0; rr < LRPW; ++rr) acc[rr][u] = 0.0f;
    }
    // query rows any offset of this block can touch: i in [max(0, T-1-rmax), min(T-1, 2T-2-rmin)]
    const int rmin = rb, rmax = min(rb + LQB - 1, 2 * Tn - 2);
    const int iblo = max(0, Tn - 1 - rmax), ibhi = min(Tn - 1, 2 * Tn - 2 - rmin);
    for (int i0 = (iblo / LTK) * LTK; i0 <= ibhi; i0 += LTK) {
        __syncthreads();
        stage_qkv(Qt, qkv_b, i0, LTK, 0, Tn, h, d, ds, G, D1, Tf);
        __syncthreads();
#pragma unroll
        for (int rr = 0; rr < LRPW; ++rr) {
            const int r = rb + warp * LRPW + rr;
            if (r > 2 * Tn - 2) continue;
            const int lo = max(max(0, Tn - 1 - r), i0), hi = min(min(Tn - 1, 2 * Tn - 2 - r), i0 + LTK - 1);
            for (int ib = lo; ib <= hi; ib += 32) {
                const int i = ib + lane;
                const float sij = i <= hi ? dS[(size_t)i * Tn + (r - (Tn - 1) + i)] : 0.0f;
                const int cnt = min(32, hi - ib + 1);
                for (int ii = 0; ii < cnt; ++ii) {
                    const float sv = __shfl_sync(0xffffffffu, sij, ii);
                    const T* qrow = Qt + (ib + ii - i0) * ds;
#pragma unroll
                    for (int u = 0; u < MAX_CPL; ++u) { const int c = lane + u * 32; if (c < d) acc[rr][u] = fmaf(sv, ldf(qrow + c) + vbv[u], acc[rr][u]); }
                }
            }
        }
    }
#pragma unroll
    for (int rr = 0; rr < LRPW; ++rr) {
        const int r = rb + warp * LRPW + rr;
        if (r > 2 * Tn - 2) continue;
#pragma unroll
        for (int u = 0; u < MAX_CPL; ++u) { const int c = lane + u * 32; if (c < d) atomicAdd(de + (size_t)r * D + h * d + c, acc[rr][u]); }
    }
}

template <typename T> size_t long_q_smem(int Tn, int d, int qrows) {   // qrows: fp32 rows of d+1 next to the tiles (2 per query row fwd, 1 bwd)
    return (size_t)(2 * LTK + LQB - 1) * att_ds<T>(d) * sizeof(T) + ((size_t)LQB * qrows * (d + 1) + (size_t)LQB * Tn) * sizeof(float);
}

template <typename K> bool set_smem(K kfn, size_t smem) {
    return cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess;
}

}  // namespace

int avec_attn_long_fwd(const void* qkv, const void* e, const int* klen, int qlen, void* o, float* probs, int B, int T, int H, int d, int G,
                       int Tf, const float* u, const float* v, int dtype, cudaStream_t st) {
    if (d > 32 * MAX_CPL || (long long)B * H > 65535) return AVEC_ERR_UNSUPPORTED;
    const int D1 = H * d / G;
    dim3 grid(cdiv(T, LQB), B * H);
    AVEC_DISPATCH_DTYPE(dtype, Tt, {
        const size_t smem = long_q_smem<Tt>(T, d, 2);
        if (smem > 227 * 1024) return AVEC_ERR_UNSUPPORTED;
        auto kfn = attn_long_fwd_kernel<Tt>;
        if (!set_smem(kfn, smem)) return AVEC_ERR_LAUNCH;
        kfn<<<grid, ATT_THREADS, smem, st>>>((const Tt*)qkv, (const Tt*)e, klen, qlen, (Tt*)o, probs, T, H, d, G, D1, Tf, u, v);
    });
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}

int avec_attn_long_bwd(const void* d_o, const void* qkv, const void* e, const float* probs, float* ds_ws, void* dqkv, float* de, int B, int T,
                       int H, int d, int G, int Tf, const float* u, const float* v, float* du, float* dv, int dtype, cudaStream_t st) {
    if (d > 32 * MAX_CPL || (long long)B * H > 65535) return AVEC_ERR_UNSUPPORTED;
    const int D1 = H * d / G;
    AVEC_DISPATCH_DTYPE(dtype, Tt, {
        const size_t smem_q = long_q_smem<Tt>(T, d, 1);
        const size_t smem_t = (size_t)2 * LTK * att_ds<Tt>(d) * sizeof(Tt);
        if (smem_q > 227 * 1024) return AVEC_ERR_UNSUPPORTED;
        auto kq = attn_long_bwd_q_kernel<Tt>;
        auto kkv = attn_long_bwd_kv_kernel<Tt>;
        auto ke = attn_long_bwd_e_kernel<Tt>;
        if (!set_smem(kq, smem_q) || !set_smem(kkv, smem_t) || !set_smem(ke, smem_t)) return AVEC_ERR_LAUNCH;
        kq<<<dim3(cdiv(T, LQB), B * H), ATT_THREADS, smem_q, st>>>((const Tt*)d_o, (const Tt*)qkv, (const Tt*)e, probs, ds_ws, (Tt*)dqkv, T, H,
                                                                    d, G, D1, Tf, du, dv);
        avec_count_launch();
        kkv<<<dim3(cdiv(T, LQB), B * H), ATT_THREADS, smem_t, st>>>((const Tt*)d_o, (const Tt*)qkv, probs, ds_ws, (Tt*)dqkv, T, H, d, G, D1, Tf, u);
        avec_count_launch();
        ke<<<dim3(cdiv(2 * T - 1, LQB), B * H), ATT_THREADS, smem_t, st>>>((const Tt*)qkv, ds_ws, de, T, H, d, G, D1, Tf, v);
    });
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}
