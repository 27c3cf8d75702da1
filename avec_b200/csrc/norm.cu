// LayerNorm (+ patch pooling), BatchNorm (channels-last), softmax, column reductions and the small elementwise glue of
// the ConformerBlock / ResNet paths.  All of these are HBM-bound: coalesced channel-contiguous accesses, fp32 math,
// one pass over the data per kernel, per-channel reductions pre-reduced in shared memory before the global atomics.
#include "common.cuh"
#include <algorithm>

namespace {

// ------------------------------------------------------------------------------------------------------------------
// LayerNorm forward: one warp per output token; the row lives in registers (C <= 32*LN_VPT).
// ------------------------------------------------------------------------------------------------------------------
constexpr int LN_G = 4;   // 4-channel groups per lane: C <= 32 * 4 * LN_G = 512, C % 4 == 0

template <typename T>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const T* __restrict__ x, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, T* __restrict__ y,
                                                            float* __restrict__ mean, float* __restrict__ rstd, int B,
                                                            int Tn, int Tp, int C, int P, float eps, long long ldy) {
    pdl_wait();      // (launched through avec_launch_pdl: nothing before this line touches global memory)
    pdl_trigger();   // the kernel behind this one may be scheduled now
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= B * Tp) return;
    const int b = warp / Tp, tp = warp % Tp;
    float out[LN_G][4];
#pragma unroll
    for (int u = 0; u < LN_G; ++u)
#pragma unroll
        for (int j = 0; j < 4; ++j) out[u][j] = 0.0f;
    for (int p = 0; p < P; ++p) {
        int t = tp * P + p;
        if (t >= Tn) break;
        const T* xr = x + ((size_t)b * Tn + t) * C;
        float v[LN_G][4];
        float s = 0.0f;
#pragma unroll
        for (int u = 0; u < LN_G; ++u) {
            int c = (lane + u * 32) * 4;
            if (c < C) load_vec<4>(xr + c, v[u]);
            else { v[u][0] = v[u][1] = v[u][2] = v[u][3] = 0.0f; }
            s += v[u][0] + v[u][1] + v[u][2] + v[u][3];
        }
        if (gamma == nullptr) {   // identity mode (attention.forwardQKV without the module's LayerNorm): plain patch mean
#pragma unroll
            for (int u = 0; u < LN_G; ++u)
#pragma unroll
                for (int j = 0; j < 4; ++j) out[u][j] += v[u][j];
            continue;
        }
        s = warp_sum(s);
        const float mu = s / C;
        float q = 0.0f;
#pragma unroll
        for (int u = 0; u < LN_G; ++u) {
            int c = (lane + u * 32) * 4;
            if (c < C) {
#pragma unroll
                for (int j = 0; j < 4; ++j) { float dv = v[u][j] - mu; q += dv * dv; }
            }
        }
        q = warp_sum(q);
        const float rs = rsqrtf(q / C + eps);
        if (lane == 0) { mean[(size_t)b * Tn + t] = mu; rstd[(size_t)b * Tn + t] = rs; }
#pragma unroll
        for (int u = 0; u < LN_G; ++u) {
            int c = (lane + u * 32) * 4;
            if (c < C) {
                float g[4], bb[4];
                load_vec<4>(gamma + c, g);
                load_vec<4>(beta + c, bb);
#pragma unroll
                for (int j = 0; j < 4; ++j) out[u][j] += (v[u][j] - mu) * rs * g[j] + bb[j];
            }
        }
    }
    const float invP = 1.0f / P;
    T* yr = y + ((size_t)b * Tp + tp) * ldy;
#pragma unroll
    for (int u = 0; u < LN_G; ++u) {
        int c = (lane + u * 32) * 4;
        if (c < C) {
#pragma unroll
            for (int j = 0; j < 4; ++j) out[u][j] *= invP;
            store_vec<4>(yr + c, out[u]);
        }
    }
}

// LayerNorm backward: warps stride over input rows; dgamma/dbeta partials live in registers, are reduced across the
// block's warps in shared memory and then added atomically to the fp32 gradient buffers.
template <typename T>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                            const float* __restrict__ gamma, const float* __restrict__ mean,
                                                            const float* __restrict__ rstd, const T* __restrict__ dres,
                                                            int res_stride, T* __restrict__ dx, float* __restrict__ dgamma,
                                                            float* __restrict__ dbeta, int B, int Tn, int Tp, int C, int P) {
    pdl_wait();      // (launched through avec_launch_pdl: nothing before this line touches global memory)
    pdl_trigger();   // the kernel behind this one may be scheduled now
    extern __shared__ float red[];  // [2][C]
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const long long rows = (long long)B * Tn;
    const int Tr = res_stride > 0 ? (Tn - 1) / res_stride + 1 : 0;
    float dg[LN_G][4], db[LN_G][4], gm[LN_G][4];
#pragma unroll
    for (int u = 0; u < LN_G; ++u) {
        int c = (lane + u * 32) * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) { dg[u][j] = 0.0f; db[u][j] = 0.0f; gm[u][j] = 0.0f; }
        if (c < C && gamma) load_vec<4>(gamma + c, gm[u]);
    }
    const float invP = 1.0f / P;
    for (long long row = (long long)blockIdx.x * wpb + wib; row < rows; row += (long long)gridDim.x * wpb) {
        const int b = (int)(row / Tn), t = (int)(row % Tn);
        const T* xr = x + row * C;
        const T* dyr = dy + ((size_t)b * Tp + t / P) * C;
        if (gamma == nullptr) {   // identity mode: dx = expand(dy) / P (+ dres)
            const bool hr = dres != nullptr && (t % res_stride) == 0;
            const T* r0 = hr ? dres + ((size_t)b * Tr + t / res_stride) * C : nullptr;
#pragma unroll
            for (int u = 0; u < LN_G; ++u) {
                int c = (lane + u * 32) * 4;
                if (c < C) {
                    float d[4];
                    load_vec<4>(dyr + c, d);
#pragma unroll
                    for (int j = 0; j < 4; ++j) d[j] *= invP;
                    if (hr) {
                        float r4[4];
                        load_vec<4>(r0 + c, r4);
#pragma unroll
                        for (int j = 0; j < 4; ++j) d[j] += r4[j];
                    }
                    store_vec<4>(dx + row * C + c, d);
                }
            }
            continue;
        }
        const float mu = mean[row], rs = rstd[row];
        float g[LN_G][4], xh[LN_G][4];
        float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
        for (int u = 0; u < LN_G; ++u) {
            int c = (lane + u * 32) * 4;
            if (c < C) {
                float d[4], xv[4];
                load_vec<4>(dyr + c, d);
                load_vec<4>(xr + c, xv);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float dd = d[j] * invP;
                    xh[u][j] = (xv[j] - mu) * rs;
                    dg[u][j] += dd * xh[u][j];
                    db[u][j] += dd;
                    g[u][j] = dd * gm[u][j];
                    s1 += g[u][j];
                    s2 += g[u][j] * xh[u][j];
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) { g[u][j] = 0.0f; xh[u][j] = 0.0f; }
            }
        }
        s1 = warp_sum(s1) / C;
        s2 = warp_sum(s2) / C;
        const bool has_res = dres != nullptr && (t % res_stride) == 0;
        const T* rr = has_res ? dres + ((size_t)b * Tr + t / res_stride) * C : nullptr;
        T* dxr = dx + row * C;
#pragma unroll
        for (int u = 0; u < LN_G; ++u) {
            int c = (lane + u * 32) * 4;
            if (c < C) {
                float o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) o[j] = rs * (g[u][j] - s1 - xh[u][j] * s2);
                if (has_res) {
                    float r4[4];
                    load_vec<4>(rr + c, r4);
#pragma unroll
                    for (int j = 0; j < 4; ++j) o[j] += r4[j];
                }
                store_vec<4>(dxr + c, o);
            }
        }
    }
    for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) red[c] = 0.0f;
    __syncthreads();
#pragma unroll
    for (int u = 0; u < LN_G; ++u) {
        int c = (lane + u * 32) * 4;
        if (c < C) {
#pragma unroll
            for (int j = 0; j < 4; ++j) { atomicAdd(&red[c + j], dg[u][j]); atomicAdd(&red[C + c + j], db[u][j]); }
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        if (dgamma) atomicAdd(dgamma + c, red[c]);
        if (dbeta) atomicAdd(dbeta + c, red[C + c]);
    }
}

template <typename T, int V>
__global__ void upsample_add_kernel(const T* __restrict__ x, const T* __restrict__ o, T* __restrict__ y, int Tn, int Tp, int C,
                                    int P, long long total) {
    pdl_wait();      // (launched through avec_launch_pdl: nothing before this line touches global memory)
    pdl_trigger();   // the kernel behind this one may be scheduled now
    const int Cv = C / V;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % Cv) * V;
        long long row = i / Cv;
        int t = (int)(row % Tn);
        long long b = row / Tn;
        float a[V], w[V];
        load_vec<V>(x + row * C + c, a);
        load_vec<V>(o + (b * Tp + t / P) * C + c, w);
#pragma unroll
        for (int j = 0; j < V; ++j) a[j] += w[j];
        store_vec<V>(y + row * C + c, a);
    }
}

template <typename T, int V>
__global__ void pool_sum_kernel(const T* __restrict__ dy, T* __restrict__ dout, int Tn, int Tp, int C, int P, long long total, long long ldo) {
    pdl_wait();      // (launched through avec_launch_pdl: nothing before this line touches global memory)
    pdl_trigger();   // the kernel behind this one may be scheduled now
    const int Cv = C / V;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % Cv) * V;
        long long row = i / Cv;
        int tp = (int)(row % Tp);
        long long b = row / Tp;
        float s[V];
#pragma unroll
        for (int j = 0; j < V; ++j) s[j] = 0.0f;
        for (int p = 0; p < P; ++p) {
            int t = tp * P + p;
            if (t < Tn) {
                float a[V];
                load_vec<V>(dy + (b * Tn + t) * C + c, a);
#pragma unroll
                for (int j = 0; j < V; ++j) s[j] += a[j];
            }
        }
        store_vec<V>(dout + row * ldo + c, s);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Row softmax (C <= 32*SM_VPT), warp per row.
// ------------------------------------------------------------------------------------------------------------------
constexpr int SM_VPT = 16;
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) softmax_fwd_kernel(const TI* __restrict__ x, TO* __restrict__ y, long long rows, int C) {
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    float v[SM_VPT];
    float mx = -INFINITY;
#pragma unroll
    for (int u = 0; u < SM_VPT; ++u) { int c = lane + u * 32; v[u] = c < C ? ldf(x + row * C + c) : -INFINITY; mx = fmaxf(mx, v[u]); }
    mx = warp_max(mx);
    float s = 0.0f;
#pragma unroll
    for (int u = 0; u < SM_VPT; ++u) { int c = lane + u * 32; v[u] = c < C ? __expf(v[u] - mx) : 0.0f; s += v[u]; }
    s = warp_sum(s);
    const float inv = 1.0f / s;
#pragma unroll
    for (int u = 0; u < SM_VPT; ++u) { int c = lane + u * 32; if (c < C) stf(y + row * C + c, v[u] * inv); }
}

template <typename T, typename TO>
__global__ void __launch_bounds__(256) softmax_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ y,
                                                          const float* __restrict__ dadd, TO* __restrict__ dx, long long rows, int C) {
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    float d[SM_VPT], p[SM_VPT];
    float s = 0.0f;
#pragma unroll
    for (int u = 0; u < SM_VPT; ++u) {
        int c = lane + u * 32;
        d[u] = c < C ? ldf(dy + row * C + c) : 0.0f;
        p[u] = c < C ? ldf(y + row * C + c) : 0.0f;
        s += d[u] * p[u];
    }
    s = warp_sum(s);
#pragma unroll
    for (int u = 0; u < SM_VPT; ++u) {
        int c = lane + u * 32;
        if (c < C) {
            float v = p[u] * (d[u] - s);
            if (dadd) v += dadd[row * C + c];
            stf(dx + row * C + c, v);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Column reductions over [rows, C] (C contiguous): thread x owns V consecutive channels, thread y strides rows.
// F(row, c0, v0[V], v1[V]) produces up to two values per channel to be summed.
// ------------------------------------------------------------------------------------------------------------------
constexpr int CR_THREADS = 512;

// blockDim = (X, 512 / X) with X = min(32, pow2ceil(C / V)): for narrow tensors (C = 64: X = 8) a warp then reads 4 whole
// consecutive rows (512 contiguous bytes) instead of leaving 3/4 of its lanes idle.
template <typename F, int V>
__global__ void __launch_bounds__(CR_THREADS) colreduce_kernel(F f_in, long long rows, int C, long long rows_per_block,
                                                               float* __restrict__ out0, float* __restrict__ out1, float alpha) {
    pdl_wait();      // (launched through avec_launch_pdl: nothing before this line touches global memory)
    pdl_trigger();   // the kernel behind this one may be scheduled now
    __shared__ float s0[CR_THREADS * V];
    __shared__ float s1[CR_THREADS * V];
    F f = f_in;
    const int X = blockDim.x, Y = blockDim.y;
    const int c = (blockIdx.x * X + threadIdx.x) * V;
    const long long r0 = (long long)blockIdx.y * rows_per_block;
    const long long r1 = min(rows, r0 + rows_per_block);
    float a0[V], a1[V];
#pragma unroll
    for (int j = 0; j < V; ++j) { a0[j] = 0.0f; a1[j] = 0.0f; }
    if (c < C) {
        f.prep(c);   // per-channel constants (this thread's channel group is fixed)
        long long r = r0 + threadIdx.y;
        // four rows per iteration: 4x the loads in flight per thread (these kernels are pure HBM streams)
        for (; r + 3LL * Y < r1; r += 4LL * Y) {
            float v0[4][V], v1[4][V];
#pragma unroll
            for (int q = 0; q < 4; ++q) f(r + (long long)q * Y, c, v0[q], v1[q]);
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int j = 0; j < V; ++j) { a0[j] += v0[q][j]; a1[j] += v1[q][j]; }
        }
        for (; r < r1; r += Y) {
            float v0[V], v1[V];
            f(r, c, v0, v1);
#pragma unroll
            for (int j = 0; j < V; ++j) { a0[j] += v0[j]; a1[j] += v1[j]; }
        }
    }
    const int W = X * V;   // channels covered by this block
#pragma unroll
    for (int j = 0; j < V; ++j) { s0[threadIdx.y * W + threadIdx.x * V + j] = a0[j]; s1[threadIdx.y * W + threadIdx.x * V + j] = a1[j]; }
    __syncthreads();
    // tree over y in shared memory, then one atomic per channel per block
    for (int h = Y >> 1; h >= 1; h >>= 1) {
        for (int i = threadIdx.y * X + threadIdx.x; i < h * W; i += CR_THREADS) { s0[i] += s0[i + h * W]; s1[i] += s1[i + h * W]; }
        __syncthreads();
    }
    for (int cc = threadIdx.y * X + threadIdx.x; cc < W; cc += CR_THREADS) {
        const int cg = blockIdx.x * W + cc;
        if (cg < C) {
            atomicAdd(out0 + cg, alpha * s0[cc]);
            if (out1) atomicAdd(out1 + cg, alpha * s1[cc]);
        }
    }
}

template <int V, typename F>
int launch_colreduce(const F& f, long long rows, int C, float* out0, float* out1, float alpha, cudaStream_t st) {
    const int Cv = cdiv(C, V);
    int X = 1;
    while (X < 32 && X < Cv) X <<= 1;
    const int Y = CR_THREADS / X;
    int gx = cdiv(Cv, X);
    long long want = cdivll(148 * 8, gx);
    long long nchunk = min(want, cdivll(rows, (long long)Y * 4));
    if (nchunk < 1) nchunk = 1;
    if (nchunk > 65535) nchunk = 65535;
    long long rpb = cdivll(rows, nchunk);
    nchunk = cdivll(rows, rpb);
    dim3 grid(gx, (unsigned)nchunk), block(X, Y);
    avec_launch_pdl(colreduce_kernel<F, V>, dim3(grid), dim3(block), 0, st, false, f, rows, C, rpb, out0, out1, alpha);
    return 0;
}

template <typename T, int V>
struct ColsumF {
    const T* x; long long ldx;
    __device__ __forceinline__ void prep(int) {}
    __device__ __forceinline__ void operator()(long long r, int c, float (&v0)[V], float (&v1)[V]) const {
        load_vec<V>(x + r * ldx + c, v0);
#pragma unroll
        for (int j = 0; j < V; ++j) v1[j] = 0.0f;
    }
};
template <typename T, int V>
struct StatsF {
    const T* x; int C;
    __device__ __forceinline__ void prep(int) {}
    __device__ __forceinline__ void operator()(long long r, int c, float (&v0)[V], float (&v1)[V]) const {
        load_vec<V>(x + r * C + c, v0);
#pragma unroll
        for (int j = 0; j < V; ++j) v1[j] = v0[j] * v0[j];
    }
};

__device__ __forceinline__ float act_fwd(float z, int act) { return act == AVEC_ACT_RELU ? fmaxf(z, 0.0f) : (act == AVEC_ACT_SWISH ? swishf_(z) : z); }
__device__ __forceinline__ float act_bwd(float z, int act) { return act == AVEC_ACT_RELU ? (z > 0.0f ? 1.0f : 0.0f) : (act == AVEC_ACT_SWISH ? dswishf_(z) : 1.0f); }

// gradient reaching input site (n, hi, wi) of a 3x3 / stride-2 / pad-1 max pool: dy of every window whose saved argmax code
// points at this site (code 255 = the ReLU floor won: no gradient)
template <typename T, int V>
__device__ __forceinline__ void pool_gather(const T* __restrict__ dy, const uint8_t* __restrict__ idx, long long n, int hi, int wi, int c,
                                            int C, int Ho, int Wo, float (&acc)[V]) {
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = 0.0f;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
        const int a = hi + 1 - kh;
        if (a < 0 || (a & 1)) continue;
        const int ho = a >> 1;
        if (ho >= Ho) continue;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
            const int b = wi + 1 - kw;
            if (b < 0 || (b & 1)) continue;
            const int wo = b >> 1;
            if (wo >= Wo) continue;
            const size_t o = ((size_t)(n * Ho + ho) * Wo + wo) * C + c;
            float d[V];
            load_vec<V>(dy + o, d);
            const int code = kh * 3 + kw;
            if (V == 8) {
                const uint2 w = *reinterpret_cast<const uint2*>(idx + o);
#pragma unroll
                for (int j = 0; j < V; ++j) { const uint32_t word = j < 4 ? w.x : w.y; if ((int)((word >> ((j & 3) * 8)) & 255u) == code) acc[j] += d[j]; }
            } else {
                const uint32_t w = *reinterpret_cast<const uint32_t*>(idx + o);
#pragma unroll
                for (int j = 0; j < V; ++j) if ((int)((w >> ((j & 3) * 8)) & 255u) == code) acc[j] += d[j];
            }
        }
    }
}

// BatchNorm backward reduction whose upstream gradient is the max-pool backward, computed on the fly (the [N,Hi,Wi,C]
// gradient tensor - 1.6 GB for the visual stem at B = 64 - is never written or re-read)
template <typename T, int V>
struct BnBwdPoolF {
    const T* dyp; const uint8_t* idx; const T* u; const float* mean; const float* rstd; int C, Hi, Wi, Ho, Wo;
    float mu[V], rs[V];
    __device__ __forceinline__ void prep(int c) { load_vec<V>(mean + c, mu); load_vec<V>(rstd + c, rs); }
    __device__ __forceinline__ void operator()(long long r, int c, float (&v0)[V], float (&v1)[V]) const {
        const int wi = (int)(r % Wi); const long long t = r / Wi; const int hi = (int)(t % Hi); const long long n = t / Hi;
        float uu[V];
        load_vec<V>(u + (size_t)r * C + c, uu);
        pool_gather<T, V>(dyp, idx, n, hi, wi, c, C, Ho, Wo, v0);
#pragma unroll
        for (int j = 0; j < V; ++j) v1[j] = v0[j] * (uu[j] - mu[j]) * rs[j];
    }
};

template <typename T, int V>
__global__ void bn_bwd_apply_pool_kernel(const T* __restrict__ dyp, const uint8_t* __restrict__ idx, const T* __restrict__ u,
                                         const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                                         const float* __restrict__ sums, T* __restrict__ du, long long totalv, int C, int Hi, int Wi, int Ho,
                                         int Wo, float inv_count) {
    const int Cv = C / V;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < totalv; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % Cv) * V;
        const long long r = i / Cv;
        const int wi = (int)(r % Wi); const long long t = r / Wi; const int hi = (int)(t % Hi); const long long n = t / Hi;
        float mu[V], rs[V], g[V], s0[V], s1[V], uu[V], dz[V], o[V];
        load_vec<V>(u + (size_t)i * V, uu);
        load_vec<V>(mean + c, mu); load_vec<V>(rstd + c, rs);
        load_vec<V>(sums + c, s0); load_vec<V>(sums + C + c, s1);
        if (gamma) load_vec<V>(gamma + c, g);
        pool_gather<T, V>(dyp, idx, n, hi, wi, c, C, Ho, Wo, dz);
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const float gg = (gamma ? g[j] : 1.0f) * rs[j];
            const float xh = (uu[j] - mu[j]) * rs[j];
            o[j] = gg * (dz[j] - s0[j] * inv_count - xh * s1[j] * inv_count);
        }
        store_vec<V>(du + (size_t)i * V, o);
    }
}

template <typename T, int V>
struct BnBwdF {
    const T* dy; const T* u; const T* res; const float* scale; const float* shift; const float* mean; const float* rstd; int C; int act;
    float sc[V], sh[V], mu[V], rs[V];
    __device__ __forceinline__ void prep(int c) {
        load_vec<V>(scale + c, sc);
        load_vec<V>(shift + c, sh);
        load_vec<V>(mean + c, mu);
        load_vec<V>(rstd + c, rs);
    }
    __device__ __forceinline__ void operator()(long long r, int c, float (&v0)[V], float (&v1)[V]) const {
        size_t i = (size_t)r * C + c;
        float uu[V], d[V], rr[V];
        load_vec<V>(u + i, uu);
        load_vec<V>(dy + i, d);
        if (res) load_vec<V>(res + i, rr);
#pragma unroll
        for (int j = 0; j < V; ++j) {
            float z = sc[j] * uu[j] + sh[j];
            if (res) z += rr[j];
            float dz = d[j] * act_bwd(z, act);
            v0[j] = dz;
            v1[j] = dz * (uu[j] - mu[j]) * rs[j];
        }
    }
};

// BatchNorm-backward column reduction (sum dz, sum dz * xhat), the heaviest of the column reductions (39 launches per AV step
// over up to 400 MB tensors).  Same block shape as colreduce_kernel, but the loop is memory-level-parallelism bound (3.7 TB/s with
// 12 loads in flight per thread, 2.7 TB/s with 8), so the raw 16-byte vectors of ROWS rows are all requested before the first
// one is consumed: 12 (no residual) / 15 (residual) 16-byte loads in flight per thread for bf16.
template <typename T, int V> struct alignas(16) RawVec { T v[V]; };
template <int V> struct alignas(16) RawVec<bf16, V> { uint32_t w[V / 2]; };   // packed pairs: unpacked with shifts, never addressed
template <int V> __device__ __forceinline__ void raw_unpack(const RawVec<float, V>& r, float (&f)[V]) {
#pragma unroll
    for (int j = 0; j < V; ++j) f[j] = r.v[j];
}
template <int V> __device__ __forceinline__ void raw_unpack(const RawVec<bf16, V>& r, float (&f)[V]) {
#pragma unroll
    for (int j = 0; j < V / 2; ++j) { f[2 * j] = __uint_as_float(r.w[j] << 16); f[2 * j + 1] = __uint_as_float(r.w[j] & 0xFFFF0000u); }
}

template <typename T, int V, int ROWS, bool HAS_RES>
__global__ void __launch_bounds__(CR_THREADS) bn_bwd_reduce_kernel(const T* __restrict__ dy, const T* __restrict__ u, const T* __restrict__ res,
                                                                   const float* __restrict__ scale, const float* __restrict__ shift,
                                                                   const float* __restrict__ mean, const float* __restrict__ rstd, long long rows, int C,
                                                                   long long rows_per_block, int act, float* __restrict__ out0, float* __restrict__ out1) {
    pdl_wait();      // (launched through avec_launch_pdl: nothing before this line touches global memory)
    __shared__ float s0[CR_THREADS * V];
    __shared__ float s1[CR_THREADS * V];
    const int X = blockDim.x, Y = blockDim.y;
    const int c = (blockIdx.x * X + threadIdx.x) * V;
    const long long r0 = (long long)blockIdx.y * rows_per_block;
    const long long r1 = min(rows, r0 + rows_per_block);
    float a0[V], a1[V];
#pragma unroll
    for (int j = 0; j < V; ++j) { a0[j] = 0.0f; a1[j] = 0.0f; }
    if (c < C) {
        // sum dz * xhat = rstd * (sum dz * u - mean * sum dz): the loop accumulates sum dz * u (mean / rstd applied once per
        // block below), which keeps 16 registers free for loads in flight
        float sc[V], sh[V];
        load_vec<V>(scale + c, sc); load_vec<V>(shift + c, sh);
        auto consume = [&](const RawVec<T, V>& ru, const RawVec<T, V>& rd, const RawVec<T, V>& rr) {
            float fu[V], fd[V], fr[V];
            raw_unpack(ru, fu); raw_unpack(rd, fd);
            if (HAS_RES) raw_unpack(rr, fr);
#pragma unroll
            for (int j = 0; j < V; ++j) {
                const float uu = fu[j];
                float z = sc[j] * uu + sh[j];
                if (HAS_RES) z += fr[j];
                const float dz = fd[j] * act_bwd(z, act);
                a0[j] += dz;
                a1[j] = fmaf(dz, uu, a1[j]);
            }
        };
        long long r = r0 + threadIdx.y;
        for (; r + (long long)(ROWS - 1) * Y < r1; r += (long long)ROWS * Y) {
            RawVec<T, V> ru[ROWS], rd[ROWS], rr[HAS_RES ? ROWS : 1];
#pragma unroll
            for (int q = 0; q < ROWS; ++q) {
                const size_t i = (size_t)(r + (long long)q * Y) * C + c;
                ru[q] = *reinterpret_cast<const RawVec<T, V>*>(u + i);
                rd[q] = *reinterpret_cast<const RawVec<T, V>*>(dy + i);
                if (HAS_RES) rr[q] = *reinterpret_cast<const RawVec<T, V>*>(res + i);
            }
#pragma unroll
            for (int q = 0; q < ROWS; ++q) consume(ru[q], rd[q], rr[HAS_RES ? q : 0]);
        }
        for (; r < r1; r += Y) {
            const size_t i = (size_t)r * C + c;
            RawVec<T, V> ru = *reinterpret_cast<const RawVec<T, V>*>(u + i), rd = *reinterpret_cast<const RawVec<T, V>*>(dy + i), rr = ru;
            if (HAS_RES) rr = *reinterpret_cast<const RawVec<T, V>*>(res + i);
            consume(ru, rd, rr);
        }
    }
    const int W = X * V;
#pragma unroll
    for (int j = 0; j < V; ++j) { s0[threadIdx.y * W + threadIdx.x * V + j] = a0[j]; s1[threadIdx.y * W + threadIdx.x * V + j] = a1[j]; }
    __syncthreads();
    for (int h = Y >> 1; h >= 1; h >>= 1) {
        for (int i = threadIdx.y * X + threadIdx.x; i < h * W; i += CR_THREADS) { s0[i] += s0[i + h * W]; s1[i] += s1[i + h * W]; }
        __syncthreads();
    }
    for (int cc = threadIdx.y * X + threadIdx.x; cc < W; cc += CR_THREADS) {
        const int cg = blockIdx.x * W + cc;
        if (cg < C) { atomicAdd(out0 + cg, s0[cc]); atomicAdd(out1 + cg, rstd[cg] * (s1[cc] - mean[cg] * s0[cc])); }
    }
}

template <typename T, int V>
int launch_bn_bwd_reduce(const T* dy, const T* u, const T* res, const float* scale, const float* shift, const float* mean, const float* rstd,
                         long long rows, int C, int act, float* sums, cudaStream_t st) {
    const int Cv = cdiv(C, V);
    int X = 1;
    while (X < 32 && X < Cv) X <<= 1;
    const int Y = CR_THREADS / X;
    const int gx = cdiv(Cv, X);
    constexpr bool HALF = sizeof(T) == 2;        // rows in flight are bounded by the 128-register budget of a 512-thread block
    constexpr int RR = HALF ? 5 : 2, RN = HALF ? 6 : 3;
    const int ROWS_ = res ? RR : RN;
    long long nchunk = std::min<long long>(cdivll(148 * 8, gx), cdivll(rows, (long long)Y * ROWS_));
    if (nchunk < 1) nchunk = 1;
    if (nchunk > 65535) nchunk = 65535;
    const long long rpb = cdivll(rows, nchunk);
    nchunk = cdivll(rows, rpb);
    dim3 grid(gx, (unsigned)nchunk), block(X, Y);
    if (res) avec_launch_pdl(bn_bwd_reduce_kernel<T, V, RR, true>, dim3(grid), dim3(block), 0, st, false, dy, u, res, scale, shift, mean, rstd, rows, C, rpb, act, sums, sums + C);
    else avec_launch_pdl(bn_bwd_reduce_kernel<T, V, RN, false>, dim3(grid), dim3(block), 0, st, false, dy, u, res, scale, shift, mean, rstd, rows, C, rpb, act, sums, sums + C);
    return 0;
}

__global__ void zero_kernel(float* p, int n) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = 0.0f; }

__global__ void bn_finalize_kernel(const float* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean,
                                   float* __restrict__ rstd, float* __restrict__ rmean, float* __restrict__ rvar,
                                   float inv_count, float unbias, int C, float eps, float momentum, int replicas) {
    pdl_wait();      // (launched through avec_launch_pdl: nothing before this line touches global memory)
    pdl_trigger();   // the kernel behind this one may be scheduled now
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float s1 = 0.0f, s2 = 0.0f;
    for (int r = 0; r < replicas; ++r) { s1 += stats[(size_t)r * 2 * C + c]; s2 += stats[(size_t)r * 2 * C + C + c]; }
    float mu = s1 * inv_count;
    float var = fmaxf(s2 * inv_count - mu * mu, 0.0f);
    float rs = rsqrtf(var + eps);
    float g = gamma ? gamma[c] : 1.0f, b = beta ? beta[c] : 0.0f;
    scale[c] = g * rs;
    shift[c] = b - mu * g * rs;
    if (mean) mean[c] = mu;
    if (rstd) rstd[c] = rs;
    if (rmean) rmean[c] = (1.0f - momentum) * rmean[c] + momentum * mu;
    if (rvar) rvar[c] = (1.0f - momentum) * rvar[c] + momentum * var * unbias;
}

__global__ void bn_eval_affine_kernel(const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ rmean,
                                      const float* __restrict__ rvar, float* __restrict__ scale, float* __restrict__ shift, int C, float eps) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float rs = rsqrtf(rvar[c] + eps);
    float g = gamma ? gamma[c] : 1.0f, b = beta ? beta[c] : 0.0f;
    scale[c] = g * rs;
    shift[c] = b - rmean[c] * g * rs;
}

template <typename T, int V>
__global__ void bn_apply_kernel(const T* __restrict__ u, const float* __restrict__ scale, const float* __restrict__ shift,
                                const T* __restrict__ res, T* __restrict__ y, long long totalv, int C, int act, long long ldy) {
    pdl_wait();      // (launched through avec_launch_pdl: nothing before this line touches global memory)
    const int Cv = C / V;
    // output element offset of flat (dense) element e: rows of y may be pitched (GEMM operands with TMA-able rows)
    auto yoff = [&](size_t e) -> size_t { return ldy == C ? e : (e / C) * (size_t)ldy + e % C; };
    const long long stride = (long long)gridDim.x * blockDim.x;
    const bool fixed_c = (stride % Cv) == 0;   // every iteration of this thread hits the same channel group
    float sc[V], sh[V];
    const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (fixed_c) { const int c = (int)(i0 % Cv) * V; load_vec<V>(scale + c, sc); load_vec<V>(shift + c, sh); }
    long long i = i0;
    if (fixed_c) {
        for (; i + stride < totalv; i += 2 * stride) {
            const size_t e0 = (size_t)i * V, e1 = (size_t)(i + stride) * V;
            float u0[V], u1[V], r0[V], r1[V];
            load_vec<V>(u + e0, u0); load_vec<V>(u + e1, u1);
            if (res) { load_vec<V>(res + e0, r0); load_vec<V>(res + e1, r1); }
#pragma unroll
            for (int j = 0; j < V; ++j) {
                float z0 = sc[j] * u0[j] + sh[j], z1 = sc[j] * u1[j] + sh[j];
                if (res) { z0 += r0[j]; z1 += r1[j]; }
                u0[j] = act_fwd(z0, act); u1[j] = act_fwd(z1, act);
            }
            store_vec<V>(y + yoff(e0), u0);
            store_vec<V>(y + yoff(e1), u1);
        }
    }
    for (; i < totalv; i += stride) {
        const int c = (int)(i % Cv) * V;
        const size_t e = (size_t)i * V;
        float uu[V], rr[V];
        load_vec<V>(u + e, uu);
        if (!fixed_c) { load_vec<V>(scale + c, sc); load_vec<V>(shift + c, sh); }
        if (res) load_vec<V>(res + e, rr);
#pragma unroll
        for (int j = 0; j < V; ++j) {
            float z = sc[j] * uu[j] + sh[j];
            if (res) z += rr[j];
            uu[j] = act_fwd(z, act);
        }
        store_vec<V>(y + yoff(e), uu);
    }
}

template <typename T, int V>
__global__ void bn_bwd_apply_kernel(const T* __restrict__ dy, const T* __restrict__ u, const float* __restrict__ scale,
                                    const float* __restrict__ shift, const T* __restrict__ res, const float* __restrict__ mean,
                                    const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ sums,
                                    T* __restrict__ du, T* __restrict__ dres, long long totalv, int C, int act, float inv_count) {
    pdl_wait();      // (launched through avec_launch_pdl: nothing before this line touches global memory)
    const int Cv = C / V;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const bool fixed_c = (stride % Cv) == 0;
    // du = k1*dz + k2*u + k3 with per-channel k1 = g*rs, k2 = -g*rs^2*s1/n, k3 = -g*rs*s0/n + g*rs^2*mu*s1/n
    float sc[V], sh[V], k1[V], k2[V], k3[V];
    auto load_coef = [&](int c) {
        float mu[V], rs[V], g[V], s0[V], s1[V];
        load_vec<V>(scale + c, sc);
        load_vec<V>(shift + c, sh);
        load_vec<V>(mean + c, mu);
        load_vec<V>(rstd + c, rs);
        load_vec<V>(sums + c, s0);
        load_vec<V>(sums + C + c, s1);
        if (gamma) load_vec<V>(gamma + c, g);
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const float gg = (gamma ? g[j] : 1.0f) * rs[j];
            k1[j] = gg;
            k2[j] = -gg * rs[j] * s1[j] * inv_count;
            k3[j] = -gg * s0[j] * inv_count + gg * rs[j] * mu[j] * s1[j] * inv_count;
        }
    };
    const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (fixed_c) load_coef((int)(i0 % Cv) * V);
    auto body = [&](const float (&uu)[V], const float (&d)[V], const float (&rr)[V], size_t e) {
        float o[V], dzv[V];
#pragma unroll
        for (int j = 0; j < V; ++j) {
            float z = sc[j] * uu[j] + sh[j];
            if (res) z += rr[j];
            float dz = d[j] * act_bwd(z, act);
            o[j] = k1[j] * dz + k2[j] * uu[j] + k3[j];
            dzv[j] = dz;
        }
        store_vec<V>(du + e, o);
        if (dres) store_vec<V>(dres + e, dzv);
    };
    long long i = i0;
    if (fixed_c) {
        // two elements per iteration with all six loads issued first: twice the bytes in flight per thread (pure HBM stream)
        for (; i + stride < totalv; i += 2 * stride) {
            const size_t e0 = (size_t)i * V, e1 = (size_t)(i + stride) * V;
            float u0[V], d0[V], r0[V], u1[V], d1[V], r1[V];
            load_vec<V>(u + e0, u0); load_vec<V>(u + e1, u1);
            load_vec<V>(dy + e0, d0); load_vec<V>(dy + e1, d1);
            if (res) { load_vec<V>(res + e0, r0); load_vec<V>(res + e1, r1); }
            body(u0, d0, r0, e0);
            body(u1, d1, r1, e1);
        }
    }
    for (; i < totalv; i += stride) {
        const int c = (int)(i % Cv) * V;
        const size_t e = (size_t)i * V;
        float uu[V], d[V], rr[V];
        load_vec<V>(u + e, uu);
        load_vec<V>(dy + e, d);
        if (!fixed_c) load_coef(c);
        if (res) load_vec<V>(res + e, rr);
        body(uu, d, rr, e);
    }
}

// BN + ReLU + MaxPool 3x3 / stride 2 / zero pad 1 (post-ReLU values are >= 0, so the zero padding never wins a strict max)
template <typename T, int V>
__global__ void bn_relu_maxpool_fwd_kernel(const T* __restrict__ u, const float* __restrict__ scale, const float* __restrict__ shift,
                                           T* __restrict__ y, uint8_t* __restrict__ idx, int Hi, int Wi, int C, int Ho, int Wo, long long totalv) {
    const int Cv = C / V;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < totalv; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % Cv) * V;
        long long t = i / Cv;
        int wo = (int)(t % Wo); t /= Wo;
        int ho = (int)(t % Ho);
        long long n = t / Ho;
        float best[V], sc[V], sh[V];
        int bi[V];
        load_vec<V>(scale + c, sc);
        load_vec<V>(shift + c, sh);
#pragma unroll
        for (int j = 0; j < V; ++j) { best[j] = 0.0f; bi[j] = 255; }   // the zero padding / ReLU floor
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            int hi = ho * 2 + kh - 1;
            if ((unsigned)hi >= (unsigned)Hi) continue;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                int wi = wo * 2 + kw - 1;
                if ((unsigned)wi >= (unsigned)Wi) continue;
                float uu[V];
                load_vec<V>(u + ((n * Hi + hi) * Wi + wi) * C + c, uu);
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    float z = sc[j] * uu[j] + sh[j];
                    if (z > best[j]) { best[j] = z; bi[j] = kh * 3 + kw; }
                }
            }
        }
        store_vec<V>(y + (size_t)i * V, best);
#pragma unroll
        for (int j = 0; j < V; ++j) idx[(size_t)i * V + j] = (uint8_t)bi[j];
    }
}

// gather-form backward: input site (hi,wi) receives dy of every window whose saved argmax is this site (z > 0 there)
template <typename T, int V>
__global__ void bn_relu_maxpool_bwd_kernel(const T* __restrict__ dy, const uint8_t* __restrict__ idx, T* __restrict__ dz, int Hi,
                                           int Wi, int C, int Ho, int Wo, long long totalv) {
    const int Cv = C / V;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < totalv; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % Cv) * V;
        long long t = i / Cv;
        int wi = (int)(t % Wi); t /= Wi;
        int hi = (int)(t % Hi);
        long long n = t / Hi;
        float acc[V];
#pragma unroll
        for (int j = 0; j < V; ++j) acc[j] = 0.0f;
        // windows ho with ho*2+kh-1 == hi  ->  kh = hi+1-2*ho in [0,3)
        for (int kh = 0; kh < 3; ++kh) {
            int a = hi + 1 - kh;
            if (a < 0 || (a & 1)) continue;
            int ho = a >> 1;
            if (ho >= Ho) continue;
            for (int kw = 0; kw < 3; ++kw) {
                int b = wi + 1 - kw;
                if (b < 0 || (b & 1)) continue;
                int wo = b >> 1;
                if (wo >= Wo) continue;
                size_t o = ((size_t)(n * Ho + ho) * Wo + wo) * C + c;
                float d[V];
                load_vec<V>(dy + o, d);
                const int code = kh * 3 + kw;
                if (V == 8) {
                    uint2 w = *reinterpret_cast<const uint2*>(idx + o);
#pragma unroll
                    for (int j = 0; j < V; ++j) { uint32_t word = j < 4 ? w.x : w.y; if ((int)((word >> ((j & 3) * 8)) & 255u) == code) acc[j] += d[j]; }
                } else {
                    uint32_t w = *reinterpret_cast<const uint32_t*>(idx + o);
#pragma unroll
                    for (int j = 0; j < V; ++j) if ((int)((w >> ((j & 3) * 8)) & 255u) == code) acc[j] += d[j];
                }
            }
        }
        store_vec<V>(dz + (size_t)i * V, acc);
    }
}

template <typename T>
__global__ void avgpool_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int HW, int C, long long total) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        long long n = i / C;
        float s = 0.0f;
        for (int p = 0; p < HW; ++p) s += ldf(x + (n * HW + p) * C + c);
        stf(y + i, s / HW);
    }
}
template <typename T>
__global__ void avgpool_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int HW, int C, long long total) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        long long n = i / ((long long)C * HW);
        stf(dx + i, ldf(dy + n * C + c) / HW);
    }
}

// zero insertion: out[n, i*s, j*s, :] = in[n, i, j, :], zeros elsewhere (turns the dgrad of a strided conv into a stride-1 conv)
template <typename T, int V>
__global__ void zero_upsample_kernel(const T* __restrict__ in, T* __restrict__ out, int Ho, int Wo, int Hi, int Wi, int C, int s, long long totalv) {
    const int Cv = C / V;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < totalv; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % Cv) * V;
        long long t = i / Cv;
        const int w = (int)(t % Wi); t /= Wi;
        const int h = (int)(t % Hi);
        const long long n = t / Hi;
        float v[V];
#pragma unroll
        for (int j = 0; j < V; ++j) v[j] = 0.0f;
        if (h % s == 0 && w % s == 0 && h / s < Ho && w / s < Wo) load_vec<V>(in + ((n * Ho + h / s) * Wo + w / s) * C + c, v);
        store_vec<V>(out + (size_t)i * V, v);
    }
}

template <typename TI, typename TO>
__global__ void convert_kernel(const TI* __restrict__ src, long long lds, TO* __restrict__ dst, long long ldd, int C, long long total) {
    pdl_wait();      // (launched through avec_launch_pdl: nothing before this line touches global memory)
    pdl_trigger();   // the kernel behind this one may be scheduled now
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        long long r = i / C;
        stf(dst + r * ldd + c, ldf(src + r * lds + c));
    }
}
// contiguous fp32 -> bf16, 8 elements per thread
__global__ void convert_f32_bf16_vec_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long long totalv) {
    pdl_wait();      // (launched through avec_launch_pdl: nothing before this line touches global memory)
    pdl_trigger();   // the kernel behind this one may be scheduled now
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < totalv; i += (long long)gridDim.x * blockDim.x) {
        float v[8];
        load_vec<8>(src + i * 8, v);
        store_vec<8>(dst + i * 8, v);
    }
}

inline int ew_blocks(long long total, int threads = 256) {
    long long b = cdivll(total, threads);
    long long cap = 148LL * 16;
    return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace

extern "C" int avec_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd, int B,
                                  int T, int C, int P, float eps, int dtype, long long ldy, avec_stream_t stream) {
    if (ldy <= 0) ldy = C;
    AVEC_CHECK_ARG(ldy >= C && ldy % 4 == 0);
    AVEC_CHECK_ARG(x && y && (gamma == nullptr || (beta && mean && rstd)) && B > 0 && T > 0 && P >= 1 && C > 0 && C <= 128 * LN_G && C % 4 == 0);
    const int Tp = cdiv(T, P);
    const long long warps = (long long)B * Tp;
    const int blocks = (int)cdivll(warps * 32, 256);
    AVEC_DISPATCH_DTYPE(dtype, Tt, (avec_launch_pdl(layernorm_fwd_kernel<Tt>, dim3(blocks), dim3(256), 0, as_stream(stream), false, 
        (const Tt*)x, gamma, beta, (Tt*)y, mean, rstd, B, T, Tp, C, P, eps, ldy)));
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}

extern "C" int avec_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                                  const void* dres, int res_stride, void* dx, float* dgamma, float* dbeta, int B, int T, int C,
                                  int P, int dtype, avec_stream_t stream) {
    AVEC_CHECK_ARG(dy && x && (gamma == nullptr || (mean && rstd)) && dx && B > 0 && T > 0 && P >= 1 && C > 0 && C <= 128 * LN_G && C % 4 == 0);
    AVEC_CHECK_ARG(!dres || res_stride >= 1);
    const int Tp = cdiv(T, P);
    const long long rows = (long long)B * T;
    // >= 4 rows per warp: every CTA ends with 2*C atomics onto the same dgamma / dbeta addresses, so few, longer CTAs
    // (Conformer sizes are 3-13 k rows: ~100-300 CTAs) beat one row per warp (profiles/r02_ncu_norm_kernels.md)
    int blocks = (int)std::min<long long>(std::max<long long>(cdivll(rows, 32), 1), 148LL * 2);
    size_t smem = 2 * (size_t)C * sizeof(float);
    AVEC_DISPATCH_DTYPE(dtype, Tt, (avec_launch_pdl(layernorm_bwd_kernel<Tt>, dim3(blocks), dim3(256), smem, as_stream(stream), false, 
        (const Tt*)dy, (const Tt*)x, gamma, mean, rstd, (const Tt*)dres, dres ? res_stride : 0, (Tt*)dx, dgamma, dbeta, B, T, Tp, C, P)));
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}

extern "C" int avec_upsample_add(const void* x, const void* o, void* y, int B, int T, int Tp, int C, int P, int dtype,
                                 avec_stream_t stream) {
    AVEC_CHECK_ARG(x && o && y && B > 0 && T > 0 && P >= 1 && Tp == cdiv(T, P));
    long long total = (long long)B * T * C;
    AVEC_DISPATCH_DTYPE_VEC(dtype, C, Tt, V, (avec_launch_pdl(upsample_add_kernel<Tt, V>, dim3(ew_blocks(total / V)), dim3(256), 0, as_stream(stream), false, 
        (const Tt*)x, (const Tt*)o, (Tt*)y, T, Tp, C, P, total / V)));
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}

extern "C" int avec_pool_sum(const void* dy, void* dout, int B, int T, int Tp, int C, int P, int dtype, long long ldo, avec_stream_t stream) {
    if (ldo <= 0) ldo = C;
    AVEC_CHECK_ARG(dy && dout && B > 0 && T > 0 && P >= 1 && Tp == cdiv(T, P) && ldo >= C && ldo % 4 == 0);
    long long total = (long long)B * Tp * C;
    AVEC_DISPATCH_DTYPE_VEC(dtype, C, Tt, V, (avec_launch_pdl(pool_sum_kernel<Tt, V>, dim3(ew_blocks(total / V)), dim3(256), 0, as_stream(stream), false, 
        (const Tt*)dy, (Tt*)dout, T, Tp, C, P, total / V, ldo)));
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}

extern "C" int avec_softmax_fwd(const void* x, int x_dtype, void* y, int y_dtype, long long rows, int C, avec_stream_t stream) {
    AVEC_CHECK_ARG(x && y && rows > 0 && C > 0 && C <= 32 * SM_VPT);
    const int blocks = (int)cdivll(rows * 32, 256);
    cudaStream_t st = as_stream(stream);
    if (x_dtype == AVEC_F32 && y_dtype == AVEC_F32) softmax_fwd_kernel<float, float><<<blocks, 256, 0, st>>>((const float*)x, (float*)y, rows, C);
    else if (x_dtype == AVEC_F32 && y_dtype == AVEC_BF16) softmax_fwd_kernel<float, bf16><<<blocks, 256, 0, st>>>((const float*)x, (bf16*)y, rows, C);
    else if (x_dtype == AVEC_BF16 && y_dtype == AVEC_BF16) softmax_fwd_kernel<bf16, bf16><<<blocks, 256, 0, st>>>((const bf16*)x, (bf16*)y, rows, C);
    else if (x_dtype == AVEC_BF16 && y_dtype == AVEC_F32) softmax_fwd_kernel<bf16, float><<<blocks, 256, 0, st>>>((const bf16*)x, (float*)y, rows, C);
    else return AVEC_ERR_INVALID;
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}

extern "C" int avec_softmax_bwd(const void* dy, const void* y, int dtype, const float* dadd, void* dx, int dx_dtype, long long rows,
                                int C, avec_stream_t stream) {
    AVEC_CHECK_ARG(dy && y && dx && rows > 0 && C > 0 && C <= 32 * SM_VPT);
    const int blocks = (int)cdivll(rows * 32, 256);
    cudaStream_t st = as_stream(stream);
    if (dtype == AVEC_F32 && dx_dtype == AVEC_F32) softmax_bwd_kernel<float, float><<<blocks, 256, 0, st>>>((const float*)dy, (const float*)y, dadd, (float*)dx, rows, C);
    else if (dtype == AVEC_BF16 && dx_dtype == AVEC_BF16) softmax_bwd_kernel<bf16, bf16><<<blocks, 256, 0, st>>>((const bf16*)dy, (const bf16*)y, dadd, (bf16*)dx, rows, C);
    else if (dtype == AVEC_BF16 && dx_dtype == AVEC_F32) softmax_bwd_kernel<bf16, float><<<blocks, 256, 0, st>>>((const bf16*)dy, (const bf16*)y, dadd, (float*)dx, rows, C);
    else if (dtype == AVEC_F32 && dx_dtype == AVEC_BF16) softmax_bwd_kernel<float, bf16><<<blocks, 256, 0, st>>>((const float*)dy, (const float*)y, dadd, (bf16*)dx, rows, C);
    else return AVEC_ERR_INVALID;
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}

extern "C" int avec_colsum(const void* x, int dtype, long long rows, int C, long long ldx, float alpha, float* out, int accumulate,
                           avec_stream_t stream) {
    AVEC_CHECK_ARG(x && out && rows > 0 && C > 0 && ldx >= C);
    cudaStream_t st = as_stream(stream);
    if (!accumulate) { zero_kernel<<<cdiv(C, 256), 256, 0, st>>>(out, C); avec_count_launch(); }
    if (C % 4 != 0 || ldx % 4 != 0) return AVEC_ERR_INVALID;
    if (C % 8 == 0 && ldx % 8 == 0) { AVEC_DISPATCH_DTYPE(dtype, Tt, { ColsumF<Tt, 8> f{(const Tt*)x, ldx}; launch_colreduce<8>(f, rows, C, out, nullptr, alpha, st); }); }
    else { AVEC_DISPATCH_DTYPE(dtype, Tt, { ColsumF<Tt, 4> f{(const Tt*)x, ldx}; launch_colreduce<4>(f, rows, C, out, nullptr, alpha, st); }); }
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}

extern "C" int avec_bn_stats(const void* u, int dtype, long long rows, int C, float* stats, avec_stream_t stream) {
    AVEC_CHECK_ARG(u && stats && rows > 0 && C > 0);
    cudaStream_t st = as_stream(stream);
    AVEC_DISPATCH_DTYPE_VEC(dtype, C, Tt, V, { StatsF<Tt, V> f{(const Tt*)u, C}; launch_colreduce<V>(f, rows, C, stats, stats + C, 1.0f, st); });
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}

extern "C" int avec_bn_finalize(const float* stats, const float* gamma, const float* beta, float* scale, float* shift, float* mean,
                                float* rstd, float* running_mean, float* running_var, long long count, int C, float eps,
                                float momentum, int replicas, avec_stream_t stream) {
    AVEC_CHECK_ARG(stats && scale && shift && count > 0 && C > 0 && replicas >= 1);
    float unbias = count > 1 ? (float)((double)count / (double)(count - 1)) : 1.0f;
    avec_launch_pdl(bn_finalize_kernel, dim3(cdiv(C, 128)), dim3(128), 0, as_stream(stream), false, stats, gamma, beta, scale, shift, mean, rstd, running_mean,
                                                                   running_var, (float)(1.0 / (double)count), unbias, C, eps, momentum, replicas);
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}

extern "C" int avec_bn_eval_affine(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                                   float* scale, float* shift, int C, float eps, avec_stream_t stream) {
    AVEC_CHECK_ARG(running_mean && running_var && scale && shift && C > 0);
    bn_eval_affine_kernel<<<cdiv(C, 128), 128, 0, as_stream(stream)>>>(gamma, beta, running_mean, running_var, scale, shift, C, eps);
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}

extern "C" int avec_bn_apply(const void* u, const float* scale, const float* shift, const void* res, void* y, long long rows, int C,
                             int act, int dtype, long long ldy, avec_stream_t stream) {
    if (ldy <= 0) ldy = C;
    AVEC_CHECK_ARG(u && scale && shift && y && rows > 0 && C > 0 && ldy >= C && ldy % 4 == 0);
    long long total = rows * C;
    AVEC_DISPATCH_DTYPE_VEC(dtype, C, Tt, V, (avec_launch_pdl(bn_apply_kernel<Tt, V>, dim3(ew_blocks(total / V)), dim3(256), 0, as_stream(stream), false, 
        (const Tt*)u, scale, shift, (const Tt*)res, (Tt*)y, total / V, C, act, ldy)));
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}

extern "C" int avec_bn_bwd_reduce(const void* dy, const void* u, const float* scale, const float* shift, const void* res,
                                  const float* mean, const float* rstd, float* sums, long long rows, int C, int act, int dtype,
                                  avec_stream_t stream) {
    AVEC_CHECK_ARG(dy && u && scale && shift && mean && rstd && sums && rows > 0 && C > 0);
    cudaStream_t st = as_stream(stream);
    AVEC_DISPATCH_DTYPE_VEC(dtype, C, Tt, V, {
        launch_bn_bwd_reduce<Tt, V>((const Tt*)dy, (const Tt*)u, (const Tt*)res, scale, shift, mean, rstd, rows, C, act, sums, st);
    });
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}

extern "C" int avec_bn_bwd_apply(const void* dy, const void* u, const float* scale, const float* shift, const void* res,
                                 const float* mean, const float* rstd, const float* gamma, const float* sums, void* du, void* dres,
                                 long long rows, int C, int act, int dtype, avec_stream_t stream) {
    AVEC_CHECK_ARG(dy && u && scale && shift && mean && rstd && sums && du && rows > 0 && C > 0);
    long long total = rows * C;
    AVEC_DISPATCH_DTYPE_VEC(dtype, C, Tt, V, (avec_launch_pdl(bn_bwd_apply_kernel<Tt, V>, dim3(ew_blocks(total / V)), dim3(256), 0, as_stream(stream), false, 
        (const Tt*)dy, (const Tt*)u, scale, shift, (const Tt*)res, mean, rstd, gamma, sums, (Tt*)du, (Tt*)dres, total / V, C, act,
        (float)(1.0 / (double)rows))));
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}

extern "C" int avec_bn_bwd_pool(const void* dyp, const uint8_t* idx, const void* u, const float* mean, const float* rstd, const float* gamma,
                                float* sums, void* du, int N, int Hi, int Wi, int C, int Ho, int Wo, int dtype, avec_stream_t stream) {
    AVEC_CHECK_ARG(dyp && idx && u && mean && rstd && sums && du && N > 0 && C > 0 && Ho == (Hi - 1) / 2 + 1 && Wo == (Wi - 1) / 2 + 1);
    cudaStream_t st = as_stream(stream);
    const long long rows = (long long)N * Hi * Wi, total = rows * C;
    AVEC_DISPATCH_DTYPE_VEC(dtype, C, Tt, V, {
        BnBwdPoolF<Tt, V> f{(const Tt*)dyp, idx, (const Tt*)u, mean, rstd, C, Hi, Wi, Ho, Wo, {}, {}};
        launch_colreduce<V>(f, rows, C, sums, sums + C, 1.0f, st);
        bn_bwd_apply_pool_kernel<Tt, V><<<ew_blocks(total / V), 256, 0, st>>>((const Tt*)dyp, idx, (const Tt*)u, mean, rstd, gamma, sums, (Tt*)du,
                                                                            total / V, C, Hi, Wi, Ho, Wo, (float)(1.0 / (double)rows));
    });
    avec_count_launch();
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}

extern "C" int avec_bn_relu_maxpool_fwd(const void* u, const float* scale, const float* shift, void* y, uint8_t* idx, int N, int Hi,
                                        int Wi, int C, int Ho, int Wo, int dtype, avec_stream_t stream) {
    AVEC_CHECK_ARG(u && scale && shift && y && idx && N > 0 && Ho == (Hi - 1) / 2 + 1 && Wo == (Wi - 1) / 2 + 1);
    long long total = (long long)N * Ho * Wo * C;
    AVEC_DISPATCH_DTYPE_VEC(dtype, C, Tt, V, (bn_relu_maxpool_fwd_kernel<Tt, V><<<ew_blocks(total / V), 256, 0, as_stream(stream)>>>(
        (const Tt*)u, scale, shift, (Tt*)y, idx, Hi, Wi, C, Ho, Wo, total / V)));
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}

extern "C" int avec_bn_relu_maxpool_bwd(const void* dy, const uint8_t* idx, void* dz, int N, int Hi, int Wi, int C, int Ho, int Wo,
                                        int dtype, avec_stream_t stream) {
    AVEC_CHECK_ARG(dy && idx && dz && N > 0);
    long long total = (long long)N * Hi * Wi * C;
    AVEC_DISPATCH_DTYPE_VEC(dtype, C, Tt, V, (bn_relu_maxpool_bwd_kernel<Tt, V><<<ew_blocks(total / V), 256, 0, as_stream(stream)>>>(
        (const Tt*)dy, idx, (Tt*)dz, Hi, Wi, C, Ho, Wo, total / V)));
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}

extern "C" int avec_avgpool_fwd(const void* x, void* y, int N, int HW, int C, int dtype, avec_stream_t stream) {
    AVEC_CHECK_ARG(x && y && N > 0 && HW > 0 && C > 0);
    long long total = (long long)N * C;
    AVEC_DISPATCH_DTYPE(dtype, Tt, (avgpool_fwd_kernel<Tt><<<ew_blocks(total), 256, 0, as_stream(stream)>>>((const Tt*)x, (Tt*)y, HW, C, total)));
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}
extern "C" int avec_avgpool_bwd(const void* dy, void* dx, int N, int HW, int C, int dtype, avec_stream_t stream) {
    AVEC_CHECK_ARG(dy && dx && N > 0 && HW > 0 && C > 0);
    long long total = (long long)N * HW * C;
    AVEC_DISPATCH_DTYPE(dtype, Tt, (avgpool_bwd_kernel<Tt><<<ew_blocks(total), 256, 0, as_stream(stream)>>>((const Tt*)dy, (Tt*)dx, HW, C, total)));
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}

extern "C" int avec_zero_upsample(const void* in, void* out, int N, int Ho, int Wo, int Hi, int Wi, int C, int s, int dtype,
                                  avec_stream_t stream) {
    AVEC_CHECK_ARG(in && out && N > 0 && s >= 1 && (Ho - 1) * s < Hi && (Wo - 1) * s < Wi);
    long long total = (long long)N * Hi * Wi * C;
    AVEC_DISPATCH_DTYPE_VEC(dtype, C, Tt, V, (zero_upsample_kernel<Tt, V><<<ew_blocks(total / V), 256, 0, as_stream(stream)>>>(
        (const Tt*)in, (Tt*)out, Ho, Wo, Hi, Wi, C, s, total / V)));
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}

// ---- multi-tensor strided copy / conversion: every per-step weight re-layout of a model in ONE launch ----------------------
// job j copies a logical 4-d index space (n0, n1, n2, n3) from src (fp32 or bf16, element strides ss) to dst (fp32 or bf16,
// element strides ds).  `chunks` maps every CTA to (job, chunk of AVEC_COPY_CHUNK elements of that job): no per-element search,
// 32-bit index arithmetic, and 4 elements per thread when the innermost dimension is contiguous on both sides.
constexpr int COPY_CHUNK = 4096;
__global__ void __launch_bounds__(256) convert_multi_kernel(const avec_copy_job* __restrict__ jobs, const int2* __restrict__ chunks) {
    const int2 bc = chunks[blockIdx.x];
    const avec_copy_job j = jobs[bc.x];
    const unsigned n1 = j.n[1], n2 = j.n[2], n3 = j.n[3];
    const unsigned numel = (unsigned)j.n[0] * n1 * n2 * n3;
    const unsigned base = (unsigned)bc.y * COPY_CHUNK, end = min(base + COPY_CHUNK, numel);
    const bool vec = j.ss[3] == 1 && j.ds[3] == 1 && (n3 & 3u) == 0 && j.src_dtype == AVEC_F32 &&
                     ((j.ss[0] | j.ss[1] | j.ss[2] | j.ds[0] | j.ds[1] | j.ds[2]) & 3) == 0 &&
                     (reinterpret_cast<uintptr_t>(j.src) & 15) == 0 && (reinterpret_cast<uintptr_t>(j.dst) & (j.dst_dtype == AVEC_F32 ? 15 : 7)) == 0;
    if (vec) {
        for (unsigned e = base + threadIdx.x * 4; e < end; e += 256 * 4) {
            unsigned r = e;
            const unsigned i3 = r % n3; r /= n3;
            const unsigned i2 = r % n2; r /= n2;
            const unsigned i1 = r % n1;
            const unsigned i0 = r / n1;
            const long long so = i0 * j.ss[0] + i1 * j.ss[1] + i2 * j.ss[2] + i3;
            const long long d_o = i0 * j.ds[0] + i1 * j.ds[1] + i2 * j.ds[2] + i3;
            const float4 v = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(j.src) + so);
            if (j.dst_dtype == AVEC_F32) *reinterpret_cast<float4*>(reinterpret_cast<float*>(j.dst) + d_o) = v;
            else {
                __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
                *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(j.dst) + d_o) = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
            }
        }
        return;
    }
    for (unsigned e = base + threadIdx.x; e < end; e += 256) {
        unsigned r = e;
        const unsigned i3 = r % n3; r /= n3;
        const unsigned i2 = r % n2; r /= n2;
        const unsigned i1 = r % n1;
        const unsigned i0 = r / n1;
        const long long so = i0 * j.ss[0] + i1 * j.ss[1] + i2 * j.ss[2] + i3 * j.ss[3];
        const long long d_o = i0 * j.ds[0] + i1 * j.ds[1] + i2 * j.ds[2] + i3 * j.ds[3];
        const float v = j.src_dtype == AVEC_F32 ? reinterpret_cast<const float*>(j.src)[so] : __bfloat162float(reinterpret_cast<const bf16*>(j.src)[so]);
        if (j.dst_dtype == AVEC_F32) reinterpret_cast<float*>(j.dst)[d_o] = v;
        else reinterpret_cast<bf16*>(j.dst)[d_o] = __float2bfloat16_rn(v);
    }
}

extern "C" int avec_convert_multi(const avec_copy_job* jobs_dev, const int* chunks_dev, int nchunks, avec_stream_t stream) {
    AVEC_CHECK_ARG(jobs_dev && chunks_dev && nchunks > 0);
    convert_multi_kernel<<<nchunks, 256, 0, as_stream(stream)>>>(jobs_dev, reinterpret_cast<const int2*>(chunks_dev));
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}

__global__ void __launch_bounds__(256) unpad_heads_kernel(const float* __restrict__ src, float* __restrict__ dst, int d, int dp, long long K, long long total) {
    pdl_wait();      // (launched through avec_launch_pdl: nothing before this line touches global memory)
    pdl_trigger();   // the kernel behind this one may be scheduled now
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long c = i % K, row = i / K;
        const long long g = row / d, r = row - g * d;
        dst[i] = src[(g * dp + r) * K + c];
    }
}
extern "C" int avec_unpad_heads(const float* src, float* dst, long long groups, int d, int dp, long long K, avec_stream_t stream) {
    AVEC_CHECK_ARG(src && dst && groups > 0 && d > 0 && dp >= d && K > 0);
    const long long total = groups * d * K;
    avec_launch_pdl(unpad_heads_kernel, dim3(ew_blocks(total)), dim3(256), 0, as_stream(stream), false, src, dst, d, dp, K, total);
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}

extern "C" int avec_convert(const void* src, int src_dtype, long long lds, void* dst, int dst_dtype, long long ldd, long long rows,
                            int C, avec_stream_t stream) {
    AVEC_CHECK_ARG(src && dst && rows > 0 && C > 0);
    long long total = rows * C;
    cudaStream_t st = as_stream(stream);
    int blocks = ew_blocks(total);
    if (src_dtype == AVEC_F32 && dst_dtype == AVEC_BF16 && lds == C && ldd == C && total % 8 == 0 &&
        ((uintptr_t)src & 31) == 0 && ((uintptr_t)dst & 15) == 0)
        avec_launch_pdl(convert_f32_bf16_vec_kernel, dim3(ew_blocks(total / 8)), dim3(256), 0, st, false, (const float*)src, (bf16*)dst, total / 8);
    else if (src_dtype == AVEC_F32 && dst_dtype == AVEC_BF16) avec_launch_pdl(convert_kernel<float, bf16>, dim3(blocks), dim3(256), 0, st, false, (const float*)src, lds, (bf16*)dst, ldd, C, total);
    else if (src_dtype == AVEC_BF16 && dst_dtype == AVEC_F32) avec_launch_pdl(convert_kernel<bf16, float>, dim3(blocks), dim3(256), 0, st, false, (const bf16*)src, lds, (float*)dst, ldd, C, total);
    else if (src_dtype == AVEC_F32 && dst_dtype == AVEC_F32) avec_launch_pdl(convert_kernel<float, float>, dim3(blocks), dim3(256), 0, st, false, (const float*)src, lds, (float*)dst, ldd, C, total);
    else if (src_dtype == AVEC_BF16 && dst_dtype == AVEC_BF16) avec_launch_pdl(convert_kernel<bf16, bf16>, dim3(blocks), dim3(256), 0, st, false, (const bf16*)src, lds, (bf16*)dst, ldd, C, total);
    else return AVEC_ERR_INVALID;
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}
