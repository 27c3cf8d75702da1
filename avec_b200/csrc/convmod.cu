// ConvolutionModule middle section: GLU -> depthwise Conv1d (k taps, stride s, zero "same" padding) -> BatchNorm batch
// statistics, forward and backward.  Reference: nnet/modules.py:374-377 (nn.GLU, depthwise layers.Conv1d with
// groups = channels, BatchNorm1d statistics over all B*T' positions, padded frames included).
// HBM-bound: one pass over the [B,T,2C] pre-activation, the GLU output is staged in shared memory with its halo so the
// k-tap window is read from SMEM; thread = channel so every global access is channel-contiguous (coalesced).
#include "common.cuh"

namespace {

constexpr int DW_TO = 8;     // output frames per CTA (short tiles: these launches are latency-, not bandwidth-bound)
constexpr int DW_CH = 128;   // channels per CTA (= threads)
constexpr int DW_MAXK = 15;
// backward tile: 16 output frames (the kernel is instruction bound: the 14-frame halo is amortised over twice as many outputs as
// in the forward, while the grid still fills the 148 SMs)
constexpr int DW_TO_BWD = 16;

template <typename T>
__global__ void __launch_bounds__(DW_CH) glu_dwconv_fwd_kernel(const T* __restrict__ pre, const float* __restrict__ w,
                                                               const float* __restrict__ bias, T* __restrict__ u,
                                                               float* __restrict__ stats, int Tn, int To, int C, int ksize,
                                                               int stride, int pad) {
    pdl_wait();      // (launched through avec_launch_pdl: nothing before this line touches global memory)
    pdl_trigger();   // the kernel behind this one may be scheduled now
    extern __shared__ float gs[];  // [rows][DW_CH]
    const int b = blockIdx.z, c = blockIdx.y * DW_CH + threadIdx.x;
    const int to0 = blockIdx.x * DW_TO;
    const int t_lo = to0 * stride - pad;
    const int rows = (DW_TO - 1) * stride + ksize;
    const bool cv = c < C;
    for (int r = 0; r < rows; ++r) {
        int t = t_lo + r;
        float g = 0.0f;
        if (cv && t >= 0 && t < Tn) {
            const T* p = pre + ((size_t)b * Tn + t) * 2 * C;
            g = ldf(p + c) * sigmoidf_(ldf(p + C + c));
        }
        gs[r * DW_CH + threadIdx.x] = g;
    }
    // each thread only reads its own column: no barrier needed
    if (!cv) return;
    float wk[DW_MAXK];
#pragma unroll
    for (int k = 0; k < DW_MAXK; ++k) wk[k] = k < ksize ? w[c * ksize + k] : 0.0f;
    const float bs = bias ? bias[c] : 0.0f;
    float s1 = 0.0f, s2 = 0.0f;
    for (int j = 0; j < DW_TO; ++j) {
        int to = to0 + j;
        if (to >= To) break;
        float acc = bs;
#pragma unroll
        for (int k = 0; k < DW_MAXK; ++k)
            if (k < ksize) acc = fmaf(wk[k], gs[(j * stride + k) * DW_CH + threadIdx.x], acc);
        stf(u + ((size_t)b * To + to) * C + c, acc);
        s1 += acc; s2 += acc * acc;
    }
    if (stats) { atomicAdd(stats + c, s1); atomicAdd(stats + C + c, s2); }
}

template <typename T, int STRIDE>
__global__ void __launch_bounds__(DW_CH) glu_dwconv_bwd_kernel(const T* __restrict__ du, const T* __restrict__ pre,
                                                               const float* __restrict__ w, T* __restrict__ dpre,
                                                               float* __restrict__ dw, float* __restrict__ db, int Tn, int To,
                                                               int C, int ksize, int stride_rt, int pad) {
    pdl_wait();      // (launched through avec_launch_pdl: nothing before this line touches global memory)
    pdl_trigger();   // the kernel behind this one may be scheduled now
    const int stride = STRIDE > 0 ? STRIDE : stride_rt;   // compile-time stride (1 / 2): no division in the tap loops
    extern __shared__ float sm[];
    const int b = blockIdx.z, c = blockIdx.y * DW_CH + threadIdx.x;
    const int to0 = blockIdx.x * DW_TO_BWD;
    const bool cv = c < C;
    // g tile: input frames [to0*s - pad, to0*s - pad + grows)
    const int g_lo = to0 * stride - pad;
    const int grows = (DW_TO_BWD - 1) * stride + ksize;
    // du tile: output frames needed by dg of input frames [t0, t0 + DW_TO_BWD*s): to in [floor((t0+pad-(k-1))/s), (t0+DW_TO_BWD*s-1+pad)/s]
    const int t0 = to0 * stride;
    const int d_lo = (t0 + pad - (ksize - 1) - (stride - 1)) / stride - 1;  // conservative lower bound (may be negative)
    const int d_hi = (t0 + DW_TO_BWD * stride - 1 + pad) / stride;
    const int drows = d_hi - d_lo + 1;
    float* gs = sm;                    // [grows][DW_CH]
    float* dsm = sm + grows * DW_CH;   // [drows][DW_CH]
    for (int r = 0; r < grows; ++r) {
        int t = g_lo + r;
        float g = 0.0f;
        if (cv && t >= 0 && t < Tn) {
            const T* p = pre + ((size_t)b * Tn + t) * 2 * C;
            g = ldf(p + c) * sigmoidf_(ldf(p + C + c));
        }
        gs[r * DW_CH + threadIdx.x] = g;
    }
    for (int r = 0; r < drows; ++r) {
        int to = d_lo + r;
        float v = 0.0f;
        if (cv && to >= 0 && to < To) v = ldf(du + ((size_t)b * To + to) * C + c);
        dsm[r * DW_CH + threadIdx.x] = v;
    }
    if (!cv) return;
    float wk[DW_MAXK], dwk[DW_MAXK];
#pragma unroll
    for (int k = 0; k < DW_MAXK; ++k) { wk[k] = k < ksize ? w[c * ksize + k] : 0.0f; dwk[k] = 0.0f; }
    // weight / bias gradients from this CTA's output frames
    float dbs = 0.0f;
    for (int j = 0; j < DW_TO_BWD; ++j) {
        int to = to0 + j;
        if (to >= To) break;
        float d = dsm[(to - d_lo) * DW_CH + threadIdx.x];
        dbs += d;
#pragma unroll
        for (int k = 0; k < DW_MAXK; ++k)
            if (k < ksize) dwk[k] = fmaf(d, gs[(j * stride + k) * DW_CH + threadIdx.x], dwk[k]);
    }
#pragma unroll
    for (int k = 0; k < DW_MAXK; ++k)
        if (k < ksize) atomicAdd(dw + c * ksize + k, dwk[k]);
    if (db) atomicAdd(db + c, dbs);
    // input gradients for this CTA's input frames, then back through the GLU
    for (int j = 0; j < DW_TO_BWD * stride; ++j) {
        int t = t0 + j;
        if (t >= Tn) break;
        float dg = 0.0f;
        if (STRIDE == 1) {
            // dg[t] = sum_k w[k] du[t + pad - k]; rows outside [0, To) of the staged tile are zero
#pragma unroll
            for (int k = 0; k < DW_MAXK; ++k) {
                const int to = t + pad - k;
                if (k < ksize) dg = fmaf(wk[k], dsm[(to - d_lo) * DW_CH + threadIdx.x], dg);   // d_lo <= to <= d_hi by construction
            }
        } else {
#pragma unroll
            for (int k = 0; k < DW_MAXK; ++k) {
                if (k < ksize) {
                    int a = t + pad - k;
                    if (a >= 0 && a % stride == 0) {
                        int to = a / stride;
                        if (to < To) dg = fmaf(wk[k], dsm[(to - d_lo) * DW_CH + threadIdx.x], dg);
                    }
                }
            }
        }
        const T* p = pre + ((size_t)b * Tn + t) * 2 * C;
        float v = ldf(p + c), gt = ldf(p + C + c);
        float sg = sigmoidf_(gt);
        T* q = dpre + ((size_t)b * Tn + t) * 2 * C;
        stf(q + c, dg * sg);
        stf(q + C + c, dg * v * sg * (1.0f - sg));
    }
}

}  // namespace

extern "C" int avec_glu_dwconv_fwd(const void* pre, const float* w, const float* bias, void* u, float* stats, int B, int T, int To,
                                   int C, int ksize, int stride, int pad, int dtype, avec_stream_t stream) {
    AVEC_CHECK_ARG(pre && w && u && B > 0 && T > 0 && C > 0 && ksize >= 1 && ksize <= DW_MAXK && stride >= 1);
    AVEC_CHECK_ARG(To == (T + 2 * pad - ksize) / stride + 1 && B <= 65535);
    dim3 grid(cdiv(To, DW_TO), cdiv(C, DW_CH), B);
    size_t smem = (size_t)((DW_TO - 1) * stride + ksize) * DW_CH * sizeof(float);
    AVEC_DISPATCH_DTYPE(dtype, Tt, {
        auto kfn = glu_dwconv_fwd_kernel<Tt>;
        if (smem > 48 * 1024 && cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return AVEC_ERR_LAUNCH;
        avec_launch_pdl(kfn, grid, dim3(DW_CH), smem, as_stream(stream), false, (const Tt*)pre, w, bias, (Tt*)u, stats, T, To, C, ksize, stride, pad);
    });
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}

extern "C" int avec_glu_dwconv_bwd(const void* du, const void* pre, const float* w, void* dpre, float* dw, float* db, int B, int T,
                                   int To, int C, int ksize, int stride, int pad, int dtype, avec_stream_t stream) {
    AVEC_CHECK_ARG(du && pre && w && dpre && dw && B > 0 && T > 0 && C > 0 && ksize >= 1 && ksize <= DW_MAXK && stride >= 1);
    AVEC_CHECK_ARG(To == (T + 2 * pad - ksize) / stride + 1 && B <= 65535);
    dim3 grid(cdiv(To, DW_TO_BWD), cdiv(C, DW_CH), B);
    const int grows = (DW_TO_BWD - 1) * stride + ksize;
    const int drows = DW_TO_BWD + (ksize + stride) / stride + 4;  // upper bound of d_hi - d_lo + 1 for any tile
    size_t smem = (size_t)(grows + drows) * DW_CH * sizeof(float);
    AVEC_DISPATCH_DTYPE(dtype, Tt, {
        auto kfn = stride == 1 ? glu_dwconv_bwd_kernel<Tt, 1> : (stride == 2 ? glu_dwconv_bwd_kernel<Tt, 2> : glu_dwconv_bwd_kernel<Tt, 0>);
        if (smem > 48 * 1024 && cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return AVEC_ERR_LAUNCH;
        avec_launch_pdl(kfn, grid, dim3(DW_CH), smem, as_stream(stream), false, (const Tt*)du, (const Tt*)pre, w, (Tt*)dpre, dw, db, T, To, C, ksize, stride, pad);
    });
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}
