// CTC negative log-likelihood and its gradient w.r.t. the logits, fused with log-softmax.
// Replaces log_softmax + transpose + nn.CTCLoss(reduction="none") of the reference's CTCLoss.forward
// (nnet/losses.py:292-334) and their autograd backward.  One CTA per utterance: alpha recursion over time with one
// thread per extended-label state, alphas kept in a caller-provided workspace, beta recursion fused with the gradient:
//   d nll / d logit[t,c] = softmax[t,c] - exp(logsumexp_{s: l'_s = c}(alpha[t,s] + beta[t,s]) + nll - logp[t,c])
// Lengths stay on the device (no host sync - the whole training step can be captured in a CUDA graph).
#include "common.cuh"

namespace {

constexpr int CTC_THREADS = 256;
constexpr float NEG_INF = -INFINITY;

__device__ __forceinline__ float logaddexp2f_(float a, float b) {
    if (a == NEG_INF) return b;
    if (b == NEG_INF) return a;
    float m = fmaxf(a, b);
    return m + log1pf(__expf(-fabsf(a - b)));
}

__global__ void __launch_bounds__(CTC_THREADS) ctc_kernel(const float* __restrict__ logits, const long long* __restrict__ labels,
                                                          const long long* __restrict__ in_len, const long long* __restrict__ lab_len,
                                                          float* __restrict__ nll, float* __restrict__ grad, float* __restrict__ ws,
                                                          int T, int V, int Lmax, int blank, int zero_infinity) {
    extern __shared__ float sm[];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = CTC_THREADS / 32;
    const int Tin = in_len ? (int)min((long long)T, in_len[b]) : T;
    const int Lb = (int)min((long long)Lmax, lab_len[b]);
    const int S = 2 * Lb + 1, Smax = 2 * Lmax + 1;
    float* lse = sm;                 // [T]
    float* beta = lse + T;           // [2][Smax]
    int* lab = reinterpret_cast<int*>(beta + 2 * Smax);   // [Smax] extended labels
    float* red = reinterpret_cast<float*>(lab + Smax);    // [2]
    const float* lg = logits + (size_t)b * T * V;
    float* gr = grad + (size_t)b * T * V;
    float* alpha = ws + (size_t)b * T * Smax;   // [T][Smax]
    // labels outside [0, V) cannot be reported from the device: they are clamped (never an out-of-bounds read of the logits row)
    for (int s = tid; s < Smax; s += CTC_THREADS) {
        long long l = (s & 1) ? labels[(size_t)b * Lmax + (s >> 1)] : (long long)blank;
        lab[s] = (int)min(max(l, 0LL), (long long)V - 1);
    }
    // ---- log-softmax denominators and softmax rows (the gradient's first term)
    for (int t = warp; t < T; t += nw) {
        const float* row = lg + (size_t)t * V;
        float mx = NEG_INF;
        for (int c = lane; c < V; c += 32) mx = fmaxf(mx, row[c]);
        mx = warp_max(mx);
        float sum = 0.0f;
        for (int c = lane; c < V; c += 32) sum += __expf(row[c] - mx);
        sum = warp_sum(sum);
        const float l = mx + logf(sum);
        if (lane == 0) lse[t] = l;
        for (int c = lane; c < V; c += 32) gr[(size_t)t * V + c] = t < Tin ? __expf(row[c] - l) : 0.0f;
    }
    __syncthreads();
    // ---- alpha recursion (thread = state)
    for (int t = 0; t < Tin; ++t) {
        for (int s = tid; s < S; s += CTC_THREADS) {
            float a;
            if (t == 0) {
                a = (s < 2) ? 0.0f : NEG_INF;
            } else {
                const float* ap = alpha + (size_t)(t - 1) * Smax;
                a = ap[s];
                if (s >= 1) a = logaddexp2f_(a, ap[s - 1]);
                if (s >= 2 && (s & 1) && lab[s] != lab[s - 2]) a = logaddexp2f_(a, ap[s - 2]);
            }
            alpha[(size_t)t * Smax + s] = a == NEG_INF ? NEG_INF : a + lg[(size_t)t * V + lab[s]] - lse[t];
        }
        __syncthreads();
    }
    if (tid == 0) {
        float l = NEG_INF;
        if (Tin > 0) {
            const float* al = alpha + (size_t)(Tin - 1) * Smax;
            l = al[S - 1];
            if (S >= 2) l = logaddexp2f_(l, al[S - 2]);
        }
        red[0] = -l;
    }
    __syncthreads();
    float loss = red[0];
    const bool infeasible = !(loss < INFINITY);   // inf or nan
    if (infeasible) {
        if (zero_infinity) {
            for (int i = tid; i < T * V; i += CTC_THREADS) gr[i] = 0.0f;
            if (tid == 0) nll[b] = 0.0f;
        } else {
            // torch.nn.functional.ctc_loss(zero_infinity=False): loss = inf and a NaN gradient for the whole utterance
            for (int i = tid; i < T * V; i += CTC_THREADS) gr[i] = __int_as_float(0x7fc00000);
            if (tid == 0) nll[b] = INFINITY;
        }
        return;
    }
    if (tid == 0) nll[b] = loss;
    // ---- beta recursion fused with the gradient's second term
    for (int t = Tin - 1; t >= 0; --t) {
        float* bc = beta + (t & 1) * Smax;
        const float* bn = beta + ((t + 1) & 1) * Smax;
        for (int s = tid; s < S; s += CTC_THREADS) {
            const float lp = lg[(size_t)t * V + lab[s]] - lse[t];
            float bt;
            if (t == Tin - 1) {
                bt = (s >= S - 2) ? 0.0f : NEG_INF;
            } else {
                bt = bn[s];
                if (s + 1 < S) bt = logaddexp2f_(bt, bn[s + 1]);
                if (s + 2 < S && (s & 1) && lab[s] != lab[s + 2]) bt = logaddexp2f_(bt, bn[s + 2]);
            }
            bt = bt == NEG_INF ? NEG_INF : bt + lp;
            bc[s] = bt;
            const float a = alpha[(size_t)t * Smax + s];
            // alpha and beta both include logp[t, l'_s]:  d nll / d logit[t,c] = y - (1 / (p * y)) * sum_{s: l'_s = c} alpha * beta
            if (a != NEG_INF && bt != NEG_INF) atomicAdd(gr + (size_t)t * V + lab[s], -__expf(a + bt - lp + loss));
        }
        __syncthreads();
    }
}

}  // namespace

extern "C" int avec_ctc_loss(const float* logits, const long long* labels, const long long* in_len, const long long* lab_len,
                             float* nll, float* grad, float* ws, int B, int T, int V, int Lmax, int blank, int zero_infinity,
                             avec_stream_t stream) {
    AVEC_CHECK_ARG(logits && labels && lab_len && nll && grad && ws && B > 0 && T > 0 && V > 0 && Lmax > 0 && blank >= 0 && blank < V);
    const int Smax = 2 * Lmax + 1;
    size_t smem = (size_t)(T + 2 * Smax) * sizeof(float) + (size_t)Smax * sizeof(int) + 16;
    if (smem > 200 * 1024) return AVEC_ERR_UNSUPPORTED;
    if (smem > 48 * 1024 && cudaFuncSetAttribute(ctc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return AVEC_ERR_LAUNCH;
    ctc_kernel<<<B, CTC_THREADS, smem, as_stream(stream)>>>(logits, labels, in_len, lab_len, nll, grad, ws, T, V, Lmax, blank, zero_infinity);
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}
