// Training-step kernels around the encoder hot path:
//   counter-based dropout (Philox4x32-10; masks are regenerated in the backward, never stored),
//   on-device SpecAugment, greedy CTC decoding, fused Adam (L2 weight decay, Noam / constant LR, global-norm clip, EMA).
// Every random draw is a pure function of (seed, step, site, row, column group) read from DEVICE memory, so a captured
// CUDA graph draws fresh masks on every replay (the step counter is advanced by a kernel inside the graph).
#include "common.cuh"

// (Philox4x32-10, dropout_bits and bits16 live in common.cuh: the tcgen05 GEMM epilogue draws the same masks)

__global__ void counter_advance_kernel(unsigned long long* ctr) { ctr[0] += 1ull; }

extern "C" int avec_counter_advance(unsigned long long* counter, avec_stream_t stream) {
    AVEC_CHECK_ARG(counter);
    counter_advance_kernel<<<1, 1, 0, as_stream(stream)>>>(counter);
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}

// y[r][c] = (res ? res[r][c] : 0) + alpha * keep(r,c) * x[r'][c] / (1 - p);  keep <=> bits16 >= thresh;  r' = r, or the
// patch row of frame r when the dropout follows the patch attention's nearest-neighbour upsampling (attentions.py:368-372)
template <typename T, int V>
__global__ void __launch_bounds__(256) dropout_kernel(const T* __restrict__ x, const T* __restrict__ res, T* __restrict__ y,
                                                      long long rows, int C, uint32_t thresh, float scale,
                                                      const unsigned long long* __restrict__ rng, uint32_t site, int Tf, int Tp,
                                                      int P, long long ldy) {
    pdl_wait();      // (launched through avec_launch_pdl: nothing before this line touches global memory)
    pdl_trigger();   // the kernel behind this one may be scheduled now
    const int per_row = (C + V - 1) / V;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * per_row) return;
    const long long row = i / per_row;
    const int c0 = (int)(i - row * per_row) * V;
    const uint4 r = dropout_bits(rng, site, (uint32_t)row, (uint32_t)(c0 >> 3));
    const size_t o = (size_t)row * C + c0;
    // P > 1: x holds one row per patch of P frames ([B, Tp, C]) and is repeated over the frames of y ([B, T, C])
    const size_t ox = P > 1 ? (size_t)((row / Tf) * Tp + (row % Tf) / P) * C + c0 : o;
    const size_t oy = (size_t)row * ldy + c0;     // y rows may be pitched (GEMM operand with TMA-able rows); x / res are dense
    if constexpr (V == 1) {
        float v = bits16(r, c0 & 7) >= thresh ? ldf(x + ox) * scale : 0.0f;
        if (res) v += ldf(res + o);
        stf(y + oy, v);
    } else {
        float xv[V], rv[V];
        load_vec<V>(x + ox, xv);
        if (res) load_vec<V>(res + o, rv);
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const float v = bits16(r, (c0 + j) & 7) >= thresh ? xv[j] * scale : 0.0f;
            xv[j] = res ? rv[j] + v : v;
        }
        store_vec<V>(y + oy, xv);
    }
}

extern "C" int avec_dropout(const void* x, const void* res, void* y, long long rows, int C, int dtype, float p, float alpha,
                            const unsigned long long* rng_state, int site, int Tf, int Tp, int P, long long ldy, avec_stream_t stream) {
    if (ldy <= 0) ldy = C;
    AVEC_CHECK_ARG(x && y && rng_state && rows > 0 && C > 0 && p >= 0.0f && p < 1.0f && ldy >= C && (ldy == C || ldy % 4 == 0));
    AVEC_CHECK_ARG(P <= 1 || (Tf > 0 && Tp > 0 && rows % Tf == 0 && (Tf + P - 1) / P <= Tp && x != y));
    const uint32_t thresh = (uint32_t)(p * 65536.0f + 0.5f);
    const float scale = alpha / (1.0f - p);
    cudaStream_t st = as_stream(stream);
#define AVEC_DROP_LAUNCH(T, V)                                                                                         \
    do {                                                                                                               \
        const long long n = rows * ((C + V - 1) / V);                                                                  \
        avec_launch_pdl(dropout_kernel<T, V>, dim3((unsigned)cdivll(n, 256)), dim3(256), 0, st, false, (const T*)x, (const T*)res, (T*)y, rows, C, thresh, \
                                                                       scale, rng_state, (uint32_t)site, Tf, Tp, P, ldy); \
    } while (0)
    AVEC_DISPATCH_DTYPE(dtype, T, {
        if (C % 8 == 0) AVEC_DROP_LAUNCH(T, 8);
        else if (C % 4 == 0) AVEC_DROP_LAUNCH(T, 4);
        else AVEC_DROP_LAUNCH(T, 1);
    });
#undef AVEC_DROP_LAUNCH
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}

// ------------------------------------------------------------------------------------------------------ SpecAugment
// mel [B, F, M] fp32 (frame-major).  mF frequency masks shared by the whole batch (FrequencyMasking(iid_masks=False)),
// mT time masks per utterance over its own length with maximum width int(pS * len) (nnet/preprocessing.py:118-127,
// torchaudio mask_along_axis: width = int(u1 * param), start = int(u2 * (size - u1 * param)), zeros in [start, start+width)).
// Draw k of utterance b: Philox(ctr = (b, k, site, step)); frequency masks use b = 0xFFFFFFFF.
#define AVEC_SPEC_MAX_MASKS 16
__device__ __forceinline__ float u01(uint32_t w) { return (float)(w >> 8) * (1.0f / 16777216.0f); }

__device__ __forceinline__ void mask_interval(const unsigned long long* rng, uint32_t site, uint32_t b, uint32_t k, float param,
                                              float size, int* lo, int* hi) {
    const unsigned long long seed = rng[0], step = rng[1];
    const uint4 r = philox4x32_10(make_uint4(b, k, site, (uint32_t)step), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const float value = __fmul_rn(u01(r.x), param);                       // explicit roundings: no FMA contraction, so the
    const float minv = __fmul_rn(u01(r.y), __fsub_rn(size, value));       // numpy restatement reproduces the intervals exactly
    *lo = (int)minv;
    *hi = (int)minv + (int)value;
}

__global__ void __launch_bounds__(256) spec_augment_kernel(float* __restrict__ mel, const long long* __restrict__ lengths, int B,
                                                           int F, int M, int mF, int Fmax, int mT, float pS,
                                                           const unsigned long long* __restrict__ rng, uint32_t site,
                                                           int* __restrict__ intervals) {
    __shared__ int lo[AVEC_SPEC_MAX_MASKS], hi[AVEC_SPEC_MAX_MASKS];
    const int b = blockIdx.y;
    const int len = lengths ? (int)min((long long)F, lengths[b]) : F;
    if (threadIdx.x < mF) {
        const float param = (float)min(Fmax, M);
        if (param < 1.0f) { lo[threadIdx.x] = hi[threadIdx.x] = 0; }
        else mask_interval(rng, site, 0xFFFFFFFFu, threadIdx.x, param, (float)M, &lo[threadIdx.x], &hi[threadIdx.x]);
    } else if (threadIdx.x < mF + mT) {
        const int Tb = min((int)(pS * (float)len), len);
        if (Tb < 1) { lo[threadIdx.x] = hi[threadIdx.x] = 0; }
        else mask_interval(rng, site, (uint32_t)b, threadIdx.x - mF, (float)Tb, (float)len, &lo[threadIdx.x], &hi[threadIdx.x]);
    }
    __syncthreads();
    if (intervals && blockIdx.x == 0 && threadIdx.x < mF + mT) {
        intervals[((size_t)b * (mF + mT) + threadIdx.x) * 2] = lo[threadIdx.x];
        intervals[((size_t)b * (mF + mT) + threadIdx.x) * 2 + 1] = hi[threadIdx.x];
    }
    const int total = F * M;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int f = i / M, m = i - f * M;
        bool z = false;
        for (int k = 0; k < mF; ++k) z |= (m >= lo[k] && m < hi[k]);
        if (f < len) for (int k = mF; k < mF + mT; ++k) z |= (f >= lo[k] && f < hi[k]);
        if (z) mel[(size_t)b * total + i] = 0.0f;
    }
}

extern "C" int avec_spec_augment(float* mel, const long long* lengths, int B, int F, int M, int mF, int Fmax, int mT, float pS,
                                 const unsigned long long* rng_state, int site, int* intervals, avec_stream_t stream) {
    AVEC_CHECK_ARG(mel && rng_state && B > 0 && F > 0 && M > 0 && mF >= 0 && mT >= 0 && mF + mT <= AVEC_SPEC_MAX_MASKS);
    dim3 grid((unsigned)min(cdiv(F * M, 256), 64), (unsigned)B);
    spec_augment_kernel<<<grid, 256, 0, as_stream(stream)>>>(mel, lengths, B, F, M, mF, Fmax, mT, pS, rng_state, (uint32_t)site,
                                                            intervals);
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}

// ------------------------------------------------------------------------------------------- video augmentation
// The training_video_transform of the reference's LRS2/3 configs (configs/LRS23/AV/EffConfInterCTC.py:82-88), per sample, on the
// device and on the whole padded batch at once:  torchvision RandomCrop(Ho, Wo) -> RandomHorizontalFlip(p) -> nnet.TimeMaskSecond
// (nnet/transforms.py:108-126: int(T_b / fps * num_mask_second) masks over time, each torchaudio mask_along_axis(mask_param =
// int(T_second * fps), mask_value = mean of the CURRENT clip): value = rand * mask_param, min = rand * (T_b - value), frames
// [int(min), int(min) + int(value)) replaced).  Draws: Philox(ctr = (b, k, site, step)): k = 0 -> crop row (word 0), crop column
// (word 1), flip (word 2);  k = 1 + m -> mask m (words 0, 1).  Frames beyond the sample's length are written as zeros (the
// collate padding of nnet/collate_fn.py).
#define AVEC_VIDEO_MAX_MASKS 32
struct VideoDraw { int oy, ox, flip; };
__device__ __forceinline__ VideoDraw video_draw(const unsigned long long* rng, uint32_t site, uint32_t b, int Hi, int Wi, int Ho, int Wo, float flip_p) {
    const unsigned long long seed = rng[0], step = rng[1];
    const uint4 r = philox4x32_10(make_uint4(b, 0u, site, (uint32_t)step), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    VideoDraw d;
    d.oy = min((int)__fmul_rn(u01(r.x), (float)(Hi - Ho + 1)), Hi - Ho);
    d.ox = min((int)__fmul_rn(u01(r.y), (float)(Wi - Wo + 1)), Wi - Wo);
    d.flip = u01(r.z) < flip_p ? 1 : 0;
    return d;
}

// fsum[b][t] = sum over the crop window of frame t (a horizontal flip does not change it); one CTA per (b, t)
__global__ void __launch_bounds__(256) video_frame_sum_kernel(const float* __restrict__ in, const long long* __restrict__ lengths, float* __restrict__ fsum,
                                                              int T, int Hi, int Wi, int Ho, int Wo, float flip_p,
                                                              const unsigned long long* __restrict__ rng, uint32_t site) {
    __shared__ float red[8];
    const int t = blockIdx.x, b = blockIdx.y;
    const int len = lengths ? (int)min((long long)T, lengths[b]) : T;
    float s = 0.0f;
    if (t < len) {
        const VideoDraw d = video_draw(rng, site, (uint32_t)b, Hi, Wi, Ho, Wo, flip_p);
        const float* fr = in + ((size_t)b * T + t) * Hi * Wi;
        for (int i = threadIdx.x; i < Ho * Wo; i += blockDim.x) s += fr[(size_t)(d.oy + i / Wo) * Wi + d.ox + i % Wo];
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.0f;
        for (int w = 0; w < 8; ++w) tot += red[w];
        fsum[(size_t)b * T + t] = tot;
    }
}

__global__ void __launch_bounds__(256) video_augment_kernel(const float* __restrict__ in, const long long* __restrict__ lengths, const float* __restrict__ fsum,
                                                            float* __restrict__ out, int T, int Hi, int Wi, int Ho, int Wo, float flip_p, int mask_T,
                                                            float fps, float num_mask_second, const unsigned long long* __restrict__ rng, uint32_t site,
                                                            int* __restrict__ draws) {
    __shared__ float s_val;
    __shared__ int s_masked;
    const int t = blockIdx.x, b = blockIdx.y;
    const int len = lengths ? (int)min((long long)T, lengths[b]) : T;
    float* dst = out + ((size_t)b * T + t) * Ho * Wo;
    if (t >= len) {
        for (int i = threadIdx.x; i < Ho * Wo; i += blockDim.x) dst[i] = 0.0f;
        return;
    }
    const VideoDraw d = video_draw(rng, site, (uint32_t)b, Hi, Wi, Ho, Wo, flip_p);
    if (threadIdx.x == 0) {
        // the masks are sequential (each fills with the mean of the clip as the previous masks left it): replay them on the
        // per-frame sums; this CTA only needs the value its own frame ends up with
        const int nmask = min((int)((float)len / fps * num_mask_second), AVEC_VIDEO_MAX_MASKS);
        const float hw = (float)(Ho * Wo);
        const float* fs = fsum + (size_t)b * T;
        int lo[AVEC_VIDEO_MAX_MASKS], hi[AVEC_VIDEO_MAX_MASKS];
        float val[AVEC_VIDEO_MAX_MASKS];
        int masked = 0;
        float mine = 0.0f;
        for (int m = 0; m < nmask; ++m) {
            mask_interval(rng, site, (uint32_t)b, (uint32_t)(1 + m), (float)mask_T, (float)len, &lo[m], &hi[m]);
            double tot = 0.0;
            for (int f = 0; f < len; ++f) {
                float v = fs[f];
                for (int q = m - 1; q >= 0; --q) if (f >= lo[q] && f < hi[q]) { v = val[q] * hw; break; }
                tot += (double)v;
            }
            val[m] = (float)(tot / ((double)hw * (double)len));
            if (t >= lo[m] && t < hi[m]) { masked = 1; mine = val[m]; }
            if (draws && t == 0) { draws[((size_t)b * (3 + 2 * AVEC_VIDEO_MAX_MASKS)) + 3 + 2 * m] = lo[m]; draws[((size_t)b * (3 + 2 * AVEC_VIDEO_MAX_MASKS)) + 4 + 2 * m] = hi[m]; }
        }
        if (draws && t == 0) {
            int* dr = draws + (size_t)b * (3 + 2 * AVEC_VIDEO_MAX_MASKS);
            dr[0] = d.oy; dr[1] = d.ox; dr[2] = d.flip;
            for (int m = nmask; m < AVEC_VIDEO_MAX_MASKS; ++m) { dr[3 + 2 * m] = 0; dr[4 + 2 * m] = 0; }
        }
        s_masked = masked; s_val = mine;
    }
    __syncthreads();
    const float* fr = in + ((size_t)b * T + t) * Hi * Wi;
    const bool masked = s_masked != 0;
    const float mv = s_val;
    for (int i = threadIdx.x; i < Ho * Wo; i += blockDim.x) {
        const int y = i / Wo, x = i - y * Wo;
        dst[i] = masked ? mv : fr[(size_t)(d.oy + y) * Wi + d.ox + (d.flip ? Wo - 1 - x : x)];
    }
}

extern "C" int avec_video_augment(const float* in, const long long* lengths, float* out, float* frame_sums, int B, int T, int Hi, int Wi, int Ho,
                                  int Wo, float flip_p, int mask_T, float fps, float num_mask_second, const unsigned long long* rng_state,
                                  int site, int* draws, avec_stream_t stream) {
    AVEC_CHECK_ARG(in && out && frame_sums && rng_state && B > 0 && T > 0 && Ho > 0 && Wo > 0 && Hi >= Ho && Wi >= Wo && fps > 0.0f && mask_T >= 0);
    AVEC_CHECK_ARG(B <= 65535 && (int)((float)T / fps * num_mask_second) <= AVEC_VIDEO_MAX_MASKS);
    dim3 grid((unsigned)T, (unsigned)B);
    video_frame_sum_kernel<<<grid, 256, 0, as_stream(stream)>>>(in, lengths, frame_sums, T, Hi, Wi, Ho, Wo, flip_p, rng_state, (uint32_t)site);
    AVEC_LAUNCH_CHECK();
    video_augment_kernel<<<grid, 256, 0, as_stream(stream)>>>(in, lengths, frame_sums, out, T, Hi, Wi, Ho, Wo, flip_p, mask_T, fps, num_mask_second,
                                                             rng_state, (uint32_t)site, draws);
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}

// ------------------------------------------------------------------------------------------------ greedy CTC decode
// nnet/decoders.py:97-120: argmax over the vocabulary (first maximum, as torch.argmax), frames beyond the utterance
// length dropped, consecutive repeats merged, blanks removed.  One CTA per utterance: warps take frames round-robin,
// thread 0 compacts.  align [B,T] int32 (frame-level argmax, -1 beyond the length), tokens [B,T] int32 (padded with -1),
// ntok [B] int32.
__global__ void __launch_bounds__(128) ctc_greedy_kernel(const float* __restrict__ logits, const long long* __restrict__ in_len,
                                                         int* __restrict__ align, int* __restrict__ tokens, int* __restrict__ ntok,
                                                         int T, int V, int blank) {
    extern __shared__ int pred[];
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int len = in_len ? (int)max(0ll, min((long long)T, in_len[b])) : T;
    for (int t = warp; t < T; t += 4) {
        int best = -1;
        if (t < len) {
            const float* row = logits + ((size_t)b * T + t) * V;
            float bv = -INFINITY;
            best = 0x7FFFFFFF;
            if (lane < V) { bv = row[lane]; best = lane; }
            for (int c = lane + 32; c < V; c += 32) {
                const float v = row[c];
                if (v > bv || (v != v && bv == bv)) { bv = v; best = c; }   // NaN counts as the maximum, as in torch
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, best, o);
                const bool onan = ov != ov, bnan = bv != bv;
                const bool greater = (onan && !bnan) || (!onan && !bnan && ov > bv);
                const bool equal = (onan && bnan) || ov == bv;
                if (greater || (equal && oi < best)) { bv = ov; best = oi; }
            }
        }
        if (lane == 0) { pred[t] = best; if (align) align[(size_t)b * T + t] = best; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int n = 0, prev = -1;
        for (int t = 0; t < len; ++t) {
            const int p = pred[t];
            if (p != prev && p != blank) tokens[(size_t)b * T + n++] = p;
            prev = p;
        }
        ntok[b] = n;
        for (int t = n; t < T; ++t) tokens[(size_t)b * T + t] = -1;
    }
}

extern "C" int avec_ctc_greedy_decode(const float* logits, const long long* in_len, int* align, int* tokens, int* ntok, int B,
                                      int T, int V, int blank, avec_stream_t stream) {
    AVEC_CHECK_ARG(logits && tokens && ntok && B > 0 && T > 0 && V > 0 && (size_t)T * sizeof(int) <= 48 * 1024);
    ctc_greedy_kernel<<<B, 128, (size_t)T * sizeof(int), as_stream(stream)>>>(logits, in_len, align, tokens, ntok, T, V, blank);
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}

// -------------------------------------------------------------------------------------------------------- fused Adam
// One launch over the FLAT fp32 parameter / gradient / moment buffers of the whole model (the same flat gradient buffer
// the NCCL all-reduce uses), replacing ~1100 per-tensor launches of torch.optim.Adam + clip_grad_norm_ + the EMA loop
// (nnet/optimizers.py:61-93, nnet/schedulers.py:120-137, nnet/model.py:378-407):
//   g' = clip * g + wd * p;  m = b1 m + (1-b1) g';  v = b2 v + (1-b2) g'^2;
//   p -= lr_t / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps);  ema = tau * ema + (1 - tau) * p
// t = *step (device, already advanced by avec_counter_advance), lr_t = lr_a (constant) or the Noam schedule
// lr_a * min(t * lr_b^-1.5, t^-0.5)  (lr_a = val_factor * dim_decay^-0.5, lr_b = warmup steps);
// clip = min(1, max_norm / (sqrt(*sumsq) + 1e-6)) when sumsq != NULL (torch.nn.utils.clip_grad_norm_).
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ out) {
    float acc = 0.0f;
    const long long n4 = n >> 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = reinterpret_cast<const float4*>(g)[i];
        acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) { const float v = g[(n4 << 2) + threadIdx.x]; acc += v * v; }
    acc = warp_sum(acc);
    __shared__ float part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 8) {
        float v = part[threadIdx.x];
        v += __shfl_xor_sync(0xffu, v, 4); v += __shfl_xor_sync(0xffu, v, 2); v += __shfl_xor_sync(0xffu, v, 1);
        if (threadIdx.x == 0) atomicAdd(out, v);
    }
}

extern "C" int avec_sumsq(const float* g, long long n, float* out, avec_stream_t stream) {
    AVEC_CHECK_ARG(g && out && n > 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0);
    const int blocks = (int)min(cdivll(n >> 2, 256 * 8) + 1, (long long)148 * 8);
    sumsq_kernel<<<blocks, 256, 0, as_stream(stream)>>>(g, n, out);
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}

struct AdamHyper { float b1, b2, eps, wd, lr_a, lr_b, max_norm, ema_tau; int lr_mode; };

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, float* __restrict__ ema, long long n, AdamHyper h,
                                                   const unsigned long long* __restrict__ step, const float* __restrict__ sumsq,
                                                   float* __restrict__ lr_out) {
    const float t = (float)step[0];
    float lr = h.lr_a;
    if (h.lr_mode == 1) lr = h.lr_a * fminf(t * rsqrtf(h.lr_b) / h.lr_b, rsqrtf(t));
    const float bc1 = 1.0f - powf(h.b1, t), bc2s = sqrtf(1.0f - powf(h.b2, t));
    const float step_size = lr / bc1;
    float clip = 1.0f;
    if (sumsq) clip = fminf(1.0f, h.max_norm / (sqrtf(sumsq[0]) + 1e-6f));
    if (lr_out && blockIdx.x == 0 && threadIdx.x == 0) { lr_out[0] = lr; lr_out[1] = sumsq ? sqrtf(sumsq[0]) : 0.0f; }
    const long long n4 = n >> 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 P = reinterpret_cast<float4*>(p)[i], G = reinterpret_cast<const float4*>(g)[i];
        float4 Mv = reinterpret_cast<float4*>(m)[i], Vv = reinterpret_cast<float4*>(v)[i];
        float* pp = &P.x; float* gg = &G.x; float* mm = &Mv.x; float* vv = &Vv.x;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float gj = clip * gg[j] + h.wd * pp[j];
            mm[j] = h.b1 * mm[j] + (1.0f - h.b1) * gj;
            vv[j] = h.b2 * vv[j] + (1.0f - h.b2) * gj * gj;
            pp[j] -= step_size * mm[j] / (sqrtf(vv[j]) / bc2s + h.eps);
        }
        reinterpret_cast<float4*>(p)[i] = P;
        reinterpret_cast<float4*>(m)[i] = Mv;
        reinterpret_cast<float4*>(v)[i] = Vv;
        if (ema) {
            float4 E = reinterpret_cast<float4*>(ema)[i];
            E.x = h.ema_tau * E.x + (1.0f - h.ema_tau) * P.x; E.y = h.ema_tau * E.y + (1.0f - h.ema_tau) * P.y;
            E.z = h.ema_tau * E.z + (1.0f - h.ema_tau) * P.z; E.w = h.ema_tau * E.w + (1.0f - h.ema_tau) * P.w;
            reinterpret_cast<float4*>(ema)[i] = E;
        }
    }
}

extern "C" int avec_adam_step(float* p, const float* g, float* m, float* v, float* ema, long long n, float beta1, float beta2,
                              float eps, float weight_decay, int lr_mode, float lr_a, float lr_b, float max_norm, float ema_tau,
                              const unsigned long long* step, const float* sumsq, float* lr_out, avec_stream_t stream) {
    AVEC_CHECK_ARG(p && g && m && v && step && n > 0 && (n & 3) == 0 && (lr_mode == 0 || lr_mode == 1));
    AVEC_CHECK_ARG(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(ema)) & 15) == 0);
    AdamHyper h{beta1, beta2, eps, weight_decay, lr_a, lr_b, max_norm, ema_tau, lr_mode};
    const int blocks = (int)min(cdivll(n >> 2, 256 * 4), (long long)148 * 8);
    adam_kernel<<<blocks, 256, 0, as_stream(stream)>>>(p, g, m, v, ema, n, h, step, sumsq, lr_out);
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}
