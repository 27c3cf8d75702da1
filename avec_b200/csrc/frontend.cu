// Audio front-end: framing + Hann window + 512-point FFT + power spectrum + mel filterbank + log, one kernel.
// Follows torchaudio.transforms.Spectrogram(n_fft=512, win_length=400, hop_length=160, center=True, pad_mode="reflect",
// power=2, hann periodic, window zero-padded to n_fft on both sides) -> MelScale(80 mels, htk, norm=None) ->
// log(x + 1e-9), as called by AudioPreprocessing.forward (reference nnet/preprocessing.py:57-85).
// cuFFT-free: one warp = one frame, radix-2 DIT in shared memory; HBM traffic = 4 B/sample in, 320 B/frame out.
#include "common.cuh"
#include <algorithm>

namespace {

constexpr int NFFT = 512, WIN = 400, HOP = 160, NBIN = 257, NMEL = 80;
constexpr int WOFF = (NFFT - WIN) / 2;  // 56
constexpr int FR_PER_CTA = 4;

__device__ __forceinline__ int bitrev9(int x) { return (int)(__brev((unsigned)x) >> 23); }

__global__ void __launch_bounds__(FR_PER_CTA * 32) stft_mel_log_kernel(const float* __restrict__ wave, const float* __restrict__ fb,
                                                                     float* __restrict__ out, int L, int F, int layout) {
    __shared__ float2 tw[NFFT / 2];
    __shared__ float win[WIN];
    __shared__ float2 buf[FR_PER_CTA][NFFT];
    __shared__ float pw[FR_PER_CTA][NBIN + 3];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y;
    for (int k = tid; k < NFFT / 2; k += blockDim.x) {
        float s, c;
        sincospif(-2.0f * (float)k / (float)NFFT, &s, &c);
        tw[k] = make_float2(c, s);
    }
    for (int n = tid; n < WIN; n += blockDim.x) win[n] = 0.5f - 0.5f * cospif(2.0f * (float)n / (float)WIN);
    __syncthreads();
    const int f = blockIdx.x * FR_PER_CTA + warp;
    if (f >= F) return;  // no block-level barrier below
    const float* x = wave + (size_t)b * L;
    float2* xb = buf[warp];
    // load (reflect padding of 256 on both sides), window, bit-reversed placement
    for (int n = lane; n < NFFT; n += 32) {
        float v = 0.0f;
        if (n >= WOFF && n < WOFF + WIN) {
            int idx = f * HOP + n - NFFT / 2;
            if (idx < 0) idx = -idx;
            if (idx >= L) idx = 2 * (L - 1) - idx;
            idx = max(0, min(L - 1, idx));
            v = x[idx] * win[n - WOFF];
        }
        xb[bitrev9(n)] = make_float2(v, 0.0f);
    }
    __syncwarp();
#pragma unroll 1
    for (int len = 2; len <= NFFT; len <<= 1) {
        const int half = len >> 1, step = NFFT / len;
        for (int j = lane; j < NFFT / 2; j += 32) {
            int grp = j / half, pos = j % half;
            int i0 = grp * len + pos, i1 = i0 + half;
            float2 w = tw[pos * step];
            float2 a = xb[i0], c = xb[i1];
            float2 t = make_float2(w.x * c.x - w.y * c.y, w.x * c.y + w.y * c.x);
            xb[i0] = make_float2(a.x + t.x, a.y + t.y);
            xb[i1] = make_float2(a.x - t.x, a.y - t.y);
        }
        __syncwarp();
    }
    for (int k = lane; k < NBIN; k += 32) { float2 v = xb[k]; pw[warp][k] = v.x * v.x + v.y * v.y; }
    __syncwarp();
    for (int m = lane; m < NMEL; m += 32) {
        float acc = 0.0f;
        for (int k = 0; k < NBIN; ++k) acc = fmaf(pw[warp][k], __ldg(fb + k * NMEL + m), acc);
        float v = logf(acc + 1e-9f);
        if (layout == 0) out[((size_t)b * F + f) * NMEL + m] = v;
        else out[((size_t)b * NMEL + m) * F + f] = v;
    }
}

// single-channel im2col: col[site][tap] = x[n, to*st+kt-pt, ho*sh+kh-ph, wo*sw+kw-pw] (0 outside), taps padded with zeros to
// Kpad.  One CTA = one (n, to) slice x IM_BH output rows x all Wo columns: the input window (KT x IH x IW elements, zero
// padding applied once) is staged in shared memory with coalesced loads, then every warp emits whole 2*Kpad-byte rows
// (lane = 8 consecutive taps = one 16-byte store).  HBM-bound on the col write (6.4 GB at B = 64).
constexpr int IM_BH = 4;

template <typename T>
__global__ void __launch_bounds__(256) im2col_c1_kernel(const T* __restrict__ x, T* __restrict__ col, ConvGeom g, int taps, int Kpad,
                                                        int IH, int IW) {
    extern __shared__ float tile[];    // [KT][IH][IW]
    __shared__ int tapofs[256];        // tap -> offset inside the tile
    const int hob = blockIdx.y * IM_BH;
    const int to = blockIdx.x % g.To, n = blockIdx.x / g.To;
    for (int tap = threadIdx.x; tap < 256; tap += blockDim.x) {
        int t = tap;
        const int kw = t % g.KW; t /= g.KW; const int kh = t % g.KH; const int kt = t / g.KH;
        tapofs[tap] = tap < taps ? (kt * IH + kh) * IW + kw : -1;
    }
    const int t0 = to * g.st - g.pt, h0 = hob * g.sh - g.ph, w0 = -g.pw;
    const int n_el = g.KT * IH * IW;
    for (int i = threadIdx.x; i < n_el; i += blockDim.x) {
        const int iw = i % IW; int r = i / IW; const int ih = r % IH; const int kt = r / IH;
        const int ti = t0 + kt, hi = h0 + ih, wi = w0 + iw;
        float v = 0.0f;
        if ((unsigned)ti < (unsigned)g.Ti && (unsigned)hi < (unsigned)g.Hi && (unsigned)wi < (unsigned)g.Wi)
            v = ldf(x + (((long long)n * g.Ti + ti) * g.Hi + hi) * g.Wi + wi);
        tile[i] = v;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int rows = min(IM_BH, g.Ho - hob);
    for (int sidx = warp; sidx < rows * g.Wo; sidx += nw) {
        const int hl = sidx / g.Wo, wo = sidx % g.Wo;
        const long long site = (((long long)n * g.To + to) * g.Ho + hob + hl) * g.Wo + wo;
        const int base = hl * g.sh * IW + wo * g.sw;
        for (int c0 = lane * 8; c0 < Kpad; c0 += 256) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int tap = c0 + j;
                const int o = tap < 256 ? tapofs[tap] : -1;
                v[j] = o >= 0 ? tile[base + o] : 0.0f;
            }
            store_vec<8>(col + site * Kpad + c0, v);
        }
    }
}


// ---- audio stem: Conv2d(1 -> Co, 3x3, stride 2, pad 1) on the log-mel image (reference nnet/networks.py:359-368).  K = 9 with a
// single input channel is far below a tensor-core tile and the op is bound by the [sites, Co] output write (185 MB at B = 64),
// so it is a SIMT kernel: a thread owns 4 output channels (its 36 weights live in registers) and walks output sites; the
// 9 taps of a site are broadcast loads, the 45 threads of a site store one contiguous 360-byte row.  BatchNorm column sums
// (forward) and the weight gradient (backward: same walk, dY instead of the weights) are reduced per CTA, then by atomics.
constexpr int S2_ROWS = 4;      // sites processed concurrently per CTA
constexpr int S2_GROUPS = 64;   // channel groups of 4 per site row (Co <= 256)

template <typename T>
__device__ __forceinline__ void stem2d_taps(const T* __restrict__ x, int n, int ho, int wo, int H, int W, float (&t)[9]) {
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
            const int hi = 2 * ho + kh - 1, wi = 2 * wo + kw - 1;
            t[kh * 3 + kw] = ((unsigned)hi < (unsigned)H && (unsigned)wi < (unsigned)W) ? ldf(x + ((size_t)n * H + hi) * W + wi) : 0.0f;
        }
}

template <typename T>
__global__ void __launch_bounds__(S2_ROWS * S2_GROUPS) stem2d_fwd_kernel(const T* __restrict__ x, const T* __restrict__ w, const float* __restrict__ bias,
                                                                        T* __restrict__ out, float* __restrict__ stats, int N, int H, int W, int Ho,
                                                                        int Wo, int Co, long long sites, long long sites_per_cta) {
    __shared__ float red[2][S2_ROWS][S2_GROUPS * 4];
    const int grp = threadIdx.x % S2_GROUPS, row = threadIdx.x / S2_GROUPS;
    const int c0 = grp * 4;
    const bool cv = c0 < Co;
    float wr[4][9], b4[4], s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        b4[j] = (cv && bias) ? bias[c0 + j] : 0.0f;
#pragma unroll
        for (int k = 0; k < 9; ++k) wr[j][k] = cv ? ldf(w + (size_t)(c0 + j) * 9 + k) : 0.0f;
    }
    const long long sbeg = (long long)blockIdx.x * sites_per_cta, send = min(sites, sbeg + sites_per_cta);
    for (long long s = sbeg + row; s < send; s += S2_ROWS) {
        const int wo = (int)(s % Wo); const long long r = s / Wo; const int ho = (int)(r % Ho), n = (int)(r / Ho);
        float t[9];
        stem2d_taps(x, n, ho, wo, H, W, t);
        if (cv) {
            float y[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float a = b4[j];
#pragma unroll
                for (int k = 0; k < 9; ++k) a = fmaf(wr[j][k], t[k], a);
                y[j] = a; s1[j] += a; s2[j] += a * a;
            }
            store_vec<4>(out + (size_t)s * Co + c0, y);
        }
    }
    if (stats) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { red[0][row][grp * 4 + j] = s1[j]; red[1][row][grp * 4 + j] = s2[j]; }
        __syncthreads();
        float* dst = stats + (size_t)(blockIdx.x % AVEC_STATS_REPLICAS) * 2 * Co;
        for (int i = threadIdx.x; i < 2 * S2_GROUPS * 4; i += blockDim.x) {
            const int q = i / (S2_GROUPS * 4), c = i % (S2_GROUPS * 4);
            if (c < Co) {
                float v = 0.f;
#pragma unroll
                for (int rr = 0; rr < S2_ROWS; ++rr) v += red[q][rr][c];
                atomicAdd(dst + q * Co + c, v);
            }
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(S2_ROWS * S2_GROUPS) stem2d_wgrad_kernel(const T* __restrict__ x, const T* __restrict__ dy, float* __restrict__ dw, int N, int H,
                                                                          int W, int Ho, int Wo, int Co, long long sites, long long sites_per_cta) {
    __shared__ float red[S2_ROWS][S2_GROUPS * 4][9];
    const int grp = threadIdx.x % S2_GROUPS, row = threadIdx.x / S2_GROUPS;
    const int c0 = grp * 4;
    const bool cv = c0 < Co;
    float acc[4][9];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int k = 0; k < 9; ++k) acc[j][k] = 0.0f;
    const long long sbeg = (long long)blockIdx.x * sites_per_cta, send = min(sites, sbeg + sites_per_cta);
    for (long long s = sbeg + row; s < send; s += S2_ROWS) {
        const int wo = (int)(s % Wo); const long long r = s / Wo; const int ho = (int)(r % Ho), n = (int)(r / Ho);
        float t[9];
        stem2d_taps(x, n, ho, wo, H, W, t);
        if (cv) {
            float d[4];
            load_vec<4>(dy + (size_t)s * Co + c0, d);
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int k = 0; k < 9; ++k) acc[j][k] = fmaf(d[j], t[k], acc[j][k]);
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int k = 0; k < 9; ++k) red[row][grp * 4 + j][k] = acc[j][k];
    __syncthreads();
    for (int i = threadIdx.x; i < Co * 9; i += blockDim.x) {
        const int c = i / 9, k = i % 9;
        float v = 0.f;
#pragma unroll
        for (int rr = 0; rr < S2_ROWS; ++rr) v += red[rr][c][k];
        atomicAdd(dw + (size_t)c * 9 + k, v);
    }
}

}  // namespace

extern "C" int avec_im2col_c1(const void* x, void* col, const avec_conv_geom* geom, int Kpad, int dtype, avec_stream_t stream) {
    AVEC_CHECK_ARG(x && col && geom && geom->C == 1 && Kpad % 8 == 0);
    ConvGeom g = make_geom(*geom);
    const int taps = g.KT * g.KH * g.KW;
    AVEC_CHECK_ARG(taps <= 256 && taps <= Kpad);
    const int IH = (IM_BH - 1) * g.sh + g.KH, IW = (g.Wo - 1) * g.sw + g.KW;
    const size_t smem = (size_t)g.KT * IH * IW * sizeof(float);
    if (smem > 200 * 1024 || cdiv(g.Ho, IM_BH) > 65535) return AVEC_ERR_UNSUPPORTED;
    dim3 grid(g.N * g.To, cdiv(g.Ho, IM_BH));
    AVEC_DISPATCH_DTYPE(dtype, Tt, {
        auto kfn = im2col_c1_kernel<Tt>;
        if (smem > 48 * 1024 && cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return AVEC_ERR_LAUNCH;
        kfn<<<grid, 256, smem, as_stream(stream)>>>((const Tt*)x, (Tt*)col, g, taps, Kpad, IH, IW);
    });
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}

static bool stem2d_args_ok(const void* x, const void* o, int N, int H, int W, int Co) {
    return x && o && N > 0 && H > 0 && W > 0 && Co > 0 && Co % 4 == 0 && Co <= S2_GROUPS * 4;
}

extern "C" int avec_stem2d_fwd(const void* x, const void* w, const float* bias, void* out, float* colstats, int N, int H, int W, int Co, int dtype,
                               avec_stream_t stream) {
    AVEC_CHECK_ARG(stem2d_args_ok(x, out, N, H, W, Co) && w);
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    const long long sites = (long long)N * Ho * Wo;
    const int ctas = (int)std::min<long long>(cdivll(sites, 64), 148 * 8);
    const long long spc = cdivll(sites, ctas);
    AVEC_DISPATCH_DTYPE(dtype, Tt, {
        stem2d_fwd_kernel<Tt><<<(unsigned)cdivll(sites, spc), S2_ROWS * S2_GROUPS, 0, as_stream(stream)>>>((const Tt*)x, (const Tt*)w, bias, (Tt*)out, colstats, N, H, W, Ho,
                                                                                                         Wo, Co, sites, spc);
    });
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}

extern "C" int avec_stem2d_wgrad(const void* x, const void* dy, float* dw, int N, int H, int W, int Co, int dtype, avec_stream_t stream) {
    AVEC_CHECK_ARG(stem2d_args_ok(x, dy, N, H, W, Co) && dw);
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    const long long sites = (long long)N * Ho * Wo;
    const int ctas = (int)std::min<long long>(cdivll(sites, 64), 148 * 4);
    const long long spc = cdivll(sites, ctas);
    AVEC_DISPATCH_DTYPE(dtype, Tt, {
        stem2d_wgrad_kernel<Tt><<<(unsigned)cdivll(sites, spc), S2_ROWS * S2_GROUPS, 0, as_stream(stream)>>>((const Tt*)x, (const Tt*)dy, dw, N, H, W, Ho, Wo, Co, sites, spc);
    });
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}

extern "C" int avec_stft_mel_log(const float* wave, const float* fb, float* out, int B, int L, int F, int layout,
                                 avec_stream_t stream) {
    AVEC_CHECK_ARG(wave && fb && out && B > 0 && B <= 65535 && L > NFFT / 2 && F == L / HOP + 1 && (layout == 0 || layout == 1));
    dim3 grid(cdiv(F, FR_PER_CTA), B);
    stft_mel_log_kernel<<<grid, FR_PER_CTA * 32, 0, as_stream(stream)>>>(wave, fb, out, L, F, layout);
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}
