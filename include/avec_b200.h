/*
 * avec_b200.h — C ABI of the B200-native (sm_100a) hot path of AVEC's Efficient-Conformer encoder.
 *
 * The reference (burchim/AVEC) is pure Python on PyTorch: every entry point below replaces a *sequence of ATen calls*
 * issued by a reference nn.Module.forward (and its autograd backward), cited per function as nnet/<file>.py:<lines>.
 * Conventions:
 *   - device pointers only, caller owns every buffer (no allocation, no host sync, no global state besides a
 *     thread-local "last CUDA error" for diagnostics); all launches go to the caller's stream;
 *   - activations are channels-last, row-major: tokens [rows, C]; images [N, H, W, C]; videos [N, T, H, W, C];
 *   - `dtype` arguments select the activation element type (AVEC_F32 parity mode / AVEC_BF16 production mode);
 *     parameters of normalisations, biases, statistics and all gradients of parameters are fp32;
 *   - return 0 on success, negative avec_status otherwise (avec_strerror()).
 */
#ifndef AVEC_B200_H
#define AVEC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* avec_stream_t; /* cudaStream_t */

enum avec_status {
    AVEC_OK = 0,
    AVEC_ERR_INVALID = -1,     /* bad argument / unsupported shape */
    AVEC_ERR_LAUNCH = -2,      /* cudaGetLastError() != success after a launch (see avec_last_cuda_error) */
    AVEC_ERR_UNSUPPORTED = -3, /* combination not implemented by this build */
    AVEC_ERR_DRIVER = -4       /* driver entry point (tensor-map encode) unavailable or failed */
};

enum avec_dtype { AVEC_F32 = 0, AVEC_BF16 = 1 };

const char* avec_strerror(int status);
int avec_last_cuda_error(void);
int avec_version(void);
/* number of kernels launched by this library on the calling thread since the last reset (bench.py's gpu_launches) */
long long avec_launch_count(void);
void avec_reset_launch_count(void);

/* ------------------------------------------------------------------------------------------------------------------
 * GEMM / implicit-GEMM convolution:   out = epilogue( sum_k A(m,k) * B(n,k) )
 * Replaces F.linear / addmm (nnet/layers.py:29-76), the k=1 Conv1d of the ConvModule (nnet/modules.py:372-381),
 * conv2d/conv3d of the front-ends (nnet/networks.py:359-368, 459-471; nnet/blocks.py:64-91) and all their autograd
 * backward GEMMs (dgrad / wgrad).
 * ------------------------------------------------------------------------------------------------------------------ */
enum avec_gemm_mode {
    AVEC_GEMM_PLAIN = 0,      /* A(m,k) = A[m*sam + k*sak],  B(n,k) = B[n*sbn + k*sbk]                                  */
    AVEC_GEMM_CONV_FWD = 1,   /* A = im2col(X): m = output site, k = (tap, ci);  B(n=co, k) = W[co][tap][ci]            */
    AVEC_GEMM_CONV_DGRAD = 2, /* A = col2im gather of dY: m = input site, k = (tap, co);  B(n=ci,k) = Wd[ci][tap][co]    */
    AVEC_GEMM_CONV_WGRAD = 3  /* A(m=co,k=site) = dY[site][co];  B(n=(tap,ci), k=site) = X[shift(site,tap)][ci]          */
};

enum avec_epilogue {
    AVEC_EPI_LINEAR = 0,   /* out = alpha * (acc + bias)                                                              */
    AVEC_EPI_SWISH = 1,    /* pre = acc + bias; out2 = pre (if given); out = pre * sigmoid(pre)                       */
    AVEC_EPI_RESIDUAL = 2, /* out = aux + alpha * (acc + bias)                                                        */
    AVEC_EPI_DSWISH = 3,   /* out = alpha * acc * swish'(aux)            (backward through Swish, aux = saved pre)     */
    AVEC_EPI_ACCUM = 4,    /* out(fp32) += alpha * acc   (atomic; split-K partial sums of weight gradients)            */
    AVEC_EPI_RELU = 5      /* out = max(0, alpha * (acc + bias) + aux?)                                               */
};

enum avec_gemm_impl { AVEC_IMPL_AUTO = 0, AVEC_IMPL_SIMT = 1, AVEC_IMPL_TCGEN05 = 2 };

/* N-d convolution geometry (2-d and 1-d convolutions set the unused extents to 1). X is [N, Ti, Hi, Wi, C],
 * Y is [N, To, Ho, Wo, Co], W is [Co][KT][KH][KW][C] (tap-major, channel-minor). */
typedef struct avec_conv_geom {
    int N, Ti, Hi, Wi, C;
    int To, Ho, Wo, Co;
    int KT, KH, KW;
    int st, sh, sw; /* strides */
    int pt, ph, pw; /* leading zero padding ("same": (k-1)/2, nnet/layers.py:250-258) */
} avec_conv_geom;

typedef struct avec_gemm_args {
    int mode; /* avec_gemm_mode */
    int impl; /* avec_gemm_impl */
    int M, N, K;
    const void* A;
    long long sam, sak; /* element strides (PLAIN) */
    const void* B;
    long long sbn, sbk;
    int ab_dtype; /* element type of A and B */
    avec_conv_geom g;
    int epi;
    float alpha;
    const float* bias; /* [N] or NULL */
    void* out;
    int out_dtype;
    long long ldo;
    void* out2;
    int out2_dtype;
    long long ldo2;
    const void* aux;
    int aux_dtype;
    long long ldaux;
    float* colstats; /* optional [32][2*N] (32 replicas selected by CTA index, summed by avec_bn_finalize):
                        += sum_m v, += sum_m v^2 of v = acc + bias (BatchNorm batch statistics) */
    int split_k;     /* ACCUM epilogue: number of K slices (0/1 = none) */
    /* nn.Dropout fused behind the GEMM (drop_p > 0; tcgen05 path with bf16 output, else AVEC_ERR_UNSUPPORTED and the caller runs
     * avec_dropout): the epilogue's value BEFORE the residual add is multiplied by keep / (1 - p) with the mask avec_dropout draws
     * for (rng_state, site) - LINEAR alpha*drop(acc+bias), SWISH drop(swish(.)), RESIDUAL aux + alpha*drop(acc+bias),
     * DSWISH drop(alpha*acc*swish'(aux)) */
    float drop_p;
    int drop_site;
    const unsigned long long* drop_rng;
} avec_gemm_args;

int avec_gemm(const avec_gemm_args* args, avec_stream_t stream);
/* diagnostics: 0 forces the cp.async gather producers even where a TMA descriptor is possible (default 1) */
void avec_set_tma(int enabled);
/* Programmatic dependent launch of the tcgen05 GEMM kernel (default on; AVEC_PDL=0 in the environment disables it for the process):
 * the grid may be scheduled while the preceding kernel of its stream drains and overlaps its prologue (barriers, TMEM allocation,
 * descriptor prefetch) with that tail.  avec_pdl_exclude_stream: launches on this stream keep the ordinary full serialisation - for a
 * secondary stream whose parked CTAs would take shared memory from the critical-path stream (<= 8 streams; enabled = 0 clears the list) */
void avec_set_pdl(int enabled);
void avec_pdl_exclude_stream(avec_stream_t stream, int enabled);
/* diagnostics: device buffer of 232 uint64; CTA (0,0,0) of every tcgen05 GEMM launch records %globaltimer (ns) at
 * [0] start, [1] setup done, [2] first k-block in smem, [3] last MMA issued, [4] accumulator complete, [5] epilogue
 * done, [6] TMEM released; [8 + 5 j + k], j < 32: per-tile timeline of that CTA (k = 0 epilogue starts waiting,
 * 1 accumulator ready, 2 epilogue done, 3 first k-block in smem, 4 last MMA issued).  NULL disables. */
void avec_set_debug_timestamps(void* dev_buf_232_u64);

/* out[n] (+)= alpha * sum_m x[m][n]      — bias gradients (autograd of the bias add in addmm / conv) */
int avec_colsum(const void* x, int dtype, long long rows, int C, long long ldx, float alpha, float* out, int accumulate,
                avec_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * LayerNorm (eps 1e-6) with the patch-attention average pooling folded in.
 * fwd: y[b,t'] = 1/P * sum_{p<P, t'P+p<T} LN(x[b,t'P+p])     (P = 1: plain LayerNorm)
 * Replaces nn.LayerNorm (nnet/modules.py:278,302,373; nnet/blocks.py:267,304) + MultiHeadAttention.pad +
 * AvgPool1d of RelPosPatch1dMultiHeadAttention.forwardQKV (nnet/attentions.py:351-371).
 * bwd: dx = LNbwd(expand(dy)/P) + (dres ? gather(dres, res_stride) : 0);  dgamma/dbeta accumulate (atomic) into fp32.
 * gamma == NULL: identity (no normalisation: plain patch mean / its backward).  ldy (here and in avec_pool_sum, avec_bn_apply,
 * avec_dropout): row pitch of the OUTPUT in elements, 0 = dense; the tensors that feed a GEMM as an operand are written with a
 * 16-byte-aligned pitch (D = 180: 192) so that the tcgen05 kernel can fetch them by TMA instead of its gather producers.
 * ------------------------------------------------------------------------------------------------------------------ */
int avec_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd, int B,
                       int T, int C, int P, float eps, int dtype, long long ldy, avec_stream_t stream);
int avec_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                       const void* dres, int res_stride, void* dx, float* dgamma, float* dbeta, int B, int T, int C,
                       int P, int dtype, avec_stream_t stream);

/* y[b,t] = x[b,t] + o[b, t / P]      (nearest upsample + slice + residual add, nnet/attentions.py:377-380, blocks.py:295) */
int avec_upsample_add(const void* x, const void* o, void* y, int B, int T, int Tp, int C, int P, int dtype,
                      avec_stream_t stream);
/* do[b,t'] = sum_{p<P, t'P+p<T} dy[b, t'P+p]   (its backward) */
int avec_pool_sum(const void* dy, void* dout, int B, int T, int Tp, int C, int P, int dtype, long long ld_out, avec_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Relative-position multi-head self-attention core (nnet/attentions.py:280-323 incl. rel_to_abs 258-276; grouped /
 * Transformer-XL variant 579-650):
 *   S[i,j] = ((q_i+u) . k_j + (q_i+v) . e_{T-1+j-i}) / sqrt(d);  S += -1e9 where masked;  P = softmax_j(S);  o_i = sum_j P v_j
 * qkv is [B*Tf, 3*D1] per FRAME (q | k | v, D1 = H*d/G wide each); a token is G consecutive frames concatenated
 * (G = 1: token = frame), T = ceil(Tf / G) tokens, frames >= Tf are zero rows.  e is [2T-1, G*D1] (pos_layer(R) regrouped,
 * computed once, not B times); u, v: optional fp32 [D1] content / position biases (grouped attention), tiled over the group.
 * klen[b] = number of unmasked key tokens of item b, qlen = number of query rows that are not fully masked (rows >= qlen
 * see every key masked, as happens for the zero-padded last patch, nnet/attentions.py:140-171,355-363).
 * o is [B*Tf, D1]; probs [B,H,T,T] fp32 is saved for the backward.
 * bwd: dqkv [B*Tf, 3*D1]; de [2T-1, G*D1], du, dv [D1] fp32, accumulated atomically; ds_ws: [B,H,T,T] fp32 scratch (dS).
 * ------------------------------------------------------------------------------------------------------------------ */
/* Kernel selection is automatic: tensor-core kernels (bf16, T <= 128 / 112), whole-head SIMT kernels (T <= 416 keys and tiles <=
 * 227 KB), key-tiled SIMT kernels beyond that (T <= ~1500).  avec_set_attention_long(1) forces the key-tiled kernels (tests). */
void avec_set_attention_long(int force);
int avec_relpos_attn_fwd(const void* qkv, const void* e, const int* klen, int qlen, void* o, float* probs, int B, int T,
                         int H, int d, int G, int Tf, const float* u, const float* v, int dtype, avec_stream_t stream);
int avec_relpos_attn_bwd(const void* d_o, const void* qkv, const void* e, const float* probs, float* ds_ws, void* dqkv,
                         float* de, int B, int T, int H, int d, int G, int Tf, const float* u, const float* v, float* du,
                         float* dv, int dtype, avec_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * The same attention as a Blackwell tile kernel (csrc/attention_tc.cu): tcgen05.mma with score / band / output accumulators in
 * TMEM, operands by TMA, flash-style for ANY sequence length (no [B,H,T,T] tensor: the forward saves lse [B,H,T] fp32, the
 * backward recomputes the probabilities).  bf16, G = 1 (regular and - on pooled tokens - patch attention).
 * "Padded heads" layout: every head owns a dp = 64, 128, 192 or 256 wide column block (dp >= d; pad columns are zeros, produced by
 * zero-padded projection weights): qkv [B*T, >= 3*H*dp] with head h of part s (0 q, 1 k, 2 v) at columns (s*H + h)*dp,
 * e [2T-1, >= H*dp], o / d_o [B*T, >= H*dp]; ld_* are row pitches in elements (multiples of 8), bases 16-byte aligned.
 * bwd: dqkv (bf16, qkv's layout) is written directly when T <= 128; for longer sequences the partial sums of the 128 x 128
 * tiles are accumulated atomically into dqkv_ws (fp32, same shape and pitch, zero-initialised by the caller) instead.
 * de [2T-1, >= H*dp] fp32 is accumulated atomically (zero-initialised by the caller).
 * ------------------------------------------------------------------------------------------------------------------ */
int avec_relpos_attn_tc_fwd(const void* qkv, long long ld_qkv, const void* e, long long ld_e, const int* klen, int qlen, void* o,
                            long long ld_o, float* lse, int B, int T, int H, int d, int dp, int qp_part, avec_stream_t stream);
int avec_relpos_attn_tc_bwd(const void* d_o, long long ld_do, const void* qkv, long long ld_qkv, const void* e, long long ld_e,
                            const void* o, long long ld_o, const float* lse, const int* klen, int qlen, void* dqkv, long long ld_dqkv,
                            float* dqkv_ws, float* de, long long ld_de, int B, int T, int H, int d, int dp, int qp_part, avec_stream_t stream);
/* Grouped / Transformer-XL attention (GroupedRelPosMultiHeadSelfAttention, nnet/attentions.py:579-650) on the same tile kernel:
 * qp_part = 3 selects a 4-part layout [q + u | k | v | q + v] (scores use q + u against the keys, q + v against the position rows;
 * dqkv then carries d(q+u) in part 0 and d(q+v) in part 3); dp up to 256 (d = G*D1/H = 135 -> 192).  The projections run at FRAME
 * rate ([B*Tf, 3*D1], e [G*(2Tn-1), D1]); a token is G consecutive frames concatenated and split into H heads.
 * avec_attn_group_pack: frames -> padded-heads tokens (dst_parts 4: qkv with u / v added after the zero padding; 1: plain regroup of a
 * one-part matrix such as e or d_o);  avec_attn_group_unpack: tokens -> frames for a one-part matrix (o, fp32 de);
 * avec_attn_group_unpack_dqkv: dqkv tokens -> [B*Tf, 3*D1] (dq = d(q+u) + d(q+v)) and du, dv [D1] (fp32, atomic). */
int avec_attn_group_pack(const void* src, int src_dtype, long long lds, const float* u, const float* v, void* dst, long long ldd, int B,
                         int Tf, int Tn, int G, int H, int D1, int dp, int dst_parts, avec_stream_t stream);
int avec_attn_group_unpack(const void* src, int src_dtype, long long lds, void* dst, int dst_dtype, long long ldd, int B, int Tf, int Tn, int G,
                           int H, int D1, int dp, avec_stream_t stream);
int avec_attn_group_unpack_dqkv(const void* src, long long lds, void* dst, long long ldd, float* du, float* dv, int B, int Tf, int Tn, int G,
                                int H, int D1, int dp, avec_stream_t stream);

/* row softmax / its backward (InterCTCResModule, nnet/modules.py:397-398).  dadd (optional, fp32) is added to dx. */
int avec_softmax_fwd(const void* x, int x_dtype, void* y, int y_dtype, long long rows, int C, avec_stream_t stream);
int avec_softmax_bwd(const void* dy, const void* y, int dtype, const float* dadd, void* dx, int dx_dtype, long long rows,
                     int C, avec_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * ConvolutionModule middle section (nnet/modules.py:374-379): GLU -> depthwise Conv1d(k, stride s, "same") [-> BN stats]
 *   pre [B,T,2C] (value | gate) -> u [B,To,C];  stats[0:C] += sum u, stats[C:2C] += sum u^2   (stats may be NULL)
 * bwd: given du [B,To,C]: dpre [B,T,2C], dw [C,k] and db [C] (fp32, atomic accumulate).
 * ------------------------------------------------------------------------------------------------------------------ */
int avec_glu_dwconv_fwd(const void* pre, const float* w, const float* bias, void* u, float* stats, int B, int T, int To,
                        int C, int ksize, int stride, int pad, int dtype, avec_stream_t stream);
int avec_glu_dwconv_bwd(const void* du, const void* pre, const float* w, void* dpre, float* dw, float* db, int B, int T,
                        int To, int C, int ksize, int stride, int pad, int dtype, avec_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * BatchNorm over channels-last rows (BatchNorm1d/2d/3d of nnet/normalizations.py:42-170; eps 1e-5, momentum 0.1).
 * finalize: from stats ([replicas][2C]: sum, sumsq over `count` rows; GEMM colstats use 32 replicas, others 1) -> scale/shift (gamma*rstd, beta-mean*gamma*rstd), saved mean /
 *   rstd, and the running-stat update (unbiased variance) when running_mean != NULL.  In eval mode call
 *   avec_bn_eval_affine instead.
 * apply: y = act(scale*u + shift (+ res))                act: 0 none, 1 ReLU, 2 Swish
 * bwd_reduce: sums[0:C] += sum dz, sums[C:2C] += sum dz*xhat,  dz = dy * act'(.)
 * bwd_apply: du = gamma*rstd*(dz - sums0/count - xhat*sums1/count);  dres = dz (optional)
 * ------------------------------------------------------------------------------------------------------------------ */
enum avec_act { AVEC_ACT_NONE = 0, AVEC_ACT_RELU = 1, AVEC_ACT_SWISH = 2 };
int avec_bn_finalize(const float* stats, const float* gamma, const float* beta, float* scale, float* shift, float* mean,
                     float* rstd, float* running_mean, float* running_var, long long count, int C, float eps,
                     float momentum, int replicas, avec_stream_t stream);
int avec_bn_eval_affine(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                        float* scale, float* shift, int C, float eps, avec_stream_t stream);
int avec_bn_stats(const void* u, int dtype, long long rows, int C, float* stats, avec_stream_t stream);
int avec_bn_apply(const void* u, const float* scale, const float* shift, const void* res, void* y, long long rows, int C,
                  int act, int dtype, long long ldy, avec_stream_t stream);
int avec_bn_bwd_reduce(const void* dy, const void* u, const float* scale, const float* shift, const void* res,
                       const float* mean, const float* rstd, float* sums, long long rows, int C, int act, int dtype,
                       avec_stream_t stream);
int avec_bn_bwd_apply(const void* dy, const void* u, const float* scale, const float* shift, const void* res,
                      const float* mean, const float* rstd, const float* gamma, const float* sums, void* du, void* dres,
                      long long rows, int C, int act, int dtype, avec_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Audio front-end: torchaudio Spectrogram(512, 400, 160, hann, center, reflect, power 2) -> MelScale(80, htk) ->
 * log(x + 1e-9)  (nnet/preprocessing.py:57-85).  wave [B, L] fp32 -> out fp32; layout 0: [B, F, 80] (frames-major,
 * what the fused stem consumes), layout 1: [B, 80, F] (the reference module's output).  fb is the [257, 80] filterbank.
 * ------------------------------------------------------------------------------------------------------------------ */
int avec_stft_mel_log(const float* wave, const float* fb, float* out, int B, int L, int F, int layout,
                      avec_stream_t stream);

/* Visual stem Conv3d(1->64, k(5,7,7), s(1,2,2), "same") as a direct tcgen05 implicit GEMM (im2col tile built in shared memory;
 * replaces layers.Conv3d.forward of nnet/networks.py:459-471 / nnet/layers.py:326-503 for the single-channel video input).
 * x [Nb][T][H][W] bf16 (W = 88: a 128-site tile spans <= 4 output rows; H even); wp [64][320] bf16 with k = (kt*7+kh)*8 + kw (kw = 7 and k >= 280 zero);
 * out [Nb*T*(H/2)*(W/2)][64] bf16 = conv + bias; colstats (optional) [AVEC_STATS_REPLICAS][2][64] fp32 BatchNorm sums. */
int avec_stem3d_fwd(const void* x, const void* wp, const float* bias, void* out, float* colstats, int Nb, int T, int H, int W,
                    avec_stream_t stream);
/* weight gradient of the same convolution: dw [64][245] fp32 += sum_sites dy[site][co] * x[window(site)][tap] (dy bf16 [sites][64]) */
int avec_stem3d_wgrad(const void* x, const void* dy, float* dw, int Nb, int T, int H, int W, avec_stream_t stream);


/* Audio stem Conv2d(1 -> Co, 3x3, stride 2, pad 1) on the log-mel image x [N][H][W] (layers.Conv2d of the subsampling module,
 * nnet/networks.py:359-368, nnet/modules.py:70-130): out [N*Ho*Wo][Co] = conv + bias, w [Co][9] in the compute dtype,
 * colstats (optional) [AVEC_STATS_REPLICAS][2][Co] fp32 BatchNorm sums; wgrad: dw [Co][9] fp32 += sum_sites dy * taps. */
int avec_stem2d_fwd(const void* x, const void* w, const float* bias, void* out, float* colstats, int N, int H, int W, int Co, int dtype,
                    avec_stream_t stream);
int avec_stem2d_wgrad(const void* x, const void* dy, float* dw, int N, int H, int W, int Co, int dtype, avec_stream_t stream);


/* single-channel im2col for the C = 1 stems (Conv3d (5,7,7) of nnet/networks.py:460-468): x [N,Ti,Hi,Wi,1] ->
 * col [sites, Kpad] (taps then zero padding), so that the stem convolution and its weight gradient run as plain GEMMs */
int avec_im2col_c1(const void* x, void* col, const avec_conv_geom* geom, int Kpad, int dtype, avec_stream_t stream);

/* BatchNorm backward (batch statistics, no activation) whose upstream gradient is the 3x3 / stride-2 max-pool backward of
 * avec_bn_relu_maxpool_fwd, gathered on the fly from the pooled gradient dyp [N,Ho,Wo,C] and the saved argmax codes: both
 * passes (column sums -> sums [2C] = {sum dz, sum dz*xhat}, then du [N,Hi,Wi,C]) without materialising dz.  Replaces
 * MaxPool3d.backward + BatchNorm3d.backward of the visual stem (nnet/networks.py:459-471, nnet/layers.py:839-915). */
int avec_bn_bwd_pool(const void* dyp, const uint8_t* idx, const void* u, const float* mean, const float* rstd, const float* gamma,
                     float* sums, void* du, int N, int Hi, int Wi, int C, int Ho, int Wo, int dtype, avec_stream_t stream);


/* ------------------------------------------------------------------------------------------------------------------
 * Visual front-end helpers (nnet/networks.py:459-472): BN3d + ReLU + MaxPool3d((1,3,3), s (1,2,2), zero "same" pad)
 * fused: u [N,Hi,Wi,C] -> y [N,Ho,Wo,C], argmax index (0..8, uint8) saved for the backward.
 * bwd: dz [N,Hi,Wi,C] = relu'(.) * scatter(dy)  (gather form, deterministic).
 * ------------------------------------------------------------------------------------------------------------------ */
int avec_bn_relu_maxpool_fwd(const void* u, const float* scale, const float* shift, void* y, uint8_t* idx, int N, int Hi,
                             int Wi, int C, int Ho, int Wo, int dtype, avec_stream_t stream);
int avec_bn_relu_maxpool_bwd(const void* dy, const uint8_t* idx, void* dz, int N, int Hi, int Wi, int C, int Ho, int Wo,
                             int dtype, avec_stream_t stream);
/* GlobalAvgPool2d (nnet/networks.py:129-132): x [N, HW, C] -> y [N, C]; bwd broadcasts dy / HW */
int avec_avgpool_fwd(const void* x, void* y, int N, int HW, int C, int dtype, avec_stream_t stream);
int avec_avgpool_bwd(const void* dy, void* dx, int N, int HW, int C, int dtype, avec_stream_t stream);

/* zero insertion out[n, i*s, j*s, :] = in[n, i, j, :] ([N,Ho,Wo,C] -> [N,Hi,Wi,C]): the input gradient of a stride-s
 * convolution is then a stride-1 convolution over the zero-inserted dY (TMA-fed implicit GEMM) */
int avec_zero_upsample(const void* in, void* out, int N, int Ho, int Wo, int Hi, int Wi, int C, int s, int dtype,
                       avec_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * CTC loss fused with log-softmax (nnet/losses.py:292-334: log_softmax -> transpose -> nn.CTCLoss(reduction="none")).
 * logits [B,T,V] fp32, labels [B,Lmax] int64, in_len / lab_len [B] int64 on the DEVICE (in_len may be NULL = T).
 * nll [B] = per-utterance negative log-likelihood; grad [B,T,V] = d nll_b / d logits (0 beyond in_len);
 * ws: workspace of B*T*(2*Lmax+1) floats.  zero_infinity: infeasible alignments give nll = 0, grad = 0.
 * ------------------------------------------------------------------------------------------------------------------ */
int avec_ctc_loss(const float* logits, const long long* labels, const long long* in_len, const long long* lab_len, float* nll,
                  float* grad, float* ws, int B, int T, int V, int Lmax, int blank, int zero_infinity, avec_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Training-step kernels downstream / upstream of the encoder (csrc/train.cu).
 * Random draws are Philox4x32-10(counter = (row, column >> 3, site, step), key = seed) with rng_state = {seed, step} two
 * uint64 in DEVICE memory; avec_counter_advance bumps a device uint64 (the RNG step, the optimizer step) from inside a
 * captured CUDA graph, so every replay draws fresh masks.  oracle/train_oracle.py restates the generator in numpy.
 * ------------------------------------------------------------------------------------------------------------------ */
int avec_counter_advance(unsigned long long* counter, avec_stream_t stream);

/* nn.Dropout (nnet/modules.py:281-283,315,380; nnet/networks.py:221,269) fused with the residual add that follows it:
 * y[r][c] = (res ? res[r][c] : 0) + alpha * keep(r,c) * x[r][c] / (1-p);  element (r,c) is kept iff its 16 random bits
 * (16-bit lane (c & 7) of the draw of column group c >> 3) are >= round(p * 65536).  The backward calls the same
 * function on the incoming gradient with the same site (res = NULL): the mask is regenerated, never stored.  x == y ok.
 * P > 1: x is [B, Tp, C] (one row per patch of P frames), y / res are [B, T, C] with rows = B*T: the dropout that follows
 * the patch attention's upsampling (nnet/attentions.py:368-372) draws one mask element per FRAME, as the reference does. */
int avec_dropout(const void* x, const void* res, void* y, long long rows, int C, int dtype, float p, float alpha,
                 const unsigned long long* rng_state, int site, int T, int Tp, int P, long long ldy, avec_stream_t stream);

/* SpecAugment (nnet/preprocessing.py:87-129 over torchaudio mask_along_axis), in place on mel [B,F,M] fp32 (frame-major):
 * mF frequency masks of width < Fmax shared by the batch, mT time masks per utterance of width < int(pS * len_b) inside
 * [0, len_b).  lengths [B] int64 on the device (NULL = F).  intervals (optional) [B][mF+mT][2] int32 receives [lo, hi). */
int avec_spec_augment(float* mel, const long long* lengths, int B, int F, int M, int mF, int Fmax, int mT, float pS,
                      const unsigned long long* rng_state, int site, int* intervals, avec_stream_t stream);

/* Video augmentation of the reference's training configs (configs/LRS23/AV/EffConfInterCTC.py:82-88), on the padded batch:
 * RandomCrop(Ho, Wo) -> RandomHorizontalFlip(flip_p) -> nnet.TimeMaskSecond (nnet/transforms.py:108-126: int(len_b / fps *
 * num_mask_second) time masks of width < mask_T frames, filled with the running mean of the clip).  in [B,T,Hi,Wi] fp32 ->
 * out [B,T,Ho,Wo] fp32, frames >= lengths[b] zero; frame_sums: [B*T] float scratch; draws (optional) [B][3 + 2*32] int32 receives
 * {crop row, crop column, flip, (lo, hi) per mask}.  Same counter-based generator as avec_dropout (site selects the stream). */
int avec_video_augment(const float* in, const long long* lengths, float* out, float* frame_sums, int B, int T, int Hi, int Wi, int Ho,
                       int Wo, float flip_p, int mask_T, float fps, float num_mask_second, const unsigned long long* rng_state,
                       int site, int* draws, avec_stream_t stream);

/* Greedy CTC decoding (nnet/decoders.py:97-120): argmax (first maximum) -> frames < in_len -> merge repeats -> drop blanks.
 * logits [B,T,V] fp32, in_len [B] int64 device (NULL = T); align [B,T] int32 frame-level argmax (-1 beyond the length,
 * may be NULL), tokens [B,T] int32 padded with -1, ntok [B] int32. */
int avec_ctc_greedy_decode(const float* logits, const long long* in_len, int* align, int* tokens, int* ntok, int B, int T, int V,
                           int blank, avec_stream_t stream);

/* Fused optimizer over FLAT fp32 buffers of n elements (n % 4 == 0, 16-byte aligned): torch.optim.Adam with L2 weight decay
 * (nnet/optimizers.py:61-93), learning rate from the device step counter (lr_mode 0: lr_a; 1: Noam, nnet/schedulers.py:
 * 120-137, lr_a * min(t * lr_b^-1.5, t^-0.5)), optional global-norm clipping (sumsq = device float holding sum g^2, from
 * avec_sumsq; nnet/model.py:378-380) and optional EMA copy (nnet/model.py:401-404).  lr_out (optional) [2] = {lr, |g|}. */
int avec_sumsq(const float* g, long long n, float* out, avec_stream_t stream);
int avec_adam_step(float* p, const float* g, float* m, float* v, float* ema, long long n, float beta1, float beta2, float eps,
                   float weight_decay, int lr_mode, float lr_a, float lr_b, float max_norm, float ema_tau,
                   const unsigned long long* step, const float* sumsq, float* lr_out, avec_stream_t stream);

/* Multi-tensor strided copy / conversion: ONE launch re-lays out every parameter of a model for the kernels (fp32 masters ->
 * bf16 [N, K] GEMM operands with TMA-able pitch, conv filters permuted to [Co][tap][Ci] / [Ci][tap][Co], Q/K/V stacked with every
 * head zero-padded to its column block, ...).  Job j copies the 4-d index space n from src (element strides ss) to dst (element
 * strides ds); every job holds < 2^31 elements.  jobs_dev is the table in DEVICE memory; chunks_dev [nchunks][2] int32 maps CTA c to
 * (job index, chunk index): the CTA copies elements [chunk * 4096, (chunk + 1) * 4096) of that job (AVEC_COPY_CHUNK). */
#define AVEC_COPY_CHUNK 4096
typedef struct avec_copy_job {
    const void* src;
    void* dst;
    long long start;
    long long ss[4], ds[4];
    int n[4];
    int src_dtype, dst_dtype;
} avec_copy_job;
int avec_convert_multi(const avec_copy_job* jobs_dev, const int* chunks_dev, int nchunks, avec_stream_t stream);
/* padded-heads gradient -> dense (fp32): dst[(g*d + r)*K + c] = src[(g*dp + r)*K + c], g < groups, r < d, c < K */
int avec_unpad_heads(const float* src, float* dst, long long groups, int d, int dp, long long K, avec_stream_t stream);

/* dtype conversion / strided copy helper: dst[r][c] = (T)src[r][c] */
int avec_convert(const void* src, int src_dtype, long long lds, void* dst, int dst_dtype, long long ldd, long long rows,
                 int C, avec_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* AVEC_B200_H */
