// Relative-position multi-head self-attention as a Blackwell tile kernel: tcgen05.mma with the score / band / output
// accumulators in TMEM, operands brought by TMA, flash-style (no [B,H,T,T] tensor is ever written: the forward saves one
// log-sum-exp per query row, the backward recomputes the probabilities).
// Replaces RelPos1dMultiHeadAttention.forwardQKV + rel_to_abs (reference nnet/attentions.py:280-323, 234-278) - and, on pooled
// tokens, RelPosPatch1dMultiHeadAttention (attentions.py:348-382) - and their autograd backward, for any sequence length.
//
// Layout ("padded heads"): the Q/K/V/E/O matrices keep every head in its own dp = 64 * ndb column block (dp >= d, pad columns
// are exact zeros because the projection weights are zero-padded), so one TMA box = one 64-wide k-block of one head:
//   qkv [B*T, 3*H*dp]: head h of part s (0 q, 1 k, 2 v) at columns (s*H + h)*dp;  e [2T-1, H*dp];  o [B*T, H*dp].
// One CTA = one (item, head, 128-query block); it walks the key blocks of 128 keys.  Per key block:
//   S    = Q K^T                      (128 x 128, TMEM columns [0,128))
//   band = Q Eb^T                     (128 x 256, TMEM [128,384)): Eb = the 255 rows of e a 128 x 128 score tile can touch,
//                                      band[i][c] = q_i . e[T-1 + j0 - i0 - 127 + c]
//   s[i][j] = (S[i][j] + band[i][j - i + 127]) / sqrt(d)      <- rel_to_abs: a per-row shift, done by the row's own thread
//                                                                 through a private shared-memory window (64 columns per 32 keys)
//   forward:  online softmax, P (bf16) -> shared memory, O_blk = P V (TMEM [0, dp): S / band are dead by then), rescaled accumulation
//   backward: dP = dO V^T (TMEM [384,512)); p = exp(s - lse); dS = p (dP - delta) / sqrt(d); P, dS and the band-layout copy
//             dSb[i][j - i + 127] go to shared memory as bf16 operands, then per 64-wide head-dim block
//             dV = P^T dO, dK = dS^T Q, dQ = dS K + dSb Eb, dEb = dSb^T Q (-> atomically into de).
// Warps 0-3: one thread per query row (TMEM lane = row); warp 4: TMA + MMA issue by one elected lane.  The head dimension is
// streamed in 64-column blocks through one shared-memory ring (loads and MMAs of a CTA are serialised; the grid supplies the
// parallelism: B*H*ceil(T/128) CTAs).
#include "tc_common.cuh"
#include <cstring>
#include <cmath>
#include <algorithm>

namespace {
using namespace tcx;

constexpr int AT_THREADS = 160;
constexpr int STG_LD = 66;                       // private skew window: 64 columns + pad (float2 stores / shifted scalar loads, conflict free)
constexpr int STG_BYTES = 4 * 32 * STG_LD * 4;   // 33792
constexpr int TILE = 16384;                      // one 128-row x 64-column bf16 operand block
constexpr int RING_FWD = 5 * TILE;               // Q | K | Eb (2 blocks) | Qp   (the PV pieces reuse block 0 for V)
constexpr int RING_BWD = 6 * TILE;               // phase A: Q | K | Eb | Qp, then dO | V;  phase B: dO | Q | K | Eb | Qp
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

struct AttnTcParams {
    int B, T, H, d, dp, ndb, nqb, nkb;
    int qp_part;                      // column part of the band's query operand (q + v): 3 for grouped attention, 0 = the same q
    const int* klen;
    int qlen;
    float scale, scale_log2;
    // forward
    bf16* o; long long ld_o;
    float* lse;                       // [B, H, T]
    // backward
    const bf16* d_o; long long ld_do;
    const bf16* o_in; long long ld_oin;
    bf16* dqkv; long long ld_dqkv;    // direct bf16 output when the item has one query block and one key block
    float* dqkv_ws;                   // fp32 accumulation buffer (same shape / pitch as dqkv) otherwise
    float* de; long long ld_de;       // [2T-1, H*dp] fp32, atomically accumulated
};

__device__ __forceinline__ uint8_t* align1024(uint8_t* p) { return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~(uintptr_t)1023); }
__device__ __forceinline__ uint64_t desc_k(uint32_t saddr) { return make_smem_desc(saddr, 16, 1024); }
__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr) { return make_smem_desc(saddr, TILE, 1024); }
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}
// byte offset of element (row r, column c) of a [128 rows][64*nblk columns] bf16 tile stored as 64-column blocks of 128-byte
// swizzled rows: the K-major operand (m = r, k = c) and the MN-major operand (k = r, mn = c) are the same bytes
__device__ __forceinline__ uint32_t tile_off(int r, int c) {
    return (uint32_t)((c >> 6) * TILE + r * 128 + ((((c & 63) >> 3) ^ (r & 7)) << 4) + (c & 7) * 2);
}

// s_band[k] = band[row][base + 31 - lane + k], k < 32: the thread parks 64 band columns of its own row in its private window
__device__ __forceinline__ void band_window(uint32_t lane_taddr, int base_col, float* win, int lane, float (&out)[32]) {
    float v[32];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        tmem_ld32(lane_taddr + (uint32_t)(128 + base_col + half * 32), v);
#pragma unroll
        for (int c = 0; c < 32; c += 2) *reinterpret_cast<float2*>(win + half * 32 + c) = make_float2(v[c], v[c + 1]);
    }
#pragma unroll
    for (int k = 0; k < 32; ++k) out[k] = win[31 - lane + k];
}

struct Bars { uint64_t tma, ring, acc, work; uint32_t tmem; };

__device__ __forceinline__ void setup(Bars* bars, int tid, int warp, const CUtensorMap* m0, const CUtensorMap* m1, const CUtensorMap* m2) {
    if (tid == 0) {
        mbar_init(&bars->tma, 1); mbar_init(&bars->ring, 1); mbar_init(&bars->acc, 1); mbar_init(&bars->work, 128);
        fence_barrier_init();
        tma_prefetch_desc(m0); tma_prefetch_desc(m1);
        if (m2) tma_prefetch_desc(m2);
    }
    if (warp == 4) tmem_alloc(&bars->tmem, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
}

// S (+)= Q K^T and band (+)= Q Eb^T for one 64-column head-dim block resident in the ring
__device__ __forceinline__ void mma_scores(uint32_t tmem, uint32_t q_s, uint32_t k_s, uint32_t e_s, uint32_t qp_s, bool first) {
    const uint32_t id_s = make_idesc(128, 0, 0), id_b = make_idesc(256, 0, 0);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_f16(tmem, desc_k(q_s + k * 32), desc_k(k_s + k * 32), id_s, (!first || k > 0) ? 1u : 0u);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_f16(tmem + 128, desc_k(qp_s + k * 32), desc_k(e_s + k * 32), id_b, (!first || k > 0) ? 1u : 0u);
}

// ================================================================ forward
__global__ void __launch_bounds__(AT_THREADS, 1) relpos_attn_tc_fwd_kernel(const __grid_constant__ AttnTcParams p, const __grid_constant__ CUtensorMap mapQKV,
                                                                          const __grid_constant__ CUtensorMap mapE) {
    pdl_trigger();   // programmatic dependent launch: a tcgen05 GEMM behind this kernel may start its prologue now
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = align1024(smem_raw);
    uint8_t* ring = sm;
    uint8_t* ptile = sm + RING_FWD;                                   // 2 blocks
    float* oacc = reinterpret_cast<float*>(ptile + 2 * TILE);         // [128][dp + 1] when nkb > 1
    Bars* bars = reinterpret_cast<Bars*>(reinterpret_cast<uint8_t*>(oacc) + (p.nkb > 1 ? 128 * (p.dp + 1) * 4 : 0));
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    int idx = blockIdx.x;
    const int qb = idx % p.nqb; idx /= p.nqb;
    const int h = idx % p.H, b = idx / p.H;
    const int i0 = qb * 128, T = p.T, H = p.H, dp = p.dp, ndb = p.ndb;
    setup(bars, tid, warp, &mapQKV, &mapE, nullptr);
    pdl_wait();      // barriers / TMEM / descriptor prefetch above overlapped the kernel in front; global memory from here on
    const uint32_t tmem = bars->tmem;
    const uint32_t ring_s = smem_u32(ring), p_s = smem_u32(ptile);

    if (warp == 4) {
        // ------------------------------------------------ control warp: TMA + MMA issue
        uint32_t tma_ph = 0, ring_ph = 0, work_ph = 0;
        const uint32_t id_o = make_idesc(64, 0, 1);
        const bool sep_qp = p.qp_part != 0;
        for (int jb = 0; jb < p.nkb; ++jb) {
            const int j0 = jb * 128, r0 = T - 1 + j0 - i0 - 127;
            for (int db = 0; db < ndb; ++db) {
                if (elect_one()) {
                    mbar_expect_tx(&bars->tma, (sep_qp ? 5 : 4) * TILE);
                    tma_load_3d(ring_s, &mapQKV, &bars->tma, h * dp + db * 64, i0, b);
                    tma_load_3d(ring_s + TILE, &mapQKV, &bars->tma, (H + h) * dp + db * 64, j0, b);
                    tma_load_2d(ring_s + 2 * TILE, &mapE, &bars->tma, h * dp + db * 64, r0);
                    if (sep_qp) tma_load_3d(ring_s + 4 * TILE, &mapQKV, &bars->tma, (p.qp_part * H + h) * dp + db * 64, i0, b);
                }
                __syncwarp();
                mbar_wait(&bars->tma, tma_ph); tma_ph ^= 1u;
                tc_fence_after();
                if (elect_one()) {
                    mma_scores(tmem, ring_s, ring_s + TILE, ring_s + 2 * TILE, sep_qp ? ring_s + 4 * TILE : ring_s, db == 0);
                    umma_commit(&bars->ring);
                    if (db == ndb - 1) umma_commit(&bars->acc);
                }
                __syncwarp();
                mbar_wait(&bars->ring, ring_ph); ring_ph ^= 1u;
            }
            mbar_wait(&bars->work, work_ph); work_ph ^= 1u;    // P of this key block is in shared memory, S / band are drained
            tc_fence_after();
            for (int db = 0; db < ndb; ++db) {
                if (elect_one()) {
                    mbar_expect_tx(&bars->tma, TILE);
                    tma_load_3d(ring_s, &mapQKV, &bars->tma, (2 * H + h) * dp + db * 64, j0, b);
                }
                __syncwarp();
                mbar_wait(&bars->tma, tma_ph); tma_ph ^= 1u;
                tc_fence_after();
                if (elect_one()) {
                    // O[:, db block] = P (K-major, k = key) x V (MN-major: n = head dim, k = key rows)
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        umma_f16(tmem + db * 64, desc_k(p_s + (k >> 2) * TILE + (k & 3) * 32), desc_mn(ring_s + k * 2048), id_o, k > 0 ? 1u : 0u);
                    umma_commit(&bars->ring);
                    if (db == ndb - 1) umma_commit(&bars->acc);
                }
                __syncwarp();
                mbar_wait(&bars->ring, ring_ph); ring_ph ^= 1u;
            }
            mbar_wait(&bars->work, work_ph); work_ph ^= 1u;    // O block drained
            tc_fence_after();
        }
    } else {
        // ------------------------------------------------ workers: thread = query row
        const int r = warp * 32 + lane, i = i0 + r;
        const bool row_in = i < T;
        const int kl = p.klen ? min(max(p.klen[b], 0), T) : T;
        const bool full_mask = i >= p.qlen || kl <= 0;    // every key masked: the reference's -1e9 leaves a uniform softmax over all T keys
        const uint32_t lane_t = tmem + ((uint32_t)(warp * 32) << 16);
        float* win = reinterpret_cast<float*>(ring) + (warp * 32 + lane) * STG_LD;
        float* orow = oacc + r * (dp + 1);
        uint32_t acc_ph = 0;
        float m_run = -INFINITY, l_run = 0.0f;
        for (int jb = 0; jb < p.nkb; ++jb) {
            const int j0 = jb * 128;
            mbar_wait(&bars->acc, acc_ph); acc_ph ^= 1u;
            tc_fence_after();
            float s[128];
#pragma unroll
            for (int c0 = 0; c0 < 128; c0 += 32) {
                float bnd[32], sc[32];
                band_window(lane_t, c0 + 96 - 32 * warp, win, lane, bnd);
                tmem_ld32(lane_t + (uint32_t)c0, sc);
#pragma unroll
                for (int k = 0; k < 32; ++k) s[c0 + k] = (sc[k] + bnd[k]) * p.scale_log2;
            }
            float mblk = -INFINITY;
#pragma unroll
            for (int k = 0; k < 128; ++k) {
                const int j = j0 + k;
                const bool valid = full_mask ? (j < T) : (j < kl);
                s[k] = valid ? (full_mask ? 0.0f : s[k]) : -INFINITY;
                mblk = fmaxf(mblk, s[k]);
            }
            const float m_new = fmaxf(m_run, mblk);
            const float alpha = (m_run == -INFINITY) ? 0.0f : exp2f(m_run - m_new);
            float sum = 0.0f;
#pragma unroll
            for (int k = 0; k < 128; k += 8) {
                float pv[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) { pv[u] = (s[k + u] == -INFINITY) ? 0.0f : exp2f(s[k + u] - m_new); sum += pv[u]; }
                *reinterpret_cast<uint4*>(ptile + tile_off(r, k)) = make_uint4(pack2(pv[0], pv[1]), pack2(pv[2], pv[3]), pack2(pv[4], pv[5]), pack2(pv[6], pv[7]));
            }
            l_run = l_run * alpha + sum;
            m_run = m_new;
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(&bars->work);
            // ---- O block
            mbar_wait(&bars->acc, acc_ph); acc_ph ^= 1u;
            tc_fence_after();
            const bool last = jb == p.nkb - 1;
            const float inv = 1.0f / l_run;
            for (int db = 0; db < ndb; ++db) {
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    float ov[32];
                    tmem_ld32(lane_t + (uint32_t)(db * 64 + half * 32), ov);
                    const int cb = db * 64 + half * 32;
                    if (p.nkb > 1) {
#pragma unroll
                        for (int c = 0; c < 32; ++c) {
                            ov[c] += (jb == 0) ? 0.0f : orow[cb + c] * alpha;
                            if (!last) orow[cb + c] = ov[c];
                        }
                    }
                    if (last && row_in) {
                        bf16* dst = p.o + ((size_t)b * T + i) * p.ld_o + h * dp + cb;
#pragma unroll
                        for (int c = 0; c < 32; c += 8)
                            *reinterpret_cast<uint4*>(dst + c) = make_uint4(pack2(ov[c] * inv, ov[c + 1] * inv), pack2(ov[c + 2] * inv, ov[c + 3] * inv),
                                                                            pack2(ov[c + 4] * inv, ov[c + 5] * inv), pack2(ov[c + 6] * inv, ov[c + 7] * inv));
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&bars->work);
        }
        if (row_in) p.lse[((size_t)b * H + h) * T + i] = (m_run + log2f(l_run)) * LN2;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// ================================================================ backward
__global__ void __launch_bounds__(AT_THREADS, 1) relpos_attn_tc_bwd_kernel(const __grid_constant__ AttnTcParams p, const __grid_constant__ CUtensorMap mapQKV,
                                                                          const __grid_constant__ CUtensorMap mapE, const __grid_constant__ CUtensorMap mapDO) {
    pdl_trigger();   // programmatic dependent launch: a tcgen05 GEMM behind this kernel may start its prologue now
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = align1024(smem_raw);
    uint8_t* ring = sm;
    uint8_t* ptile = sm + RING_BWD;              // P   [128 i][128 j]   2 blocks
    uint8_t* dstile = ptile + 2 * TILE;          // dS  [128 i][128 j]   2 blocks
    uint8_t* dsb = dstile + 2 * TILE;            // dSb [128 i][256 c]   4 blocks
    Bars* bars = reinterpret_cast<Bars*>(dsb + 4 * TILE);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    int idx = blockIdx.x;
    const int qb = idx % p.nqb; idx /= p.nqb;
    const int h = idx % p.H, b = idx / p.H;
    const int i0 = qb * 128, T = p.T, H = p.H, dp = p.dp, ndb = p.ndb;
    setup(bars, tid, warp, &mapQKV, &mapE, &mapDO);
    pdl_wait();      // (as in the forward kernel)
    const uint32_t tmem = bars->tmem;
    const uint32_t ring_s = smem_u32(ring), p_s = smem_u32(ptile), ds_s = smem_u32(dstile), dsb_s = smem_u32(dsb);
    const bool single = p.nqb == 1 && p.nkb == 1;

    if (warp == 4) {
        uint32_t tma_ph = 0, ring_ph = 0, work_ph = 0;
        const uint32_t id_dp = make_idesc(128, 0, 0), id_t = make_idesc(64, 1, 1), id_q = make_idesc(64, 0, 1);
        const bool sep_qp = p.qp_part != 0;
        for (int jb = 0; jb < p.nkb; ++jb) {
            const int j0 = jb * 128, r0 = T - 1 + j0 - i0 - 127;
            // ---- phase A: S, band, dP, head dim streamed in 64-column blocks
            for (int db = 0; db < ndb; ++db) {
                if (elect_one()) {
                    mbar_expect_tx(&bars->tma, (sep_qp ? 5 : 4) * TILE);
                    tma_load_3d(ring_s, &mapQKV, &bars->tma, h * dp + db * 64, i0, b);
                    tma_load_3d(ring_s + TILE, &mapQKV, &bars->tma, (H + h) * dp + db * 64, j0, b);
                    tma_load_2d(ring_s + 2 * TILE, &mapE, &bars->tma, h * dp + db * 64, r0);
                    if (sep_qp) tma_load_3d(ring_s + 4 * TILE, &mapQKV, &bars->tma, (p.qp_part * H + h) * dp + db * 64, i0, b);
                }
                __syncwarp();
                mbar_wait(&bars->tma, tma_ph); tma_ph ^= 1u;
                tc_fence_after();
                if (elect_one()) { mma_scores(tmem, ring_s, ring_s + TILE, ring_s + 2 * TILE, sep_qp ? ring_s + 4 * TILE : ring_s, db == 0); umma_commit(&bars->ring); }
                __syncwarp();
                mbar_wait(&bars->ring, ring_ph); ring_ph ^= 1u;
                if (elect_one()) {
                    mbar_expect_tx(&bars->tma, 2 * TILE);
                    tma_load_3d(ring_s, &mapDO, &bars->tma, h * dp + db * 64, i0, b);
                    tma_load_3d(ring_s + TILE, &mapQKV, &bars->tma, (2 * H + h) * dp + db * 64, j0, b);
                }
                __syncwarp();
                mbar_wait(&bars->tma, tma_ph); tma_ph ^= 1u;
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_f16(tmem + 384, desc_k(ring_s + k * 32), desc_k(ring_s + TILE + k * 32), id_dp, (db > 0 || k > 0) ? 1u : 0u);
                    umma_commit(&bars->ring);
                    if (db == ndb - 1) umma_commit(&bars->acc);
                }
                __syncwarp();
                mbar_wait(&bars->ring, ring_ph); ring_ph ^= 1u;
            }
            mbar_wait(&bars->work, work_ph); work_ph ^= 1u;    // P, dS, dSb written; S / band / dP drained
            tc_fence_after();
            // ---- phase B: per head-dim block  dV | dK | dQ | dEb (2 x 128 rows)  ->  TMEM [0,64) [64,128) [128,192) [192,256) [256,320)
            for (int db = 0; db < ndb; ++db) {
                if (elect_one()) {
                    mbar_expect_tx(&bars->tma, (sep_qp ? 6 : 5) * TILE);
                    tma_load_3d(ring_s, &mapDO, &bars->tma, h * dp + db * 64, i0, b);
                    tma_load_3d(ring_s + TILE, &mapQKV, &bars->tma, h * dp + db * 64, i0, b);
                    tma_load_3d(ring_s + 2 * TILE, &mapQKV, &bars->tma, (H + h) * dp + db * 64, j0, b);
                    tma_load_2d(ring_s + 3 * TILE, &mapE, &bars->tma, h * dp + db * 64, r0);
                    if (sep_qp) tma_load_3d(ring_s + 5 * TILE, &mapQKV, &bars->tma, (p.qp_part * H + h) * dp + db * 64, i0, b);
                }
                __syncwarp();
                mbar_wait(&bars->tma, tma_ph); tma_ph ^= 1u;
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t do_s = ring_s, q_s = ring_s + TILE, k_s = ring_s + 2 * TILE, e_s = ring_s + 3 * TILE;
                    const uint32_t qp_s = sep_qp ? ring_s + 5 * TILE : q_s;
                    const uint32_t dqp_col = sep_qp ? 320u : 128u;     // grouped attention: d(q + v) is a separate output
#pragma unroll
                    for (int k = 0; k < 8; ++k) {   // reduction over the 128 query rows, 16 at a time
                        umma_f16(tmem, desc_mn(p_s + k * 2048), desc_mn(do_s + k * 2048), id_t, k > 0 ? 1u : 0u);                   // dV  = P^T dO
                        umma_f16(tmem + 64, desc_mn(ds_s + k * 2048), desc_mn(q_s + k * 2048), id_t, k > 0 ? 1u : 0u);             // dK  = dS^T Q
                        umma_f16(tmem + 192, desc_mn(dsb_s + k * 2048), desc_mn(qp_s + k * 2048), id_t, k > 0 ? 1u : 0u);          // dEb = dSb^T Qp (c < 128)
                        umma_f16(tmem + 256, desc_mn(dsb_s + 2 * TILE + k * 2048), desc_mn(qp_s + k * 2048), id_t, k > 0 ? 1u : 0u); //               (c >= 128)
                    }
#pragma unroll
                    for (int k = 0; k < 8; ++k)     // dQ = dS K: reduction over the 128 keys
                        umma_f16(tmem + 128, desc_k(ds_s + (k >> 2) * TILE + (k & 3) * 32), desc_mn(k_s + k * 2048), id_q, k > 0 ? 1u : 0u);
#pragma unroll
                    for (int k = 0; k < 16; ++k)    //    + dSb Eb: reduction over the 256 band rows
                        umma_f16(tmem + dqp_col, desc_k(dsb_s + (k >> 2) * TILE + (k & 3) * 32), desc_mn(e_s + k * 2048), id_q, (!sep_qp || k > 0) ? 1u : 0u);
                    umma_commit(&bars->ring);
                    umma_commit(&bars->acc);
                }
                __syncwarp();
                mbar_wait(&bars->ring, ring_ph); ring_ph ^= 1u;
                mbar_wait(&bars->work, work_ph); work_ph ^= 1u;    // the five accumulators are drained
                tc_fence_after();
            }
        }
    } else {
        const int r = warp * 32 + lane, i = i0 + r;
        const bool row_in = i < T;
        const int kl = p.klen ? min(max(p.klen[b], 0), T) : T;
        const bool full_mask = i >= p.qlen || kl <= 0;
        const uint32_t lane_t = tmem + ((uint32_t)(warp * 32) << 16);
        float* win = reinterpret_cast<float*>(ring) + (warp * 32 + lane) * STG_LD;
        uint32_t acc_ph = 0;
        // delta_i = dO_i . O_i, lse_i: per query row, once
        float delta = 0.0f, lse2 = 0.0f;
        if (row_in) {
            const bf16* go = p.d_o + ((size_t)b * T + i) * p.ld_do + h * dp;
            const bf16* oo = p.o_in + ((size_t)b * T + i) * p.ld_oin + h * dp;
            for (int c = 0; c < dp; c += 8) {
                float a[8], w[8];
                load_vec<8>(go + c, a);
                load_vec<8>(oo + c, w);
#pragma unroll
                for (int u = 0; u < 8; ++u) delta += a[u] * w[u];
            }
            lse2 = p.lse[((size_t)b * H + h) * T + i] * LOG2E;
        }
        for (int jb = 0; jb < p.nkb; ++jb) {
            const int j0 = jb * 128, r0 = T - 1 + j0 - i0 - 127;
            // the band-layout tile is sparse (128 of a row's 256 columns): clear this thread's row, then scatter
#pragma unroll
            for (int c = 0; c < 256; c += 8) *reinterpret_cast<uint4*>(dsb + tile_off(r, c)) = make_uint4(0u, 0u, 0u, 0u);
            mbar_wait(&bars->acc, acc_ph); acc_ph ^= 1u;
            tc_fence_after();
#pragma unroll 1
            for (int c0 = 0; c0 < 128; c0 += 32) {
                float bnd[32], sc[32], dpv[32], pv[32], dsv[32];
                band_window(lane_t, c0 + 96 - 32 * warp, win, lane, bnd);
                tmem_ld32(lane_t + (uint32_t)c0, sc);
                tmem_ld32(lane_t + (uint32_t)(384 + c0), dpv);
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                    const int j = j0 + c0 + k;
                    const bool valid = row_in && (full_mask ? (j < T) : (j < kl));
                    const float s2 = full_mask ? 0.0f : (sc[k] + bnd[k]) * p.scale_log2;
                    pv[k] = valid ? exp2f(s2 - lse2) : 0.0f;
                    dsv[k] = pv[k] * (dpv[k] - delta) * p.scale;
                }
#pragma unroll
                for (int k = 0; k < 32; k += 8) {
                    *reinterpret_cast<uint4*>(ptile + tile_off(r, c0 + k)) = make_uint4(pack2(pv[k], pv[k + 1]), pack2(pv[k + 2], pv[k + 3]), pack2(pv[k + 4], pv[k + 5]), pack2(pv[k + 6], pv[k + 7]));
                    *reinterpret_cast<uint4*>(dstile + tile_off(r, c0 + k)) = make_uint4(pack2(dsv[k], dsv[k + 1]), pack2(dsv[k + 2], dsv[k + 3]), pack2(dsv[k + 4], dsv[k + 5]), pack2(dsv[k + 6], dsv[k + 7]));
                }
                const int cbase = c0 - r + 127;    // band column of key c0 + k is cbase + k, in [0, 254]
#pragma unroll
                for (int k = 0; k < 32; ++k) *reinterpret_cast<bf16*>(dsb + tile_off(r, cbase + k)) = __float2bfloat16_rn(dsv[k]);
            }
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(&bars->work);
            // ---- phase B epilogues
            for (int db = 0; db < ndb; ++db) {
                mbar_wait(&bars->acc, acc_ph); acc_ph ^= 1u;
                tc_fence_after();
                const int j = j0 + r;     // key row of dV / dK
#pragma unroll 1
                const int nparts = p.qp_part != 0 ? 4 : 3;
                for (int part = 0; part < nparts; ++part) {      // 0: dV (row j), 1: dK (row j), 2: dQ (row i), 3: d(q + v) (row i, grouped)
                    const int row = part >= 2 ? i : j;
                    const int colblk = (part == 0 ? 2 * H + h : (part == 1 ? H + h : (part == 2 ? h : p.qp_part * H + h))) * dp + db * 64;
                    const uint32_t tcol = part == 3 ? 320u : (uint32_t)(part * 64);
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        float v[32];
                        tmem_ld32(lane_t + tcol + (uint32_t)(half * 32), v);
                        if (row < T) {
                            const size_t off = ((size_t)b * T + row) * p.ld_dqkv + colblk + half * 32;
                            if (single) {
#pragma unroll
                                for (int c = 0; c < 32; c += 8)
                                    *reinterpret_cast<uint4*>(p.dqkv + off + c) = make_uint4(pack2(v[c], v[c + 1]), pack2(v[c + 2], v[c + 3]), pack2(v[c + 4], v[c + 5]), pack2(v[c + 6], v[c + 7]));
                            } else {
#pragma unroll
                                for (int c = 0; c < 32; c += 4)
                                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p.dqkv_ws + off + c), "f"(v[c]), "f"(v[c + 1]), "f"(v[c + 2]), "f"(v[c + 3]) : "memory");
                            }
                        }
                    }
                }
#pragma unroll 1
                for (int part = 0; part < 2; ++part) {      // dEb rows c = part * 128 + r  ->  de[r0 + c]
                    const int er = r0 + part * 128 + r;
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        float v[32];
                        tmem_ld32(lane_t + (uint32_t)(192 + part * 64 + half * 32), v);
                        if (er >= 0 && er < 2 * T - 1 && part * 128 + r < 255) {
                            float* dst = p.de + (size_t)er * p.ld_de + h * dp + db * 64 + half * 32;
#pragma unroll
                            for (int c = 0; c < 32; c += 4)
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c), "f"(v[c]), "f"(v[c + 1]), "f"(v[c + 2]), "f"(v[c + 3]) : "memory");
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(&bars->work);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

bool make_maps(const void* qkv, long long ld_qkv, const void* e, long long ld_e, int B, int T, int H, int dp, int nparts, CUtensorMap* mq, CUtensorMap* me) {
    const cuuint64_t dq[3] = {(cuuint64_t)(nparts * H * dp), (cuuint64_t)T, (cuuint64_t)B};
    const cuuint64_t sq[2] = {(cuuint64_t)ld_qkv * 2, (cuuint64_t)ld_qkv * 2 * T};
    const cuuint32_t bq[3] = {64, 128, 1};
    const cuuint64_t de_[2] = {(cuuint64_t)(H * dp), (cuuint64_t)(2 * T - 1)};
    const cuuint64_t se[1] = {(cuuint64_t)ld_e * 2};
    const cuuint32_t be[2] = {64, 256};
    return encode_map(mq, qkv, 3, dq, sq, bq) && encode_map(me, e, 2, de_, se, be);
}
bool aligned16(const void* p, long long ld) { return (reinterpret_cast<uintptr_t>(p) % 16) == 0 && (ld * 2) % 16 == 0; }

int g_attr_done[2][64] = {{0}};
bool attr_once(int which, const void* fn, size_t bytes) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (g_attr_done[which][dev]) return true;
    if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) return false;
    g_attr_done[which][dev] = 1;
    return true;
}

// ================================================================ grouped attention: frames <-> tokens
// GroupedRelPosMultiHeadSelfAttention (reference nnet/attentions.py:579-650): a token is G consecutive frames concatenated
// (G*D1 wide) and split into H heads of dG = G*D1/H channels; frames past the real length Tf are zero rows; the content / position
// biases u, v (tiled over the group) are added AFTER the zero padding.  Element e = h*dG + c of token `tok` lives in frame
// tok*G + e / D1, column e % D1.  These kernels move between the frame-rate matrices of the projections and the padded-heads
// token layout of the tile kernel; G = 1 is the Transformer-XL variant of stages 2 / 3 (pure copy + bias).
template <typename T>
__global__ void __launch_bounds__(256) group_pack_kernel(const T* __restrict__ src, long long lds, int src_parts, const float* __restrict__ u,
                                                         const float* __restrict__ v, bf16* __restrict__ dst, long long ldd, int Bn, int Tf, int Tn,
                                                         int G, int H, int D1, int dG, int dp, int dst_parts) {
    // dst part s: 0 = q + u, 1 = k, 2 = v, 3 = q + v (dst_parts = 4);  dst_parts = 1: plain regroup of a 1-part matrix
    const long long total = (long long)Bn * Tn * dst_parts * H * dp;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % dp);
        long long r = i / dp;
        const int h = (int)(r % H); r /= H;
        const int s = (int)(r % dst_parts);
        const long long tokrow = r / dst_parts;
        const int tok = (int)(tokrow % Tn), b = (int)(tokrow / Tn);
        float val = 0.0f;
        if (c < dG) {
            const int e = h * dG + c, fi = e / D1, col = e - fi * D1, f = tok * G + fi;
            const int sp = dst_parts == 1 ? 0 : (s == 3 ? 0 : s);
            if (f < Tf) val = ldf(src + ((long long)b * Tf + f) * lds + (long long)sp * D1 + col);
            if (dst_parts == 4 && s == 0 && u) val += u[col];
            if (dst_parts == 4 && s == 3 && v) val += v[col];
        }
        dst[tokrow * ldd + ((long long)s * H + h) * dp + c] = __float2bfloat16_rn(val);
    }
}

// tokens -> frames for a 1-part matrix (attention output o, or the fp32 position gradient de)
template <typename TS, typename TD>
__global__ void __launch_bounds__(256) group_unpack_kernel(const TS* __restrict__ src, long long lds, TD* __restrict__ dst, long long ldd, int Bn, int Tf, int Tn,
                                                           int G, int H, int D1, int dG, int dp) {
    const long long total = (long long)Bn * Tf * D1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int col = (int)(i % D1);
        const long long fr = i / D1;
        const int f = (int)(fr % Tf), b = (int)(fr / Tf);
        const int tok = f / G, e = (f - tok * G) * D1 + col, h = e / dG, c = e - h * dG;
        stf(dst + fr * ldd + col, ldf(src + ((long long)b * Tn + tok) * lds + (long long)h * dp + c));
    }
}

// dqkv tokens [B*Tn, 4*H*dp] (d(q+u) | dk | dv | d(q+v)) -> frames [B*Tf, 3*D1] (dq = d(q+u) + d(q+v) | dk | dv);
// du[col] += sum d(q+u), dv[col] += sum d(q+v) over EVERY (token, group slot), the zero-padded frames included
__global__ void __launch_bounds__(256) group_unpack_dqkv_kernel(const bf16* __restrict__ src, long long lds, bf16* __restrict__ dst, long long ldd,
                                                                float* __restrict__ du, float* __restrict__ dv, int Bn, int Tf, int Tn, int G, int H,
                                                                int D1, int dG, int dp) {
    // thread = (column col of D1, group slot fi); blockIdx.y strides over (b, tok): partial bias sums stay in registers
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= D1) return;
    float su = 0.0f, sv = 0.0f;
    for (long long bt = blockIdx.y; bt < (long long)Bn * Tn; bt += gridDim.y) {
        const int tok = (int)(bt % Tn), b = (int)(bt / Tn);
        for (int fi = 0; fi < G; ++fi) {
            const int e = fi * D1 + col, h = e / dG, c = e - h * dG, f = tok * G + fi;
            const bf16* row = src + bt * lds;
            const float dqc = __bfloat162float(row[(long long)(0 * H + h) * dp + c]);
            const float dqp = __bfloat162float(row[(long long)(3 * H + h) * dp + c]);
            su += dqc; sv += dqp;
            if (f < Tf) {
                bf16* out = dst + ((long long)b * Tf + f) * ldd;
                out[col] = __float2bfloat16_rn(dqc + dqp);
                out[D1 + col] = row[(long long)(1 * H + h) * dp + c];
                out[2 * D1 + col] = row[(long long)(2 * H + h) * dp + c];
            }
        }
    }
    if (du) atomicAdd(du + col, su);
    if (dv) atomicAdd(dv + col, sv);
}

inline int ew_blocks_(long long total) { return (int)std::min<long long>((total + 255) / 256, 148LL * 16); }

}  // namespace


extern "C" int avec_attn_group_pack(const void* src, int src_dtype, long long lds, const float* u, const float* v, void* dst, long long ldd, int B,
                                    int Tf, int Tn, int G, int H, int D1, int dp, int dst_parts, avec_stream_t stream) {
    AVEC_CHECK_ARG(src && dst && B > 0 && Tf > 0 && Tn > 0 && G >= 1 && H > 0 && D1 > 0 && (G * D1) % H == 0 && dp >= G * D1 / H && (dst_parts == 1 || dst_parts == 4));
    AVEC_CHECK_ARG(Tf <= Tn * G && ldd >= (long long)dst_parts * H * dp);
    const int dG = G * D1 / H;
    const long long total = (long long)B * Tn * dst_parts * H * dp;
    if (src_dtype == AVEC_BF16)
        group_pack_kernel<bf16><<<ew_blocks_(total), 256, 0, as_stream(stream)>>>((const bf16*)src, lds, 3, u, v, (bf16*)dst, ldd, B, Tf, Tn, G, H, D1, dG, dp, dst_parts);
    else if (src_dtype == AVEC_F32)
        group_pack_kernel<float><<<ew_blocks_(total), 256, 0, as_stream(stream)>>>((const float*)src, lds, 3, u, v, (bf16*)dst, ldd, B, Tf, Tn, G, H, D1, dG, dp, dst_parts);
    else return AVEC_ERR_INVALID;
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}

extern "C" int avec_attn_group_unpack(const void* src, int src_dtype, long long lds, void* dst, int dst_dtype, long long ldd, int B, int Tf, int Tn, int G,
                                      int H, int D1, int dp, avec_stream_t stream) {
    AVEC_CHECK_ARG(src && dst && B > 0 && Tf > 0 && Tn > 0 && G >= 1 && H > 0 && D1 > 0 && (G * D1) % H == 0 && Tf <= Tn * G);
    const int dG = G * D1 / H;
    const long long total = (long long)B * Tf * D1;
    if (src_dtype == AVEC_BF16 && dst_dtype == AVEC_BF16)
        group_unpack_kernel<bf16, bf16><<<ew_blocks_(total), 256, 0, as_stream(stream)>>>((const bf16*)src, lds, (bf16*)dst, ldd, B, Tf, Tn, G, H, D1, dG, dp);
    else if (src_dtype == AVEC_F32 && dst_dtype == AVEC_F32)
        group_unpack_kernel<float, float><<<ew_blocks_(total), 256, 0, as_stream(stream)>>>((const float*)src, lds, (float*)dst, ldd, B, Tf, Tn, G, H, D1, dG, dp);
    else return AVEC_ERR_INVALID;
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}

extern "C" int avec_attn_group_unpack_dqkv(const void* src, long long lds, void* dst, long long ldd, float* du, float* dv, int B, int Tf, int Tn, int G,
                                           int H, int D1, int dp, avec_stream_t stream) {
    AVEC_CHECK_ARG(src && dst && B > 0 && Tf > 0 && Tn > 0 && G >= 1 && H > 0 && D1 > 0 && (G * D1) % H == 0 && Tf <= Tn * G);
    const int dG = G * D1 / H;
    dim3 grid((unsigned)cdiv(D1, 128), (unsigned)std::min<long long>((long long)B * Tn, 592));
    group_unpack_dqkv_kernel<<<grid, 128, 0, as_stream(stream)>>>((const bf16*)src, lds, (bf16*)dst, ldd, du, dv, B, Tf, Tn, G, H, D1, dG, dp);
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}

extern "C" int avec_relpos_attn_tc_fwd(const void* qkv, long long ld_qkv, const void* e, long long ld_e, const int* klen, int qlen, void* o,
                                       long long ld_o, float* lse, int B, int T, int H, int d, int dp, int qp_part, avec_stream_t stream) {
    AVEC_CHECK_ARG(qkv && e && o && lse && B > 0 && T > 0 && H > 0 && d > 0 && dp >= d && dp % 64 == 0 && dp <= 256 && (qp_part == 0 || qp_part == 3));
    const int nparts = qp_part ? 4 : 3;
    AVEC_CHECK_ARG(ld_qkv >= (long long)nparts * H * dp && ld_e >= (long long)H * dp && ld_o >= (long long)H * dp);
    AVEC_CHECK_ARG(aligned16(qkv, ld_qkv) && aligned16(e, ld_e) && aligned16(o, ld_o));
    AttnTcParams p;
    memset(&p, 0, sizeof(p));
    p.B = B; p.T = T; p.H = H; p.d = d; p.dp = dp; p.ndb = dp / 64; p.nqb = cdiv(T, 128); p.nkb = cdiv(T, 128);
    p.klen = klen; p.qlen = qlen; p.qp_part = qp_part;
    p.scale = 1.0f / sqrtf((float)d); p.scale_log2 = p.scale * LOG2E;
    p.o = reinterpret_cast<bf16*>(o); p.ld_o = ld_o; p.lse = lse;
    CUtensorMap mq, me;
    if (!make_maps(qkv, ld_qkv, e, ld_e, B, T, H, dp, nparts, &mq, &me)) return AVEC_ERR_DRIVER;
    const size_t smem = 1024 + RING_FWD + 2 * TILE + (p.nkb > 1 ? (size_t)128 * (dp + 1) * 4 : 0) + 64;
    if (smem > 227 * 1024) return AVEC_ERR_UNSUPPORTED;
    if (!attr_once(0, reinterpret_cast<const void*>(relpos_attn_tc_fwd_kernel), 227 * 1024)) return AVEC_ERR_LAUNCH;
    const long long ctas = (long long)B * H * p.nqb;
    if (ctas > 0x7fffffffLL) return AVEC_ERR_INVALID;
    avec_launch_pdl(relpos_attn_tc_fwd_kernel, dim3((unsigned)ctas), dim3(AT_THREADS), smem, as_stream(stream), false, p, mq, me);
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}

extern "C" int avec_relpos_attn_tc_bwd(const void* d_o, long long ld_do, const void* qkv, long long ld_qkv, const void* e, long long ld_e,
                                       const void* o, long long ld_o, const float* lse, const int* klen, int qlen, void* dqkv, long long ld_dqkv,
                                       float* dqkv_ws, float* de, long long ld_de, int B, int T, int H, int d, int dp, int qp_part, avec_stream_t stream) {
    AVEC_CHECK_ARG(d_o && qkv && e && o && lse && dqkv && de && B > 0 && T > 0 && H > 0 && d > 0 && dp >= d && dp % 64 == 0 && dp <= 256 && (qp_part == 0 || qp_part == 3));
    const int nparts = qp_part ? 4 : 3;
    AVEC_CHECK_ARG(ld_qkv >= (long long)nparts * H * dp && ld_dqkv >= (long long)nparts * H * dp && ld_e >= (long long)H * dp && ld_de >= (long long)H * dp && ld_de % 4 == 0);
    AVEC_CHECK_ARG(ld_do >= (long long)H * dp && ld_o >= (long long)H * dp);
    AVEC_CHECK_ARG(aligned16(qkv, ld_qkv) && aligned16(e, ld_e) && aligned16(o, ld_o) && aligned16(d_o, ld_do) && aligned16(dqkv, ld_dqkv));
    AVEC_CHECK_ARG((reinterpret_cast<uintptr_t>(de) % 16) == 0 && (T <= 128 || (dqkv_ws && (reinterpret_cast<uintptr_t>(dqkv_ws) % 16) == 0 && ld_dqkv % 4 == 0)));
    AttnTcParams p;
    memset(&p, 0, sizeof(p));
    p.B = B; p.T = T; p.H = H; p.d = d; p.dp = dp; p.ndb = dp / 64; p.nqb = cdiv(T, 128); p.nkb = cdiv(T, 128);
    p.klen = klen; p.qlen = qlen; p.qp_part = qp_part;
    p.scale = 1.0f / sqrtf((float)d); p.scale_log2 = p.scale * LOG2E;
    p.lse = const_cast<float*>(lse);
    p.d_o = reinterpret_cast<const bf16*>(d_o); p.ld_do = ld_do;
    p.o_in = reinterpret_cast<const bf16*>(o); p.ld_oin = ld_o;
    p.dqkv = reinterpret_cast<bf16*>(dqkv); p.ld_dqkv = ld_dqkv; p.dqkv_ws = dqkv_ws;
    p.de = de; p.ld_de = ld_de;
    CUtensorMap mq, me, mdo;
    if (!make_maps(qkv, ld_qkv, e, ld_e, B, T, H, dp, nparts, &mq, &me)) return AVEC_ERR_DRIVER;
    {
        const cuuint64_t dd[3] = {(cuuint64_t)(H * dp), (cuuint64_t)T, (cuuint64_t)B};
        const cuuint64_t sd[2] = {(cuuint64_t)ld_do * 2, (cuuint64_t)ld_do * 2 * T};
        const cuuint32_t bd[3] = {64, 128, 1};
        if (!encode_map(&mdo, d_o, 3, dd, sd, bd)) return AVEC_ERR_DRIVER;
    }
    const size_t smem = 1024 + RING_BWD + 8 * TILE + 64;
    if (!attr_once(1, reinterpret_cast<const void*>(relpos_attn_tc_bwd_kernel), 227 * 1024)) return AVEC_ERR_LAUNCH;
    const long long ctas = (long long)B * H * p.nqb;
    if (ctas > 0x7fffffffLL) return AVEC_ERR_INVALID;
    avec_launch_pdl(relpos_attn_tc_bwd_kernel, dim3((unsigned)ctas), dim3(AT_THREADS), smem, as_stream(stream), false, p, mq, me, mdo);
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}
