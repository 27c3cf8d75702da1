// Tensor-core relative-position attention for the shapes of the Efficient Conformer encoders (T <= 128 tokens per head in
// the forward, <= 112 in the backward; d <= 96; bf16; plain layout G = 1 without the Transformer-XL u / v biases).  Same
// semantics as the SIMT kernels in attention.cu (reference nnet/attentions.py:258-323):
//   S[i,j] = (q_i.k_j + q_i.e_{T-1+j-i}) / sqrt(d) + (masked ? -1e9 : 0),  P = softmax_j S,  o_i = sum_j P_ij v_j
// One CTA per (batch item, head), 8 warps, each warp owns 16 query rows.  All contractions run on warp-level
// mma.sync.m16n8k16 (bf16 in, fp32 accumulate) fed by ldmatrix from padded (conflict-free) shared-memory tiles: the problem
// is a 101 x 101 x 64 tile per head - latency / occupancy bound, far below one tcgen05 tile of work - so the TMEM path of
// the GEMM kernel would buy nothing here, while the register-resident accumulators make the softmax and the rel_to_abs
// skew cheap:  the position scores R = Q E^T are computed only on the (T + 15)-wide window a warp's 16 rows touch and are
// re-read through shared memory at the per-row shift (r = T-1+j-i), never materialised as a (T, 2T-1) tensor in HBM.
#include "common.cuh"

namespace {

constexpr int AM_THREADS = 256;
constexpr int AM_WARPS = 8;
constexpr int E_PAD = 16;   // zero rows in front of the staged E table (window starts may be negative for padded query rows)

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm_x4(uint32_t a, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t a, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&t);
}

// A fragment (16 rows x 16 k) of a row-major [m][k] tile
__device__ __forceinline__ void load_a(uint32_t (&a)[4], const bf16* tile, int ld, int m0, int k0, int lane) {
    ldsm_x4(s_u32(tile + (size_t)(m0 + (lane & 15)) * ld + k0 + (lane >> 4) * 8), a[0], a[1], a[2], a[3]);
}
// A fragment of A[m][k] = tile[k][m] (tile stored [k][m])
__device__ __forceinline__ void load_a_t(uint32_t (&a)[4], const bf16* tile, int ld, int m0, int k0, int lane) {
    ldsm_x4_t(s_u32(tile + (size_t)(k0 + (lane & 7) + ((lane >> 4) & 1) * 8) * ld + m0 + ((lane >> 3) & 1) * 8), a[0], a[1], a[2], a[3]);
}
// B fragments of two adjacent 8-wide n tiles from a tile stored [n][k]: (b[0], b[1]) = tile n0, (b[2], b[3]) = tile n0 + 8
__device__ __forceinline__ void load_b_nk(uint32_t (&b)[4], const bf16* tile, int ld, int n0, int k0, int lane) {
    ldsm_x4(s_u32(tile + (size_t)(n0 + (lane & 7) + ((lane >> 4) & 1) * 8) * ld + k0 + ((lane >> 3) & 1) * 8), b[0], b[1], b[2], b[3]);
}
// same from a tile stored [k][n]
__device__ __forceinline__ void load_b_kn(uint32_t (&b)[4], const bf16* tile, int ld, int n0, int k0, int lane) {
    ldsm_x4_t(s_u32(tile + (size_t)(k0 + (lane & 7) + ((lane >> 3) & 1) * 8) * ld + n0 + ((lane >> 4) & 1) * 8), b[0], b[1], b[2], b[3]);
}

// stage `rows` x d elements (row r at src + r * src_ld) as bf16 [rows_pad][DS] with zero padding of columns d..DP and rows;
// 16-byte / 4-byte / 2-byte global loads depending on what the head offset h * d allows (d = 64 / 90 / 45)
__device__ __forceinline__ void stage_rows(bf16* dst, int DS, int DP, const bf16* __restrict__ src, long long src_ld, int rows, int rows_pad, int d,
                                           int dst_row0, int tid) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(src);
    if (d % 8 == 0 && (a & 15) == 0 && (src_ld * 2) % 16 == 0) {
        const int cpr = DP / 8;
        for (int idx = tid; idx < rows_pad * cpr; idx += AM_THREADS) {
            const int r = idx / cpr, c = (idx - r * cpr) * 8;
            uint4 w = make_uint4(0u, 0u, 0u, 0u);
            if (r < rows && c < d) w = __ldg(reinterpret_cast<const uint4*>(src + (long long)r * src_ld + c));
            *reinterpret_cast<uint4*>(dst + (size_t)(dst_row0 + r) * DS + c) = w;
        }
    } else if (d % 2 == 0 && (a & 3) == 0 && (src_ld * 2) % 4 == 0) {
        const int cpr = DP / 2;
        for (int idx = tid; idx < rows_pad * cpr; idx += AM_THREADS) {
            const int r = idx / cpr, c = (idx - r * cpr) * 2;
            uint32_t w = 0u;
            if (r < rows && c < d) w = __ldg(reinterpret_cast<const uint32_t*>(src + (long long)r * src_ld + c));
            *reinterpret_cast<uint32_t*>(dst + (size_t)(dst_row0 + r) * DS + c) = w;
        }
    } else {
        const int cpr = DP / 2;   // bf16 pairs per row
        for (int idx = tid; idx < rows_pad * cpr; idx += AM_THREADS) {
            const int r = idx / cpr, c = (idx - r * cpr) * 2;
            uint32_t w = 0u;
            if (r < rows) {
                const unsigned short* s = reinterpret_cast<const unsigned short*>(src + (long long)r * src_ld);
                if (c < d) w = (uint32_t)__ldg(s + c);
                if (c + 1 < d) w |= (uint32_t)__ldg(s + c + 1) << 16;
            }
            *reinterpret_cast<uint32_t*>(dst + (size_t)(dst_row0 + r) * DS + c) = w;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------ forward
// NT = key tiles of 8 (TP = 8 NT padded tokens, NT even), DP = padded head width (multiple of 16)
template <int NT, int DP>
__global__ void __launch_bounds__(AM_THREADS, 1) relpos_attn_mma_fwd_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ e,
                                                                          const int* __restrict__ klen, int qlen, bf16* __restrict__ o,
                                                                          float* __restrict__ probs, int T, int H, int d) {
    constexpr int TP = NT * 8, DS = DP + 8, KT = DP / 16;
    constexpr int NW = NT + 2;              // window tiles: covers T + 15 <= TP + 15 < 8 (NT + 2)
    constexpr int RW = NW * 8 + 2;          // fp32 row stride of the per-warp R window (even: 8-byte stores)
    constexpr int EP = 2 * TP + 2 * E_PAD;  // staged E rows: E_PAD zeros, 2T-1 rows, zeros
    extern __shared__ __align__(16) uint8_t sm_raw[];
    bf16* Qs = reinterpret_cast<bf16*>(sm_raw);
    bf16* Ks = Qs + TP * DS;
    bf16* Vs = Ks + TP * DS;
    bf16* Es = Vs + TP * DS;
    float* Rs = reinterpret_cast<float*>(Es + EP * DS);   // [AM_WARPS][16][RW]
    const int b = blockIdx.x / H, h = blockIdx.x % H;
    const int D = H * d;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bf16* qkv_b = qkv + (size_t)b * T * 3 * D + h * d;
    stage_rows(Qs, DS, DP, qkv_b, 3LL * D, T, TP, d, 0, tid);
    stage_rows(Ks, DS, DP, qkv_b + D, 3LL * D, T, TP, d, 0, tid);
    stage_rows(Vs, DS, DP, qkv_b + 2 * D, 3LL * D, T, TP, d, 0, tid);
    for (int idx = tid; idx < E_PAD * DS / 2; idx += AM_THREADS) reinterpret_cast<uint32_t*>(Es)[idx] = 0u;
    stage_rows(Es, DS, DP, e + h * d, (long long)D, 2 * T - 1, EP - E_PAD, d, E_PAD, tid);
    __syncthreads();
    const int i0 = warp * 16;
    if (i0 >= T) return;
    const int g = lane >> 2, q = lane & 3;
    const int kl = klen ? klen[b] : T;
    const float scale = rsqrtf((float)d);

    uint32_t qa[KT][4];
#pragma unroll
    for (int kt = 0; kt < KT; ++kt) load_a(qa[kt], Qs, DS, i0, kt * 16, lane);

    // ---- position scores on this warp's window: R[il][w] = q_{i0+il} . e_{r0+w}, r0 = T-1-i0-15
    float* Rw = Rs + (size_t)warp * 16 * RW;
    const int erow0 = T - 1 - i0 - 15 + E_PAD;   // >= 1
#pragma unroll 1
    for (int wt = 0; wt < NW; wt += 2) {
        float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int kt = 0; kt < KT; ++kt) {
            uint32_t bb[4];
            load_b_nk(bb, Es, DS, erow0 + wt * 8, kt * 16, lane);
            mma16816(c0, qa[kt], bb[0], bb[1]);
            mma16816(c1, qa[kt], bb[2], bb[3]);
        }
        *reinterpret_cast<float2*>(Rw + (size_t)g * RW + wt * 8 + 2 * q) = make_float2(c0[0], c0[1]);
        *reinterpret_cast<float2*>(Rw + (size_t)(g + 8) * RW + wt * 8 + 2 * q) = make_float2(c0[2], c0[3]);
        *reinterpret_cast<float2*>(Rw + (size_t)g * RW + wt * 8 + 8 + 2 * q) = make_float2(c1[0], c1[1]);
        *reinterpret_cast<float2*>(Rw + (size_t)(g + 8) * RW + wt * 8 + 8 + 2 * q) = make_float2(c1[2], c1[3]);
    }
    __syncwarp();

    // ---- content scores, skewed position term, mask, softmax (rows g and g + 8 of this warp's block)
    float s[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) { s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f; }
#pragma unroll
    for (int nt = 0; nt < NT; nt += 2) {
#pragma unroll
        for (int kt = 0; kt < KT; ++kt) {
            uint32_t bb[4];
            load_b_nk(bb, Ks, DS, nt * 8, kt * 16, lane);
            mma16816(s[nt], qa[kt], bb[0], bb[1]);
            mma16816(s[nt + 1], qa[kt], bb[2], bb[3]);
        }
    }
    const int ia = i0 + g, ib = i0 + g + 8;
    const bool qa_masked = ia >= qlen, qb_masked = ib >= qlen;
    float mxa = -INFINITY, mxb = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int j = nt * 8 + 2 * q + u;
            // window column of (il, j): j - il + 15
            float va = (s[nt][u] + Rw[(size_t)g * RW + j - g + 15]) * scale;
            float vb = (s[nt][2 + u] + Rw[(size_t)(g + 8) * RW + j - g + 7]) * scale;
            if (j >= kl || qa_masked) va += -1e9f;
            if (j >= kl || qb_masked) vb += -1e9f;
            if (j >= T) { va = -INFINITY; vb = -INFINITY; }
            s[nt][u] = va; s[nt][2 + u] = vb;
            mxa = fmaxf(mxa, va); mxb = fmaxf(mxb, vb);
        }
    }
    mxa = fmaxf(mxa, __shfl_xor_sync(0xffffffffu, mxa, 1)); mxa = fmaxf(mxa, __shfl_xor_sync(0xffffffffu, mxa, 2));
    mxb = fmaxf(mxb, __shfl_xor_sync(0xffffffffu, mxb, 1)); mxb = fmaxf(mxb, __shfl_xor_sync(0xffffffffu, mxb, 2));
    float sa = 0.f, sb = 0.f;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            s[nt][u] = __expf(s[nt][u] - mxa); sa += s[nt][u];
            s[nt][2 + u] = __expf(s[nt][2 + u] - mxb); sb += s[nt][2 + u];
        }
    }
    sa += __shfl_xor_sync(0xffffffffu, sa, 1); sa += __shfl_xor_sync(0xffffffffu, sa, 2);
    sb += __shfl_xor_sync(0xffffffffu, sb, 1); sb += __shfl_xor_sync(0xffffffffu, sb, 2);
    const float inva = 1.0f / sa, invb = 1.0f / sb;
    float* pra = probs + (((size_t)b * H + h) * T + ia) * T;
    float* prb = probs + (((size_t)b * H + h) * T + ib) * T;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int j = nt * 8 + 2 * q + u;
            s[nt][u] *= inva; s[nt][2 + u] *= invb;
            if (j < T) {
                if (ia < T) pra[j] = s[nt][u];
                if (ib < T) prb[j] = s[nt][2 + u];
            }
        }
    }

    // ---- O = P V   (P straight from the accumulator registers: two adjacent score tiles form one A fragment)
    float oacc[DP / 8][4];
#pragma unroll
    for (int nc = 0; nc < DP / 8; ++nc) { oacc[nc][0] = oacc[nc][1] = oacc[nc][2] = oacc[nc][3] = 0.f; }
#pragma unroll
    for (int kt = 0; kt < NT / 2; ++kt) {
        uint32_t pa[4] = {pack2(s[2 * kt][0], s[2 * kt][1]), pack2(s[2 * kt][2], s[2 * kt][3]),
                          pack2(s[2 * kt + 1][0], s[2 * kt + 1][1]), pack2(s[2 * kt + 1][2], s[2 * kt + 1][3])};
#pragma unroll
        for (int nc = 0; nc < DP / 8; nc += 2) {
            uint32_t bb[4];
            load_b_kn(bb, Vs, DS, nc * 8, kt * 16, lane);
            mma16816(oacc[nc], pa, bb[0], bb[1]);
            mma16816(oacc[nc + 1], pa, bb[2], bb[3]);
        }
    }
    bf16* oa = o + ((size_t)b * T + ia) * D + h * d;
    bf16* ob = o + ((size_t)b * T + ib) * D + h * d;
#pragma unroll
    for (int nc = 0; nc < DP / 8; ++nc) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int c = nc * 8 + 2 * q + u;
            if (c < d) {
                if (ia < T) oa[c] = __float2bfloat16_rn(oacc[nc][u]);
                if (ib < T) ob[c] = __float2bfloat16_rn(oacc[nc][2 + u]);
            }
        }
    }
}

template <int NT, int DP>
constexpr size_t fwd_smem_bytes() {
    return (size_t)(3 * NT * 8 + 2 * NT * 8 + 2 * E_PAD) * (DP + 8) * 2 + (size_t)AM_WARPS * 16 * ((NT + 2) * 8 + 2) * 4;
}

// ----------------------------------------------------------------------------------------------------------- backward
// Phase A (warp = 16 query rows): dP = dO V^T, delta, dS = P (dP - delta) / sqrt(d); dQ = dS K + skew(dS) E; P, dS and the
// skewed dS are left in shared memory as bf16.  Phase B (warp = 16 keys): dV = P^T dO, dK = dS^T Q.  Phase C (warp = 16
// relative offsets): dE = skew(dS)^T Q, accumulated over the batch with fp32 atomics.
template <int NT, int DP>
__global__ void __launch_bounds__(AM_THREADS, 1) relpos_attn_mma_bwd_kernel(const bf16* __restrict__ d_o, const bf16* __restrict__ qkv,
                                                                          const bf16* __restrict__ e, const float* __restrict__ probs,
                                                                          bf16* __restrict__ dqkv, float* __restrict__ de, int T, int H, int d) {
    constexpr int TP = NT * 8, DS = DP + 8, KT = DP / 16;
    constexpr int NW = NT + 2;
    constexpr int EP = 2 * TP + 2 * E_PAD;
    constexpr int PS = TP + 8;              // row stride of the P / dS tiles
    constexpr int SK = 2 * TP + 2 * E_PAD + 8;  // row stride of the skewed dS tile: column r + E_PAD
    extern __shared__ __align__(16) uint8_t sm_raw[];
    bf16* Qs = reinterpret_cast<bf16*>(sm_raw);
    bf16* Ks = Qs + TP * DS;
    bf16* Vs = Ks + TP * DS;
    bf16* Os = Vs + TP * DS;
    bf16* Es = Os + TP * DS;
    bf16* Pb = Es + EP * DS;       // [TP][PS]
    bf16* Sb = Pb + TP * PS;       // [TP][PS]   dS
    bf16* Sk = Sb + TP * PS;       // [TP][SK]   dS at column T-1+j-i + E_PAD
    const int b = blockIdx.x / H, h = blockIdx.x % H;
    const int D = H * d;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bf16* qkv_b = qkv + (size_t)b * T * 3 * D + h * d;
    stage_rows(Qs, DS, DP, qkv_b, 3LL * D, T, TP, d, 0, tid);
    stage_rows(Ks, DS, DP, qkv_b + D, 3LL * D, T, TP, d, 0, tid);
    stage_rows(Vs, DS, DP, qkv_b + 2 * D, 3LL * D, T, TP, d, 0, tid);
    stage_rows(Os, DS, DP, d_o + (size_t)b * T * D + h * d, (long long)D, T, TP, d, 0, tid);
    for (int idx = tid; idx < E_PAD * DS / 2; idx += AM_THREADS) reinterpret_cast<uint32_t*>(Es)[idx] = 0u;
    stage_rows(Es, DS, DP, e + h * d, (long long)D, 2 * T - 1, EP - E_PAD, d, E_PAD, tid);
    for (int idx = tid; idx < TP * (2 * PS + SK) / 2; idx += AM_THREADS) reinterpret_cast<uint32_t*>(Pb)[idx] = 0u;
    __syncthreads();
    const int g = lane >> 2, q = lane & 3;
    const float scale = rsqrtf((float)d);
    bf16* dq_b = dqkv + (size_t)b * T * 3 * D + h * d;

    // ------------------------------------------------ phase A
    const int i0 = warp * 16;
    if (i0 < T) {
        const int ia = i0 + g, ib = i0 + g + 8;
        float s[NT][4];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) { s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f; }
        {
            uint32_t oa[KT][4];
#pragma unroll
            for (int kt = 0; kt < KT; ++kt) load_a(oa[kt], Os, DS, i0, kt * 16, lane);
#pragma unroll
            for (int nt = 0; nt < NT; nt += 2) {
#pragma unroll
                for (int kt = 0; kt < KT; ++kt) {
                    uint32_t bb[4];
                    load_b_nk(bb, Vs, DS, nt * 8, kt * 16, lane);
                    mma16816(s[nt], oa[kt], bb[0], bb[1]);
                    mma16816(s[nt + 1], oa[kt], bb[2], bb[3]);
                }
            }
        }
        // probabilities of rows ia / ib, delta, dS (in place)
        const float* pra = probs + (((size_t)b * H + h) * T + ia) * T;
        const float* prb = probs + (((size_t)b * H + h) * T + ib) * T;
        float da = 0.f, db = 0.f;
        float pv[NT][4];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int j = nt * 8 + 2 * q + u;
                const float pa_ = (j < T && ia < T) ? __ldg(pra + j) : 0.f;
                const float pb_ = (j < T && ib < T) ? __ldg(prb + j) : 0.f;
                pv[nt][u] = pa_; pv[nt][2 + u] = pb_;
                da = fmaf(pa_, s[nt][u], da); db = fmaf(pb_, s[nt][2 + u], db);
            }
        }
        da += __shfl_xor_sync(0xffffffffu, da, 1); da += __shfl_xor_sync(0xffffffffu, da, 2);
        db += __shfl_xor_sync(0xffffffffu, db, 1); db += __shfl_xor_sync(0xffffffffu, db, 2);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const int j = nt * 8 + 2 * q;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                s[nt][u] = pv[nt][u] * (s[nt][u] - da) * scale;
                s[nt][2 + u] = pv[nt][2 + u] * (s[nt][2 + u] - db) * scale;
            }
            *reinterpret_cast<uint32_t*>(Pb + (size_t)ia * PS + j) = pack2(pv[nt][0], pv[nt][1]);
            *reinterpret_cast<uint32_t*>(Pb + (size_t)ib * PS + j) = pack2(pv[nt][2], pv[nt][3]);
            *reinterpret_cast<uint32_t*>(Sb + (size_t)ia * PS + j) = pack2(s[nt][0], s[nt][1]);
            *reinterpret_cast<uint32_t*>(Sb + (size_t)ib * PS + j) = pack2(s[nt][2], s[nt][3]);
            // skewed copy: column T-1+j-i + E_PAD (parity differs per row: element stores)
            if (j < T) {
                Sk[(size_t)ia * SK + T - 1 + j - ia + E_PAD] = __float2bfloat16_rn(s[nt][0]);
                Sk[(size_t)ib * SK + T - 1 + j - ib + E_PAD] = __float2bfloat16_rn(s[nt][2]);
            }
            if (j + 1 < T) {
                Sk[(size_t)ia * SK + T + j - ia + E_PAD] = __float2bfloat16_rn(s[nt][1]);
                Sk[(size_t)ib * SK + T + j - ib + E_PAD] = __float2bfloat16_rn(s[nt][3]);
            }
        }
        __syncwarp();
        // dQ = dS K  +  skew(dS) E
        float acc[DP / 8][4];
#pragma unroll
        for (int nc = 0; nc < DP / 8; ++nc) { acc[nc][0] = acc[nc][1] = acc[nc][2] = acc[nc][3] = 0.f; }
#pragma unroll
        for (int kt = 0; kt < NT / 2; ++kt) {
            uint32_t sa_[4] = {pack2(s[2 * kt][0], s[2 * kt][1]), pack2(s[2 * kt][2], s[2 * kt][3]),
                               pack2(s[2 * kt + 1][0], s[2 * kt + 1][1]), pack2(s[2 * kt + 1][2], s[2 * kt + 1][3])};
#pragma unroll
            for (int nc = 0; nc < DP / 8; nc += 2) {
                uint32_t bb[4];
                load_b_kn(bb, Ks, DS, nc * 8, kt * 16, lane);
                mma16816(acc[nc], sa_, bb[0], bb[1]);
                mma16816(acc[nc + 1], sa_, bb[2], bb[3]);
            }
        }
        // window of this warp's rows in the skewed tile: columns [T-1-i0-15, ...) + E_PAD, NW tiles of 8 (= NW / 2 k steps)
        const int w0 = T - 1 - i0 - 15 + E_PAD;   // >= 1; ldmatrix rows need 16-byte alignment -> round down to a multiple of 8
        const int w0a = w0 & ~7;
#pragma unroll 1
        for (int kt = 0; kt < (NW + 2) / 2; ++kt) {
            uint32_t sa_[4];
            load_a(sa_, Sk, SK, i0, w0a + kt * 16, lane);
#pragma unroll
            for (int nc = 0; nc < DP / 8; nc += 2) {
                uint32_t bb[4];
                load_b_kn(bb, Es, DS, nc * 8, w0a + kt * 16, lane);
                mma16816(acc[nc], sa_, bb[0], bb[1]);
                mma16816(acc[nc + 1], sa_, bb[2], bb[3]);
            }
        }
#pragma unroll
        for (int nc = 0; nc < DP / 8; ++nc) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int c = nc * 8 + 2 * q + u;
                if (c < d) {
                    if (ia < T) dq_b[(size_t)ia * 3 * D + c] = __float2bfloat16_rn(acc[nc][u]);
                    if (ib < T) dq_b[(size_t)ib * 3 * D + c] = __float2bfloat16_rn(acc[nc][2 + u]);
                }
            }
        }
    }
    __syncthreads();

    // ------------------------------------------------ phase B: key blocks
    for (int jb = warp * 16; jb < T; jb += AM_WARPS * 16) {
        float av[DP / 8][4], ak[DP / 8][4];
#pragma unroll
        for (int nc = 0; nc < DP / 8; ++nc) { av[nc][0] = av[nc][1] = av[nc][2] = av[nc][3] = 0.f; ak[nc][0] = ak[nc][1] = ak[nc][2] = ak[nc][3] = 0.f; }
#pragma unroll 1
        for (int kt = 0; kt < NT / 2; ++kt) {
            uint32_t pa[4], sa_[4];
            load_a_t(pa, Pb, PS, jb, kt * 16, lane);
            load_a_t(sa_, Sb, PS, jb, kt * 16, lane);
#pragma unroll
            for (int nc = 0; nc < DP / 8; nc += 2) {
                uint32_t bo[4], bq[4];
                load_b_kn(bo, Os, DS, nc * 8, kt * 16, lane);
                load_b_kn(bq, Qs, DS, nc * 8, kt * 16, lane);
                mma16816(av[nc], pa, bo[0], bo[1]);
                mma16816(av[nc + 1], pa, bo[2], bo[3]);
                mma16816(ak[nc], sa_, bq[0], bq[1]);
                mma16816(ak[nc + 1], sa_, bq[2], bq[3]);
            }
        }
        const int ja = jb + g, jb2 = jb + g + 8;
#pragma unroll
        for (int nc = 0; nc < DP / 8; ++nc) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int c = nc * 8 + 2 * q + u;
                if (c < d) {
                    if (ja < T) { dq_b[(size_t)ja * 3 * D + 2 * D + c] = __float2bfloat16_rn(av[nc][u]); dq_b[(size_t)ja * 3 * D + D + c] = __float2bfloat16_rn(ak[nc][u]); }
                    if (jb2 < T) { dq_b[(size_t)jb2 * 3 * D + 2 * D + c] = __float2bfloat16_rn(av[nc][2 + u]); dq_b[(size_t)jb2 * 3 * D + D + c] = __float2bfloat16_rn(ak[nc][2 + u]); }
                }
            }
        }
    }

    // ------------------------------------------------ phase C: dE_r = sum_i Sk[i][r + E_PAD] Q_i   (blocks of 16 offsets r)
    for (int rb = warp * 16; rb < 2 * T - 1; rb += AM_WARPS * 16) {
        float acc[DP / 8][4];
#pragma unroll
        for (int nc = 0; nc < DP / 8; ++nc) { acc[nc][0] = acc[nc][1] = acc[nc][2] = acc[nc][3] = 0.f; }
        // rows i with a non-zero entry in columns [rb, rb + 16): T-1-(rb+15) <= i <= 2T-2-rb
        const int ilo = max(0, T - 1 - rb - 15) & ~15, ihi = min(T - 1, 2 * T - 2 - rb);
#pragma unroll 1
        for (int k0 = ilo; k0 <= ihi; k0 += 16) {
            uint32_t sa_[4];
            load_a_t(sa_, Sk, SK, rb + E_PAD, k0, lane);
#pragma unroll
            for (int nc = 0; nc < DP / 8; nc += 2) {
                uint32_t bq[4];
                load_b_kn(bq, Qs, DS, nc * 8, k0, lane);
                mma16816(acc[nc], sa_, bq[0], bq[1]);
                mma16816(acc[nc + 1], sa_, bq[2], bq[3]);
            }
        }
        const int ra = rb + g, rb2 = rb + g + 8;
#pragma unroll
        for (int nc = 0; nc < DP / 8; ++nc) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int c = nc * 8 + 2 * q + u;
                if (c < d) {
                    if (ra < 2 * T - 1) atomicAdd(de + (size_t)ra * D + h * d + c, acc[nc][u]);
                    if (rb2 < 2 * T - 1) atomicAdd(de + (size_t)rb2 * D + h * d + c, acc[nc][2 + u]);
                }
            }
        }
    }
}

template <int NT, int DP>
constexpr size_t bwd_smem_bytes() {
    return (size_t)(4 * NT * 8 + 2 * NT * 8 + 2 * E_PAD) * (DP + 8) * 2 + (size_t)NT * 8 * (2 * (NT * 8 + 8) + 2 * NT * 8 + 2 * E_PAD + 8) * 2;
}

int g_attn_mma = -1;
bool mma_enabled() {
    if (g_attn_mma < 0) { const char* ev = getenv("AVEC_ATTN_MMA"); g_attn_mma = ev ? atoi(ev) : 1; }
    return g_attn_mma != 0;
}

template <int NT, int DP>
int launch_fwd(const bf16* qkv, const bf16* e, const int* klen, int qlen, bf16* o, float* probs, int B, int T, int H, int d, cudaStream_t st) {
    auto kfn = relpos_attn_mma_fwd_kernel<NT, DP>;
    constexpr size_t smem = fwd_smem_bytes<NT, DP>();
    static_assert(smem <= 227 * 1024, "forward tile does not fit");
    if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return AVEC_ERR_LAUNCH;
    kfn<<<B * H, AM_THREADS, smem, st>>>(qkv, e, klen, qlen, o, probs, T, H, d);
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}
template <int NT, int DP>
int launch_bwd(const bf16* d_o, const bf16* qkv, const bf16* e, const float* probs, bf16* dqkv, float* de, int B, int T, int H, int d, cudaStream_t st) {
    auto kfn = relpos_attn_mma_bwd_kernel<NT, DP>;
    constexpr size_t smem = bwd_smem_bytes<NT, DP>();
    static_assert(smem <= 227 * 1024, "backward tile does not fit");
    if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return AVEC_ERR_LAUNCH;
    kfn<<<B * H, AM_THREADS, smem, st>>>(d_o, qkv, e, probs, dqkv, de, T, H, d);
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}

}  // namespace

// returns AVEC_ERR_UNSUPPORTED when the shape is outside the tensor-core kernel's envelope (the caller then runs the SIMT kernel)
int avec_attn_mma_fwd(const void* qkv, const void* e, const int* klen, int qlen, void* o, float* probs, int B, int T, int H, int d, cudaStream_t st) {
    if (!mma_enabled() || T > 128 || d > 96 || T < 1) return AVEC_ERR_UNSUPPORTED;
    const int dp = d <= 48 ? 48 : (d <= 64 ? 64 : 96);
    const int nt = T <= 64 ? 8 : (T <= 80 ? 10 : (T <= 112 ? 14 : 16));
#define AVEC_ATT_FWD(NT_, DP_) if (nt == NT_ && dp == DP_) return launch_fwd<NT_, DP_>((const bf16*)qkv, (const bf16*)e, klen, qlen, (bf16*)o, probs, B, T, H, d, st)
    AVEC_ATT_FWD(8, 48); AVEC_ATT_FWD(8, 64); AVEC_ATT_FWD(8, 96);
    AVEC_ATT_FWD(10, 48); AVEC_ATT_FWD(10, 64); AVEC_ATT_FWD(10, 96);
    AVEC_ATT_FWD(14, 48); AVEC_ATT_FWD(14, 64); AVEC_ATT_FWD(14, 96);
    AVEC_ATT_FWD(16, 48); AVEC_ATT_FWD(16, 64);
#undef AVEC_ATT_FWD
    return AVEC_ERR_UNSUPPORTED;
}

int avec_attn_mma_bwd(const void* d_o, const void* qkv, const void* e, const float* probs, void* dqkv, float* de, int B, int T, int H, int d,
                      cudaStream_t st) {
    if (!mma_enabled() || T > 112 || d > 96 || T < 1) return AVEC_ERR_UNSUPPORTED;
    const int dp = d <= 48 ? 48 : (d <= 64 ? 64 : 96);
    const int nt = T <= 64 ? 8 : (T <= 80 ? 10 : 14);
#define AVEC_ATT_BWD(NT_, DP_) if (nt == NT_ && dp == DP_) return launch_bwd<NT_, DP_>((const bf16*)d_o, (const bf16*)qkv, (const bf16*)e, probs, (bf16*)dqkv, de, B, T, H, d, st)
    AVEC_ATT_BWD(8, 48); AVEC_ATT_BWD(8, 64); AVEC_ATT_BWD(8, 96);
    AVEC_ATT_BWD(10, 48); AVEC_ATT_BWD(10, 64); AVEC_ATT_BWD(10, 96);
    AVEC_ATT_BWD(14, 48); AVEC_ATT_BWD(14, 64);
#undef AVEC_ATT_BWD
    return AVEC_ERR_UNSUPPORTED;
}
