// tcgen05 / TMEM GEMM and implicit-GEMM convolution for sm_100a (bf16 operands, fp32 accumulate in tensor memory).
//
// One CTA computes one 128 x BN output tile (BN = multiple of 16, <= 256, chosen per problem so that awkward widths
// such as 180 / 360 / 720 / 1080 waste few columns).  Warp roles:
//   warps 0-3  producers: gather 128-byte row slices (64 bf16) of both operands from HBM/L2 into the canonical
//              128B-swizzled shared-memory layout UMMA expects (K-major: smem row = matrix row, MN-major: smem row =
//              reduction index), zero-filling out-of-range rows / taps / tails; then, once the main loop is done, the
//              same four warps are the epilogue (tcgen05.ld of their TMEM lane quarter -> fused epilogue -> HBM);
//   warp 4     lane 0 issues tcgen05.mma (UMMA 128 x BN x 16, kind::f16) over a STAGES-deep mbarrier ring and commits
//              stage-free / accumulator-ready barriers; the warp also owns the TMEM allocation.
// The gather producers are what let ONE kernel serve: plain linears with any leading dimension (D = 180 rows are not
// 16-byte aligned, which rules out a TMA descriptor), transposed operands of dgrad / wgrad (MN-major descriptors, no
// transposed copies in HBM), and the ResNet / strided convolutions as implicit GEMMs (im2col never materialised).
#include "common.cuh"
#include <cuda.h>
#include <cstring>
#include <cstdint>
#include <cstdlib>
#include <algorithm>
#include <type_traits>

namespace {

constexpr int BM = 128;
constexpr int BKE = 64;                  // reduction elements per k-block (one 128-byte swizzle row)
constexpr int PRODUCER_THREADS = 128;
constexpr int TC_THREADS = 192;   // warps 0-3 gather producers + epilogue, warp 4 MMA issuer, warp 5 TMA producer
constexpr int TC_THREADS_WIDE = 320;   // + warps 6-9: second epilogue warpgroup (tiles wider than 64 columns: it takes the odd slabs)

enum OperandKind {
    OP_PLAIN_K = 0,    // element (row, k) at base[row*ld + k]               -> K-major tile
    OP_PLAIN_MN = 1,   // element (row, k) at base[k*ld + row]               -> MN-major tile
    OP_CONV_FWD = 2,   // A: row = output site, k = (tap, ci)                 -> K-major
    OP_CONV_DGRAD = 3, // A: row = input site, k = (tap, co), gather from dY  -> K-major
    OP_CONV_WGRAD_X = 4, // B: row = (tap, ci), k = output site, gather from X  -> MN-major
    OP_CONV_TAPS = 5,    // A, C == 1 (stems): row = output site, k = tap; element-wise im2col gather -> K-major
    OP_CONV_TAPS_MN = 6, // B, C == 1 (stem wgrad): row = tap, k = output site              -> MN-major
    // ---- TMA-fed kinds (cp.async.bulk.tensor issued by one thread; 16-byte aligned strides required)
    OP_TMA_K = 8,        // 2-d tensor map, box (64 k, rows)                                -> K-major
    OP_TMA_MN = 9,       // 2-d tensor map, box (64 mn, 64 k) per 64-wide MN group          -> MN-major
    OP_TMA_CONV_K = 10,  // A of a stride-1 conv fwd / dgrad: 4-d map (C,W,H,N), box (64,W,BH,BI) at the tap's shift -> K-major
    OP_TMA_CONV_MN = 11, // wgrad operands: A = dY (no shift), B = X shifted by the group's tap  -> MN-major, k = site
    OP_TMA_CONV_HALO = 12 // A of a stride-1 3x3 conv on images wider than a tile: ONE halo tile ((BH+2) x (W+2) sites x 64 ch)
                          // per 64-channel block is loaded per output tile and all 9 taps read it through row-offset
                          // descriptors (the accumulator rows then live on the (W+2)-wide halo grid)  -> K-major
};
__host__ __device__ inline bool is_tma(int kind) { return kind >= OP_TMA_K; }
__host__ __device__ inline bool is_mn(int kind) {
    return kind == OP_PLAIN_MN || kind == OP_CONV_WGRAD_X || kind == OP_CONV_TAPS_MN || kind == OP_TMA_MN || kind == OP_TMA_CONV_MN;
}

struct TcParams {
    int M, N, K;
    int BN;
    int stages;
    int num_kb, kb_per_split;
    int a_kind, b_kind;
    const bf16* A; long long a_ld; int a_align;
    const bf16* B; long long b_ld; int b_align;
    ConvGeom g;
    int cpb;  // 64-channel blocks per filter tap
    EpiParams ep;
    int out_transposed;
    // TMA / tile geometry
    int ksteps;                 // UMMA K=16 steps per k-block (4 for 64-element k-blocks)
    int a_rows, b_rows;         // smem rows (128 B each) per stage of each operand
    int a_group_stride, b_group_stride;  // bytes between 64-wide MN groups (MN-major LBO)
    int a_tx, b_tx;             // bytes the TMA loads of one k-block deliver per operand
    int conv_tiles;             // 1: the M (or reduction) tiles are row-aligned image tiles of BH rows x BI images
    int ct_BH, ct_BI, ct_tph;   // tile rows, images per tile, tiles per image (on the conv's OUTPUT grid ct_H x ct_W)
    int ct_H, ct_W, ct_s;       // output grid and spatial stride (1 or 2; strided taps use the tensor map's element strides)
    int ct_dgrad;               // tap shift sign (dgrad reads dy[p + pad - tap])
    int halo_bytes;             // bytes of one halo tile (multiple of 1024); 0 = no halo mode
    int halo_W2;                // W + 2
    int stg_dedicated;          // 1: epilogue staging has its own shared memory (persistent launches); 0: it aliases the drained ring
    // parity-class mode of strided dgrads: filter taps come from a table (tile shift + weight k offset per tap) and output rows
    // are scattered: row = tile base + rowrel[r] (table in shared memory), tile base = ((n0 * rm_Hi + rm_sy * h0 + rm_a) * rm_Wi + rm_b)
    int gen_taps;               // number of table taps (0 = regular 3x3 / 1x1 addressing)
    signed char tap_dh[9], tap_dw[9];
    int tap_koff[9];            // element offset of the tap's 64-channel blocks inside a B row
    int rm_on, rm_Hi, rm_Wi, rm_sy, rm_sx, rm_a, rm_b;
    int epi_fast;               // 0: generic epilogue, 1: fast epilogue with 8-byte aligned rows, 2: 16-byte aligned rows
    int grid_m, grid_n, splits; // tile grid (the launch grid is min(#tiles, resident CTAs): persistent tile loop)
    unsigned long long* dbg_ts; // diagnostics (avec_set_debug_timestamps): CTA (0,0,0) records globaltimer at phase boundaries
    int dbg_mode;               // diagnostics (AVEC_DEBUG_MODE bits): 1 no TMA loads, 2 no MMAs, 4 sleeping epilogue wait, 8 one-lane MMA poll
    int dbg_rowofs;             // diagnostics (AVEC_DEBUG_ROWOFS): A tile loaded `ofs` rows early, descriptor started `ofs` rows in
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// one elected lane of a converged warp (elect.sync): code under it may feed warp-uniform operands straight to the uniform
// datapath (UTCHMMA / UTMALDG) - a plain `lane == 0` branch makes ptxas wrap every such instruction in an
// ELECT / R2UR.BROADCAST waterfall loop (~160 cycles per MMA)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
    while (!mbar_try(bar, parity)) __nanosleep(100);
}
__device__ __forceinline__ void mbar_wait_addr(uint32_t bar_saddr, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP_A:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE_A;\n\t"
        "bra WAIT_LOOP_A;\n\t"
        "WAIT_DONE_A:\n\t}" ::"r"(bar_saddr), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_commit_addr(uint32_t bar_saddr) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_saddr) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Timeline / ablation diagnostics are compiled only with -DAVEC_TIMELINE (tools/ts_probe.py needs such a build): in the
// production build the per-k-block loops of the single-thread MMA issuer and TMA producer carry no trace of them (those
// loops are instruction-latency bound: every scalar instruction per k-block costs ~4-5 cycles of issue time).
#ifdef AVEC_TIMELINE
#define AVEC_TS(slot) do { if (p.dbg_ts && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) p.dbg_ts[slot] = gtime(); } while (0)
// per-tile timeline of CTA 0 (tiles j < 32): slots 8 + 5 j + {0 epilogue starts waiting, 1 accumulator ready, 2 epilogue done,
// 3 first k-block in shared memory, 4 last MMA issued}; the debug buffer holds 8 + 5 * 32 = 168 values
// k-block level trace of tile 6 of CTA 0 in SM clocks: [168 + 2 i + {0 full-wait done, 1 MMAs + commit issued}] (MMA thread),
// [200 + 2 i + {0 empty-wait done, 1 TMA issued}] (producer thread), i < 16; buffer holds 232 values
#define AVEC_TSK(j, i, base, k) do { if (p.dbg_ts && blockIdx.x == 0 && (j) == 6 && (i) < 16) p.dbg_ts[(base) + 2 * (i) + (k)] = (unsigned long long)clock64(); } while (0)
#define AVEC_TSJ(j, k) do { if (p.dbg_ts && blockIdx.x == 0 && (j) < 32) p.dbg_ts[8 + 5 * (j) + (k)] = gtime(); } while (0)
#define AVEC_DBG_MODE(bit) ((p.dbg_mode & (bit)) != 0)
#else
#define AVEC_TS(slot) do { } while (0)
#define AVEC_TSK(j, i, base, k) do { } while (0)
#define AVEC_TSJ(j, k) do { } while (0)
#define AVEC_DBG_MODE(bit) false
#endif

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// shared-memory matrix descriptor, SWIZZLE_128B, Blackwell version bits (cute::UMMA::SmemDescriptor layout)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;  // SWIZZLE_128B
    return d;
}
// instruction descriptor for kind::f16: BF16 x BF16 -> F32, M = 128
__device__ __forceinline__ uint32_t make_idesc(int n, int a_mn, int b_mn) {
    uint32_t d = 0;
    d |= 1u << 4;                       // D format F32
    d |= 1u << 7;                       // A format BF16
    d |= 1u << 10;                      // B format BF16
    d |= (uint32_t)(a_mn & 1) << 15;    // A major (0 = K, 1 = MN)
    d |= (uint32_t)(b_mn & 1) << 16;    // B major
    d |= (uint32_t)(n >> 3) << 17;      // N / 8
    d |= (uint32_t)(BM >> 4) << 24;     // M / 16
    return d;
}

// ---------------------------------------------------------------- gather producer
// load 8 consecutive bf16 (one 16-byte smem chunk); nv = number of valid elements (0..8), the rest are zeros
__device__ __forceinline__ uint4 load_chunk(const bf16* p, int nv, int align) {
    uint4 r = make_uint4(0u, 0u, 0u, 0u);
    if (nv <= 0) return r;
    if (nv >= 8) {
        if (align >= 16) return __ldg(reinterpret_cast<const uint4*>(p));
        if (align >= 8) {
            uint2 a = __ldg(reinterpret_cast<const uint2*>(p)), b = __ldg(reinterpret_cast<const uint2*>(p) + 1);
            return make_uint4(a.x, a.y, b.x, b.y);
        }
        if (align >= 4) {
            const uint32_t* q = reinterpret_cast<const uint32_t*>(p);
            return make_uint4(__ldg(q), __ldg(q + 1), __ldg(q + 2), __ldg(q + 3));
        }
    }
    const unsigned short* q = reinterpret_cast<const unsigned short*>(p);
    uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int i = 0; i < 8; ++i)
        if (i < nv) w[i >> 1] |= (uint32_t)__ldg(q + i) << ((i & 1) * 16);
    return make_uint4(w[0], w[1], w[2], w[3]);
}

struct RowInfo { int n, t, h, w; };

// 8 consecutive filter taps of a single-channel input window whose origin is (n, t0, h0, w0); tapofs = kt | kh<<8 | kw<<16
__device__ __forceinline__ uint4 gather_taps8(const bf16* __restrict__ x, const ConvGeom& g, int n, int t0, int h0, int w0, int tap0,
                                              int ntaps, const int* __restrict__ tapofs) {
    const unsigned short* xs = reinterpret_cast<const unsigned short*>(x);
    uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int tap = tap0 + j;
        if (tap < ntaps) {
            const int o = tapofs[tap];
            const int ti = t0 + (o & 255), hi = h0 + ((o >> 8) & 255), wi = w0 + (o >> 16);
            if ((unsigned)ti < (unsigned)g.Ti && (unsigned)hi < (unsigned)g.Hi && (unsigned)wi < (unsigned)g.Wi)
                w[j >> 1] |= (uint32_t)__ldg(xs + (((long long)n * g.Ti + ti) * g.Hi + hi) * g.Wi + wi) << ((j & 1) * 16);
        }
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// ---- asynchronous global -> shared copies (LDGSTS): no register staging, several k-blocks in flight per thread
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// copy one 16-byte chunk (8 bf16, `nv` of them valid, the rest zero-filled) to shared memory
__device__ __forceinline__ void copy_chunk(uint8_t* dst, const bf16* src, int nv, int align, const bf16* safe) {
    const uint32_t d = smem_u32(dst);
    int bytes = nv <= 0 ? 0 : (nv >= 8 ? 16 : nv * 2);
    if (bytes == 0) src = safe;
    if (align >= 16) {
        cp_async16(d, src, bytes);
    } else if (align >= 8) {
        cp_async8(d, src, min(bytes, 8));
        cp_async8(d + 8, bytes > 8 ? src + 4 : safe, max(bytes - 8, 0));
    } else if (align >= 4) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int bq = min(max(bytes - 4 * q, 0), 4);
            cp_async4(d + 4 * q, bq > 0 ? src + 2 * q : safe, bq);
        }
    } else {
        *reinterpret_cast<uint4*>(dst) = load_chunk(src, nv, align);
    }
}

// Fill `rows` smem rows (128 B each, 128B-swizzled, tile base 1024-aligned) of one operand tile for k-block kb.
// `tile0` = first matrix row (K-major kinds) or first MN index (MN-major kinds) of this CTA's tile.
__device__ __forceinline__ void fill_operand(uint8_t* tile, int rows, int kind, const bf16* __restrict__ base, long long ld, int align,
                                             int R, int K, int tile0, int kb, const TcParams& p, const RowInfo* __restrict__ rinfo,
                                             const int* __restrict__ tapofs, int pt) {
    const int chunk = pt & 7;
    const int r_first = pt >> 3;  // 0..15
    // filter tap of this k-block (conv kinds)
    int kt = 0, kh = 0, kw = 0, cb = 0;
    if (kind == OP_CONV_FWD || kind == OP_CONV_DGRAD) {
        int tap = kb / p.cpb; cb = kb % p.cpb;
        kw = tap % p.g.KW; tap /= p.g.KW; kh = tap % p.g.KH; kt = tap / p.g.KH;
    }
    for (int r = r_first; r < rows; r += 16) {
        uint8_t* dst = tile + (size_t)r * 128 + ((chunk ^ (r & 7)) << 4);
        const bf16* src = nullptr;
        int nv = 0;
        if (kind == OP_CONV_TAPS) {
            const RowInfo ri = rinfo[r];
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (ri.n >= 0) v = gather_taps8(base, p.g, ri.n, ri.t, ri.h, ri.w, kb * BKE + chunk * 8, K, tapofs);
            *reinterpret_cast<uint4*>(dst) = v;
            continue;
        }
        if (kind == OP_CONV_TAPS_MN) {
            const long long site = (long long)kb * BKE + (r & 63);
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (site < K) {
                long long s = site;
                const int wo = (int)(s % p.g.Wo); s /= p.g.Wo;
                const int ho = (int)(s % p.g.Ho); s /= p.g.Ho;
                const int to = (int)(s % p.g.To); const int n = (int)(s / p.g.To);
                v = gather_taps8(base, p.g, n, to * p.g.st - p.g.pt, ho * p.g.sh - p.g.ph, wo * p.g.sw - p.g.pw,
                                 tile0 + (r >> 6) * 64 + chunk * 8, R, tapofs);
            }
            *reinterpret_cast<uint4*>(dst) = v;
            continue;
        }
        if (kind == OP_PLAIN_K) {
            const int row = tile0 + r, k = kb * BKE + chunk * 8;
            if (row < R) { src = base + (long long)row * ld + k; nv = K - k; }
        } else if (kind == OP_PLAIN_MN) {
            // smem row = (group g = r / 64, reduction index kk = r % 64); 128 bytes = MN indices tile0 + 64 g + [0, 64)
            const int g = r >> 6, kk = kb * BKE + (r & 63), mn = tile0 + g * 64 + chunk * 8;
            if (kk < K) { src = base + (long long)kk * ld + mn; nv = R - mn; }
        } else if (kind == OP_CONV_FWD) {
            const RowInfo ri = rinfo[r];
            const int ti = ri.t + kt, hi = ri.h + kh, wi = ri.w + kw;
            if (ri.n >= 0 && (unsigned)ti < (unsigned)p.g.Ti && (unsigned)hi < (unsigned)p.g.Hi && (unsigned)wi < (unsigned)p.g.Wi) {
                src = base + ((((long long)ri.n * p.g.Ti + ti) * p.g.Hi + hi) * p.g.Wi + wi) * p.g.C + cb * BKE + chunk * 8;
                nv = 8;
            }
        } else if (kind == OP_CONV_DGRAD) {
            const RowInfo ri = rinfo[r];
            const int a = ri.t - kt, b = ri.h - kh, c = ri.w - kw;
            if (ri.n >= 0 && a >= 0 && b >= 0 && c >= 0 && a % p.g.st == 0 && b % p.g.sh == 0 && c % p.g.sw == 0) {
                const int to = a / p.g.st, ho = b / p.g.sh, wo = c / p.g.sw;
                if (to < p.g.To && ho < p.g.Ho && wo < p.g.Wo) {
                    src = base + ((((long long)ri.n * p.g.To + to) * p.g.Ho + ho) * p.g.Wo + wo) * p.g.Co + cb * BKE + chunk * 8;
                    nv = 8;
                }
            }
        } else {  // OP_CONV_WGRAD_X: group g -> (tap, ci block); smem row kk -> output site
            const int g = r >> 6;
            const int G = tile0 / 64 + g;
            const long long site = (long long)kb * BKE + (r & 63);
            if (site < K && G * 64 < R) {
                int tap = G / p.cpb; const int cbl = G % p.cpb;
                const int fw = tap % p.g.KW; tap /= p.g.KW; const int fh = tap % p.g.KH; const int ft = tap / p.g.KH;
                long long s = site;
                const int wo = (int)(s % p.g.Wo); s /= p.g.Wo;
                const int ho = (int)(s % p.g.Ho); s /= p.g.Ho;
                const int to = (int)(s % p.g.To); const int n = (int)(s / p.g.To);
                const int ti = to * p.g.st + ft - p.g.pt, hi = ho * p.g.sh + fh - p.g.ph, wi = wo * p.g.sw + fw - p.g.pw;
                if ((unsigned)ti < (unsigned)p.g.Ti && (unsigned)hi < (unsigned)p.g.Hi && (unsigned)wi < (unsigned)p.g.Wi) {
                    src = base + ((((long long)n * p.g.Ti + ti) * p.g.Hi + hi) * p.g.Wi + wi) * p.g.C + cbl * BKE + chunk * 8;
                    nv = 8;
                }
            }
        }
        copy_chunk(dst, src, src ? nv : 0, align, base);
    }
}

// ---- vectorised epilogue helpers: 16 consecutive columns of one output row
__device__ __forceinline__ void load16(const void* p, int dtype, size_t idx, float (&v)[16]) {
    if (dtype == AVEC_F32) {
        const float* q = reinterpret_cast<const float*>(p) + idx;
        if ((reinterpret_cast<uintptr_t>(q) & 15) == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { float4 t = __ldg(reinterpret_cast<const float4*>(q) + i); v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w; }
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __ldg(q + i);
        }
    } else {
        const bf16* q = reinterpret_cast<const bf16*>(p) + idx;
        const uintptr_t a = reinterpret_cast<uintptr_t>(q);
        uint32_t w[8];
        if ((a & 15) == 0) {
            uint4 t0 = __ldg(reinterpret_cast<const uint4*>(q)), t1 = __ldg(reinterpret_cast<const uint4*>(q) + 1);
            w[0] = t0.x; w[1] = t0.y; w[2] = t0.z; w[3] = t0.w; w[4] = t1.x; w[5] = t1.y; w[6] = t1.z; w[7] = t1.w;
        } else if ((a & 7) == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { uint2 t = __ldg(reinterpret_cast<const uint2*>(q) + i); w[2 * i] = t.x; w[2 * i + 1] = t.y; }
        } else if ((a & 3) == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) w[i] = __ldg(reinterpret_cast<const uint32_t*>(q) + i);
        } else {
            const unsigned short* h = reinterpret_cast<const unsigned short*>(q);
#pragma unroll
            for (int i = 0; i < 8; ++i) w[i] = (uint32_t)__ldg(h + 2 * i) | ((uint32_t)__ldg(h + 2 * i + 1) << 16);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u); }
    }
}
__device__ __forceinline__ void store16(void* p, int dtype, size_t idx, const float (&v)[16]) {
    if (dtype == AVEC_F32) {
        float* q = reinterpret_cast<float*>(p) + idx;
        if ((reinterpret_cast<uintptr_t>(q) & 15) == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) reinterpret_cast<float4*>(q)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) q[i] = v[i];
        }
    } else {
        bf16* q = reinterpret_cast<bf16*>(p) + idx;
        const uintptr_t a = reinterpret_cast<uintptr_t>(q);
        uint32_t w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            w[i] = *reinterpret_cast<uint32_t*>(&t);
        }
        if ((a & 15) == 0) {
            reinterpret_cast<uint4*>(q)[0] = make_uint4(w[0], w[1], w[2], w[3]);
            reinterpret_cast<uint4*>(q)[1] = make_uint4(w[4], w[5], w[6], w[7]);
        } else if ((a & 7) == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) reinterpret_cast<uint2*>(q)[i] = make_uint2(w[2 * i], w[2 * i + 1]);
        } else if ((a & 3) == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) reinterpret_cast<uint32_t*>(q)[i] = w[i];
        } else {
            unsigned short* h = reinterpret_cast<unsigned short*>(q);
#pragma unroll
            for (int i = 0; i < 8; ++i) { h[2 * i] = (unsigned short)(w[i] & 0xFFFF); h[2 * i + 1] = (unsigned short)(w[i] >> 16); }
        }
    }
}

// m-tile (or, for the TMA wgrad, reduction-tile) index -> first image / first image row of a row-aligned conv tile
__device__ __forceinline__ void conv_tile_origin(const TcParams& p, int t, int& n0, int& h0) {
    if (p.ct_BI == 1) { n0 = t / p.ct_tph; h0 = (t % p.ct_tph) * p.ct_BH; }
    else { n0 = t * p.ct_BI; h0 = 0; }
}

// One thread issues the TMA loads of one operand for k-block kb into `tile`.
__device__ __forceinline__ void tma_fill(const TcParams& p, int kind, const CUtensorMap* map, uint8_t* tile, uint64_t* bar, int rows,
                                         int group_stride, int tile0, int kb, int mtile, bool is_a) {
    const uint32_t dst = smem_u32(tile);
    if (kind == OP_TMA_K) {
        tma_load_2d(dst, map, bar, kb * BKE, tile0 - (is_a ? p.dbg_rowofs : 0));
    } else if (kind == OP_TMA_MN) {
        for (int g = 0; g * 64 < rows; ++g) tma_load_2d(dst + g * group_stride, map, bar, tile0 + g * 64, kb * BKE);
    } else if (kind == OP_TMA_CONV_K) {
        int tap = kb / p.cpb; const int cb = kb % p.cpb;
        const int kw = tap % p.g.KW, kh = tap / p.g.KW;
        int n0, h0;
        conv_tile_origin(p, mtile, n0, h0);
        const int dh = p.ct_dgrad ? p.g.ph - kh : kh - p.g.ph, dw = p.ct_dgrad ? p.g.pw - kw : kw - p.g.pw;
        tma_load_4d(dst, map, bar, cb * BKE, dw, h0 * p.ct_s + dh, n0);
    } else {  // OP_TMA_CONV_MN: the k-block is the conv tile kb
        int n0, h0;
        conv_tile_origin(p, kb, n0, h0);
        const int ngroups = rows / (p.ksteps * 16);
        for (int g = 0; g < ngroups; ++g) {
            if (is_a) {
                tma_load_4d(dst + g * group_stride, map, bar, tile0 + g * 64, 0, h0, n0);
            } else {
                const int G = tile0 / 64 + g;
                int tap = G / p.cpb; const int cb = G % p.cpb;
                const int kw = tap % p.g.KW, kh = tap / p.g.KW;
                tma_load_4d(dst + g * group_stride, map, bar, cb * BKE, kw - p.g.pw, h0 * p.ct_s + kh - p.g.ph, n0);
            }
        }
    }
}

struct TileInfo {
    int mtile, n0, z, kb_begin, nkb, m0, rows_valid;
    long long row_base;
};

__device__ __forceinline__ TileInfo decode_tile(const TcParams& p, int t) {
    TileInfo ti;
    ti.mtile = t % p.grid_m; t /= p.grid_m;
    ti.n0 = (t % p.grid_n) * p.BN;
    ti.z = t / p.grid_n;
    ti.kb_begin = ti.z * p.kb_per_split;
    ti.nkb = min(p.num_kb, ti.kb_begin + p.kb_per_split) - ti.kb_begin;
    ti.m0 = ti.mtile * BM;
    if (p.conv_tiles && p.a_kind == OP_TMA_CONV_HALO) {
        int tn0, th0;
        conv_tile_origin(p, ti.mtile, tn0, th0);
        ti.row_base = ((long long)tn0 * p.ct_H + th0) * p.ct_W;
        ti.rows_valid = min(p.ct_BH, p.ct_H - th0);   // image rows (see tile_row)
    } else if (p.conv_tiles && p.a_kind == OP_TMA_CONV_K) {
        int tn0, th0;
        conv_tile_origin(p, ti.mtile, tn0, th0);
        ti.row_base = p.rm_on ? ((long long)tn0 * p.rm_Hi + p.rm_sy * th0 + p.rm_a) * p.rm_Wi + p.rm_b : ((long long)tn0 * p.ct_H + th0) * p.ct_W;
        ti.rows_valid = p.ct_BI == 1 ? min(p.ct_BH, p.ct_H - th0) * p.ct_W : min(p.ct_BI, p.g.N - tn0) * p.ct_H * p.ct_W;
    } else {
        ti.row_base = ti.m0;
        ti.rows_valid = min(BM, p.M - ti.m0);
    }
    return ti;
}


// accumulator row r (0..127) of a tile -> output row (false: the row holds no output)
__device__ __forceinline__ bool tile_row(const TcParams& p, const TileInfo& ti, int r, long long& row) {
    if (p.halo_bytes) {   // rows live on the (W+2)-wide halo grid
        const int hh = r / p.halo_W2, ww = r - hh * p.halo_W2;
        row = ti.row_base + (long long)hh * p.ct_W + ww;
        return ww < p.ct_W && hh < ti.rows_valid;   // rows_valid = image rows of this tile
    }
    row = ti.row_base + r;
    return r < ti.rows_valid;
}

// ---- 8 consecutive columns of one output row (transposed-domain epilogue: 8 lanes cover one 64-column row segment)
__device__ __forceinline__ void load8(const void* p, int dtype, size_t idx, float (&v)[8]) {
    if (dtype == AVEC_F32) {
        const float* q = reinterpret_cast<const float*>(p) + idx;
        if ((reinterpret_cast<uintptr_t>(q) & 15) == 0) {
            float4 a = __ldg(reinterpret_cast<const float4*>(q)), b = __ldg(reinterpret_cast<const float4*>(q) + 1);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __ldg(q + i);
        }
    } else {
        const bf16* q = reinterpret_cast<const bf16*>(p) + idx;
        const uintptr_t a = reinterpret_cast<uintptr_t>(q);
        uint32_t w[4];
        if ((a & 15) == 0) { uint4 t = __ldg(reinterpret_cast<const uint4*>(q)); w[0] = t.x; w[1] = t.y; w[2] = t.z; w[3] = t.w; }
        else if ((a & 7) == 0) { uint2 t0 = __ldg(reinterpret_cast<const uint2*>(q)), t1 = __ldg(reinterpret_cast<const uint2*>(q) + 1); w[0] = t0.x; w[1] = t0.y; w[2] = t1.x; w[3] = t1.y; }
        else if ((a & 3) == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) w[i] = __ldg(reinterpret_cast<const uint32_t*>(q) + i);
        } else {
            const unsigned short* h = reinterpret_cast<const unsigned short*>(q);
#pragma unroll
            for (int i = 0; i < 4; ++i) w[i] = (uint32_t)__ldg(h + 2 * i) | ((uint32_t)__ldg(h + 2 * i + 1) << 16);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u); }
    }
}
__device__ __forceinline__ void store8(void* p, int dtype, size_t idx, const float (&v)[8]) {
    if (dtype == AVEC_F32) {
        float* q = reinterpret_cast<float*>(p) + idx;
        if ((reinterpret_cast<uintptr_t>(q) & 15) == 0) {
            reinterpret_cast<float4*>(q)[0] = make_float4(v[0], v[1], v[2], v[3]);
            reinterpret_cast<float4*>(q)[1] = make_float4(v[4], v[5], v[6], v[7]);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) q[i] = v[i];
        }
    } else {
        bf16* q = reinterpret_cast<bf16*>(p) + idx;
        const uintptr_t a = reinterpret_cast<uintptr_t>(q);
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]); w[i] = *reinterpret_cast<uint32_t*>(&t); }
        if ((a & 15) == 0) *reinterpret_cast<uint4*>(q) = make_uint4(w[0], w[1], w[2], w[3]);
        else if ((a & 7) == 0) { reinterpret_cast<uint2*>(q)[0] = make_uint2(w[0], w[1]); reinterpret_cast<uint2*>(q)[1] = make_uint2(w[2], w[3]); }
        else if ((a & 3) == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) reinterpret_cast<uint32_t*>(q)[i] = w[i];
        } else {
            unsigned short* h = reinterpret_cast<unsigned short*>(q);
#pragma unroll
            for (int i = 0; i < 4; ++i) { h[2 * i] = (unsigned short)(w[i] & 0xFFFF); h[2 * i + 1] = (unsigned short)(w[i] >> 16); }
        }
    }
}

constexpr int STG_LD = 68;                       // floats per staged row (64 columns + pad: conflict-free 128-bit accesses)
constexpr int STG_WARP = 32 * STG_LD;            // floats per warp
constexpr int STG_BYTES = 4 * STG_WARP * 4;      // 34816 bytes per CTA

// ---- explicit shared-space accesses for the epilogue staging (the staging pointer is derived from a runtime select of two
// bases, so plain C++ accesses compile to generic LD.E / ST.E with their longer latency)
__device__ __forceinline__ void sts128(uint32_t a, float x, float y, float z, float w) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t a) {
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(a) : "memory");
    return r;
}
// 64 accumulator columns of this lane's row with ONE wait (two x32 loads in flight)
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, float (&v)[64]) {
    uint32_t r[64];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%64];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
        "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%65];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]),
          "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]),
          "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]),
          "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
        : "r"(taddr), "r"(taddr + 32u)
        : "memory");
#pragma unroll
    for (int i = 0; i < 64; ++i) v[i] = __uint_as_float(r[i]);
}
// 8 consecutive bf16 (aligned to 16 bytes when AL16, else to 8 bytes) <-> packed words
// `half`: only the first 4 of the 8 columns exist (last column group of a matrix whose width is 4 mod 8, e.g. N = 180)
template <bool AL16>
__device__ __forceinline__ uint4 ldg_bf16x8(const bf16* q, bool half = false) {
    if (half) { const uint2 a = __ldg(reinterpret_cast<const uint2*>(q)); return make_uint4(a.x, a.y, 0u, 0u); }
    if (AL16) return __ldg(reinterpret_cast<const uint4*>(q));
    const uint2 a = __ldg(reinterpret_cast<const uint2*>(q)), b = __ldg(reinterpret_cast<const uint2*>(q) + 1);
    return make_uint4(a.x, a.y, b.x, b.y);
}
template <bool AL16>
__device__ __forceinline__ void stg_bf16x8(bf16* q, const float (&v)[8], bool half = false) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]); w[i] = *reinterpret_cast<uint32_t*>(&t); }
    if (half) reinterpret_cast<uint2*>(q)[0] = make_uint2(w[0], w[1]);
    else if (AL16) *reinterpret_cast<uint4*>(q) = make_uint4(w[0], w[1], w[2], w[3]);
    else { reinterpret_cast<uint2*>(q)[0] = make_uint2(w[0], w[1]); reinterpret_cast<uint2*>(q)[1] = make_uint2(w[2], w[3]); }
}
__device__ __forceinline__ void unpack_bf16x8(const uint4& t, float (&v)[8]) {
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u); }
}

// Fast epilogue of one 128 x BN tile for one warp: the common launch-constant cases (bf16 output / auxiliary operands with
// rows aligned to 8 or 16 bytes, N % 8 == 0, fp32 reductions for ACCUM) compiled per epilogue kind, so the transposed-domain
// loop is branch-free: four row passes at a time with all their shared-memory and auxiliary loads issued before the first
// use, bias kept in registers (a lane owns the same 8 columns in every pass).
template <int KIND, bool STATS, bool AL16, bool DROP = false>
__device__ __forceinline__ void epilogue_fast(const TcParams& p, const EpiParams& ep, const TileInfo& ti, uint32_t lane_addr, uint32_t stg_s,
                                              uint32_t bias_sa, float* stats_dst, int warp, int lane, bool bias_on, uint32_t rowrel_sa = 0u,
                                              int slab0 = 0, int slab_step = 64, const uint32_t* kbits = nullptr) {
    const int BN = p.BN;
    const int q = lane & 7, rs = lane >> 3;
    const int lc = q * 8;
    const float alpha = ep.alpha;
    const bool scale_on = alpha != 1.0f;   // kernel-uniform: the common alpha == 1 / no-bias launches skip 16 FP32 ops per row segment
    constexpr bool drop_on = DROP;   // nn.Dropout fused behind this GEMM: its own instantiation, so the plain epilogues keep their register budget
    const int rows_left = ti.rows_valid - warp * 32;   // valid rows of this warp's quarter (may be <= 0 or >= 32)
    // output row of tile row r: contiguous, or scattered through the row table (parity classes of strided dgrads)
    auto out_row = [&](int rr) -> long long {   // < 0: the tile row holds no output (halo-grid columns past the image)
        const int r = warp * 32 + rr;
        if (p.rm_on) {
            int rel;
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(rel) : "r"(rowrel_sa + (uint32_t)r * 4u) : "memory");
            return rel == (int)0x80000000 ? -1LL : ti.row_base + rel;
        }
        return ti.row_base + r;
    };
    int si = 0;
    for (int c0 = slab0; c0 < BN; c0 += slab_step, ++si) {   // with two epilogue warpgroups each takes every other 64-column slab
        const int ncol = min(64, BN - c0);   // multiple of 16
        // keep bits of this slab (drawn by the caller while the MMAs were still running): bit (u * 8 + j) of word g
        uint32_t kw0 = 0u, kw1 = 0u;
        if (drop_on) {
            kw0 = si == 0 ? kbits[0] : si == 1 ? kbits[2] : si == 2 ? kbits[4] : kbits[6];
            kw1 = si == 0 ? kbits[1] : si == 1 ? kbits[3] : si == 2 ? kbits[5] : kbits[7];
        }
        const int gc = ti.n0 + c0 + lc;
        const bool col_ok = lc < ncol && gc < p.N;
        const bool half = gc + 8 > p.N;     // N % 8 == 4: the last column group holds 4 columns (8-byte accesses only)
        // ---- auxiliary operand of the first four passes: does not depend on the accumulator, so it is requested first
        uint4 xa[4];
        constexpr bool HAS_AUX = KIND == AVEC_EPI_RESIDUAL || KIND == AVEC_EPI_DSWISH || KIND == AVEC_EPI_RELU;
        const bool aux_on = HAS_AUX && ep.aux != nullptr;
        const bf16* auxp = reinterpret_cast<const bf16*>(ep.aux);
        if (HAS_AUX) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int rr = u * 4 + rs;
                xa[u] = make_uint4(0u, 0u, 0u, 0u);
                if (aux_on && col_ok && rr < rows_left) { const long long ro = out_row(rr); if (ro >= 0) xa[u] = ldg_bf16x8<AL16>(auxp + (size_t)ro * ep.ldaux + gc, half); }
            }
        }
        // ---- row domain: TMEM -> registers -> staging
        if (ncol == 64) {
            float v[64];
            tmem_ld64(lane_addr + (uint32_t)c0, v);
#pragma unroll
            for (int j = 0; j < 64; j += 4) sts128(stg_s + (uint32_t)(lane * STG_LD + j) * 4u, v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
            for (int cc = 0; cc < ncol; cc += 16) {
                float v[16];
                tmem_ld16(lane_addr + (uint32_t)(c0 + cc), v);
#pragma unroll
                for (int j = 0; j < 16; j += 4) sts128(stg_s + (uint32_t)(lane * STG_LD + cc + j) * 4u, v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
        }
        uint4 xb[4];
        if (HAS_AUX) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int rr = (4 + u) * 4 + rs;
                xb[u] = make_uint4(0u, 0u, 0u, 0u);
                if (aux_on && col_ok && rr < rows_left) { const long long ro = out_row(rr); if (ro >= 0) xb[u] = ldg_bf16x8<AL16>(auxp + (size_t)ro * ep.ldaux + gc, half); }
            }
        }
        __syncwarp();
        // ---- transposed domain
        float b[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (bias_on) {
            const float4 b0 = lds128(bias_sa + (uint32_t)(c0 + (col_ok ? lc : 0)) * 4u), b1 = lds128(bias_sa + (uint32_t)(c0 + (col_ok ? lc : 0) + 4) * 4u);
            b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
        }
        float s1[8], s2[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { s1[j] = 0.0f; s2[j] = 0.0f; }
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            float4 t[8];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int rr = (g * 4 + u) * 4 + rs;
                const uint32_t sa = stg_s + (uint32_t)(rr * STG_LD + (col_ok ? lc : 0)) * 4u;
                t[2 * u] = lds128(sa); t[2 * u + 1] = lds128(sa + 16u);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int rr = (g * 4 + u) * 4 + rs;
                if (!(col_ok && rr < rows_left)) continue;
                const long long row_ = out_row(rr);
                if (row_ < 0) continue;
                const size_t row = (size_t)row_;
                float a[8] = {t[2 * u].x, t[2 * u].y, t[2 * u].z, t[2 * u].w, t[2 * u + 1].x, t[2 * u + 1].y, t[2 * u + 1].z, t[2 * u + 1].w};
                if (bias_on) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) a[j] += b[j];
                }
                const size_t oi = row * ep.ldo + gc;
                float sa_[8];   // alpha * a (a itself feeds the BatchNorm statistics and the Swish pre-activation copy)
#pragma unroll
                for (int j = 0; j < 8; ++j) sa_[j] = a[j];
                if (scale_on) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) sa_[j] *= alpha;
                }
                if (KIND == AVEC_EPI_ACCUM) {
                    float* dst = reinterpret_cast<float*>(ep.out) + oi;
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(sa_[0]), "f"(sa_[1]), "f"(sa_[2]), "f"(sa_[3]) : "memory");
                    if (!half) asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4), "f"(sa_[4]), "f"(sa_[5]), "f"(sa_[6]), "f"(sa_[7]) : "memory");
                } else {
                    float o[8], x[8], kp[8];
                    if (HAS_AUX) unpack_bf16x8(g == 0 ? xa[u] : xb[u], x);
                    if (KIND == AVEC_EPI_SWISH && ep.out2) stg_bf16x8<AL16>(reinterpret_cast<bf16*>(ep.out2) + row * ep.ldo2 + gc, a, half);
#pragma unroll
                    for (int j = 0; j < 8; ++j) kp[j] = 1.0f;
                    if (drop_on) {   // the mask avec_dropout draws for this (row, column group)
                        const uint32_t kb = (g == 0 ? kw0 : kw1) >> (u * 8);
#pragma unroll
                        for (int j = 0; j < 8; ++j) kp[j] = ((kb >> j) & 1u) ? ep.drop_scale : 0.0f;
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (KIND == AVEC_EPI_LINEAR) o[j] = sa_[j] * kp[j];
                        else if (KIND == AVEC_EPI_SWISH) o[j] = swishf_(a[j]) * kp[j];
                        else if (KIND == AVEC_EPI_RESIDUAL) o[j] = x[j] + sa_[j] * kp[j];
                        else if (KIND == AVEC_EPI_DSWISH) o[j] = sa_[j] * dswishf_(x[j]) * kp[j];
                        else o[j] = fmaxf(sa_[j] + x[j], 0.0f);
                    }
                    stg_bf16x8<AL16>(reinterpret_cast<bf16*>(ep.out) + oi, o, half);
                }
                if (STATS) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) { s1[j] += a[j]; s2[j] += a[j] * a[j]; }
                }
            }
        }
        __syncwarp();
        if (STATS) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], 8);  s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], 8);
                s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], 16); s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], 16);
            }
            if (col_ok) {
                const bool hi = (rs & 1) != 0, sq = rs >= 2;
                const float r0 = sq ? (hi ? s2[4] : s2[0]) : (hi ? s1[4] : s1[0]);
                const float r1 = sq ? (hi ? s2[5] : s2[1]) : (hi ? s1[5] : s1[1]);
                const float r2 = sq ? (hi ? s2[6] : s2[2]) : (hi ? s1[6] : s1[2]);
                const float r3 = sq ? (hi ? s2[7] : s2[3]) : (hi ? s1[7] : s1[3]);
                float* dst = stats_dst + (sq ? p.N : 0) + gc + (hi ? 4 : 0);   // 16-byte aligned: N % 8 == 0, gc % 8 == 0
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(r0), "f"(r1), "f"(r2), "f"(r3) : "memory");
            }
        }
    }
}

// Epilogue of one 128 x BN tile for one warp (32 accumulator rows).  The accumulator comes out of TMEM one ROW per lane
// (tcgen05.ld 32x32b); writing global memory in that shape would touch 32 different cache lines per instruction, so the
// 64-column slab is staged in shared memory and re-read TRANSPOSED: 8 lanes cover one row's 64 columns (one full 128-byte
// line of bf16), 4 rows per instruction.  Bias, activation, residual / Swish' operands, fp32 split-K reductions and the
// BatchNorm column statistics are all applied in that coalesced domain.
__device__ __forceinline__ void epilogue_warp(const TcParams& p, const EpiParams& ep, const TileInfo& ti, uint32_t lane_addr, float* stg,
                                              const float* bias_s, float* stats_dst, int warp, int lane, int slab0 = 0, int slab_step = 64) {
    const int BN = p.BN;
    const int q = lane & 7, rs = lane >> 3;
    for (int c0 = slab0; c0 < BN; c0 += slab_step) {
        const int ncol = min(64, BN - c0);   // multiple of 16
        // ---- row domain: TMEM -> registers (+ bias) -> staging
        for (int cc = 0; cc < ncol; cc += 32) {
            float v[32];
            const bool two = ncol - cc >= 32;
            if (two) tmem_ld32(lane_addr + (uint32_t)(c0 + cc), v);
            else tmem_ld16(lane_addr + (uint32_t)(c0 + cc), *reinterpret_cast<float(*)[16]>(&v[0]));
            const int nv = two ? 32 : 16;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                if (j < nv) {
                    const float4 b = *reinterpret_cast<const float4*>(bias_s + c0 + cc + j);
                    *reinterpret_cast<float4*>(stg + lane * STG_LD + cc + j) = make_float4(v[j] + b.x, v[j + 1] + b.y, v[j + 2] + b.z, v[j + 3] + b.w);
                }
            }
        }
        __syncwarp();
        // ---- transposed domain
        float s1[8], s2[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { s1[j] = 0.0f; s2[j] = 0.0f; }
        const int lc = q * 8;                 // column inside the slab
        const int gc = ti.n0 + c0 + lc;       // global column
        if (lc < ncol) {
#pragma unroll 2
            for (int pass = 0; pass < 8; ++pass) {
                const int rr = pass * 4 + rs;
                long long row;
                if (!tile_row(p, ti, warp * 32 + rr, row)) continue;
                float a[8];
                {
                    const float4 t0 = *reinterpret_cast<const float4*>(stg + rr * STG_LD + lc);
                    const float4 t1 = *reinterpret_cast<const float4*>(stg + rr * STG_LD + lc + 4);
                    a[0] = t0.x; a[1] = t0.y; a[2] = t0.z; a[3] = t0.w; a[4] = t1.x; a[5] = t1.y; a[6] = t1.z; a[7] = t1.w;
                }
                if (gc + 8 <= p.N) {
                    const size_t oi = (size_t)row * ep.ldo + gc;
                    float o[8];
                    if (ep.kind == AVEC_EPI_ACCUM) {
                        float* dst = reinterpret_cast<float*>(ep.out) + oi;
                        if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(ep.alpha * a[0]), "f"(ep.alpha * a[1]),
                                         "f"(ep.alpha * a[2]), "f"(ep.alpha * a[3]) : "memory");
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4), "f"(ep.alpha * a[4]), "f"(ep.alpha * a[5]),
                                         "f"(ep.alpha * a[6]), "f"(ep.alpha * a[7]) : "memory");
                        } else {
#pragma unroll
                            for (int j = 0; j < 8; ++j) atomicAdd(dst + j, ep.alpha * a[j]);
                        }
                    } else {
                        if (ep.kind == AVEC_EPI_LINEAR) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) o[j] = ep.alpha * a[j];
                        } else if (ep.kind == AVEC_EPI_SWISH) {
                            if (ep.out2) store8(ep.out2, ep.out2_dtype, (size_t)row * ep.ldo2 + gc, a);
#pragma unroll
                            for (int j = 0; j < 8; ++j) o[j] = swishf_(a[j]);
                        } else {
                            float x[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) x[j] = 0.0f;
                            if (ep.aux) load8(ep.aux, ep.aux_dtype, (size_t)row * ep.ldaux + gc, x);
                            if (ep.kind == AVEC_EPI_RESIDUAL) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) o[j] = x[j] + ep.alpha * a[j];
                            } else if (ep.kind == AVEC_EPI_DSWISH) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) o[j] = ep.alpha * a[j] * dswishf_(x[j]);
                            } else {  // RELU
#pragma unroll
                                for (int j = 0; j < 8; ++j) o[j] = fmaxf(ep.alpha * a[j] + x[j], 0.0f);
                            }
                        }
                        store8(ep.out, ep.out_dtype, oi, o);
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) { s1[j] += a[j]; s2[j] += a[j] * a[j]; }
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (gc + j < p.N) {
                            epilogue_elem(ep, (int)row, gc + j, a[j]);   // bias already added (ep.bias == nullptr)
                            s1[j] += a[j]; s2[j] += a[j] * a[j];
                        }
                    }
                }
            }
        }
        __syncwarp();
        if (ep.colstats) {
            // lanes with equal q hold partial sums of the same 8 columns: fold the 4 row groups (every lane ends up with the
            // totals), then each row group flushes one quarter with a single 128-bit reduction: {sum 0-3, sum 4-7, sq 0-3, sq 4-7}
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], 8);  s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], 8);
                s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], 16); s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], 16);
            }
            if (lc < ncol) {
                const bool hi = (rs & 1) != 0, sq = rs >= 2;
                const float r0 = sq ? (hi ? s2[4] : s2[0]) : (hi ? s1[4] : s1[0]);
                const float r1 = sq ? (hi ? s2[5] : s2[1]) : (hi ? s1[5] : s1[1]);
                const float r2 = sq ? (hi ? s2[6] : s2[2]) : (hi ? s1[6] : s1[2]);
                const float r3 = sq ? (hi ? s2[7] : s2[3]) : (hi ? s1[7] : s1[3]);
                const int col = gc + (hi ? 4 : 0);
                float* dst = stats_dst + (sq ? p.N : 0) + col;
                if (col + 4 <= p.N && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(r0), "f"(r1), "f"(r2), "f"(r3) : "memory");
                } else {
                    if (col < p.N) atomicAdd(dst, r0);
                    if (col + 1 < p.N) atomicAdd(dst + 1, r1);
                    if (col + 2 < p.N) atomicAdd(dst + 2, r2);
                    if (col + 3 < p.N) atomicAdd(dst + 3, r3);
                }
            }
        }
    }
}

template <bool AL16>
__device__ __forceinline__ void epilogue_fast_dispatch(const TcParams& p, const EpiParams& ep, const TileInfo& ti, uint32_t lane_addr, uint32_t stg_s,
                                                       uint32_t bias_sa, float* stats_dst, int warp, int lane, bool bias_on, uint32_t rowrel_sa, int slab0, int slab_step,
                                                       const uint32_t* kbits) {
    if (ep.drop_rng) {
        switch (ep.kind) {
        case AVEC_EPI_LINEAR: epilogue_fast<AVEC_EPI_LINEAR, false, AL16, true>(p, ep, ti, lane_addr, stg_s, bias_sa, stats_dst, warp, lane, bias_on, rowrel_sa, slab0, slab_step, kbits); break;
        case AVEC_EPI_SWISH: epilogue_fast<AVEC_EPI_SWISH, false, AL16, true>(p, ep, ti, lane_addr, stg_s, bias_sa, stats_dst, warp, lane, bias_on, rowrel_sa, slab0, slab_step, kbits); break;
        case AVEC_EPI_RESIDUAL: epilogue_fast<AVEC_EPI_RESIDUAL, false, AL16, true>(p, ep, ti, lane_addr, stg_s, bias_sa, stats_dst, warp, lane, bias_on, rowrel_sa, slab0, slab_step, kbits); break;
        default: epilogue_fast<AVEC_EPI_DSWISH, false, AL16, true>(p, ep, ti, lane_addr, stg_s, bias_sa, stats_dst, warp, lane, bias_on, rowrel_sa, slab0, slab_step, kbits); break;
        }
        return;
    }
    switch (ep.kind) {
    case AVEC_EPI_LINEAR:
        if (ep.colstats) epilogue_fast<AVEC_EPI_LINEAR, true, AL16>(p, ep, ti, lane_addr, stg_s, bias_sa, stats_dst, warp, lane, bias_on, rowrel_sa, slab0, slab_step);
        else epilogue_fast<AVEC_EPI_LINEAR, false, AL16>(p, ep, ti, lane_addr, stg_s, bias_sa, stats_dst, warp, lane, bias_on, rowrel_sa, slab0, slab_step);
        break;
    case AVEC_EPI_SWISH: epilogue_fast<AVEC_EPI_SWISH, false, AL16>(p, ep, ti, lane_addr, stg_s, bias_sa, stats_dst, warp, lane, bias_on, rowrel_sa, slab0, slab_step); break;
    case AVEC_EPI_RESIDUAL: epilogue_fast<AVEC_EPI_RESIDUAL, false, AL16>(p, ep, ti, lane_addr, stg_s, bias_sa, stats_dst, warp, lane, bias_on, rowrel_sa, slab0, slab_step); break;
    case AVEC_EPI_DSWISH: epilogue_fast<AVEC_EPI_DSWISH, false, AL16>(p, ep, ti, lane_addr, stg_s, bias_sa, stats_dst, warp, lane, bias_on, rowrel_sa, slab0, slab_step); break;
    case AVEC_EPI_ACCUM: epilogue_fast<AVEC_EPI_ACCUM, false, AL16>(p, ep, ti, lane_addr, stg_s, bias_sa, stats_dst, warp, lane, bias_on, rowrel_sa, slab0, slab_step); break;
    default: epilogue_fast<AVEC_EPI_RELU, false, AL16>(p, ep, ti, lane_addr, stg_s, bias_sa, stats_dst, warp, lane, bias_on, rowrel_sa, slab0, slab_step); break;
    }
}

// Persistent, warp-specialised: every CTA walks the tile list t = blockIdx.x, blockIdx.x + gridDim.x, ... ; the smem ring
// and its phases run on across tiles, and the accumulator is double-buffered in TMEM (2 x BN columns) so that the epilogue
// of tile j overlaps the TMA + MMA main loop of tile j + 1.
__global__ void __launch_bounds__(TC_THREADS_WIDE, 1) gemm_tc_kernel(const __grid_constant__ TcParams p, const __grid_constant__ CUtensorMap mapA,
                                                               const __grid_constant__ CUtensorMap mapB) {
    pdl_trigger();   // programmatic dependent launch: a tcgen05 GEMM behind this kernel may start its prologue now
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem_al = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* halo = smem_al;                                   // [2 slots][cpb][halo_bytes] (halo mode only)
    uint8_t* smem = smem_al + 2 * p.cpb * p.halo_bytes;        // ring
    const int BN = p.BN;
    const int a_bytes = p.a_rows * 128;
    const int b_bytes = ((p.b_rows * 128 + 1023) / 1024) * 1024;
    const int stage_bytes = a_bytes + b_bytes;
    const int epi_groups = blockDim.x > TC_THREADS ? 2 : 1;
    uint8_t* ctrl = smem + (size_t)p.stages * stage_bytes + (p.stg_dedicated ? STG_BYTES * epi_groups : 0);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(ctrl);
    uint64_t* empty_bar = full_bar + 8;
    uint64_t* accum_full = empty_bar + 8;    // [2]
    uint64_t* accum_empty = accum_full + 2;  // [2]
    uint64_t* halo_full = accum_empty + 2;   // [2]
    uint64_t* halo_empty = halo_full + 2;    // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(halo_empty + 2);
    RowInfo* rinfo = reinterpret_cast<RowInfo*>(ctrl + 256);
    int* tapofs = reinterpret_cast<int*>(ctrl + 256 + BM * sizeof(RowInfo));
    float* bias_s = reinterpret_cast<float*>(ctrl + 256 + BM * sizeof(RowInfo) + 256 * sizeof(int));   // [256]

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform by construction (and known to be so by ptxas)
    if (tid == 0) AVEC_TS(0);   // kernel start
    const bool a_tma = is_tma(p.a_kind), b_tma = is_tma(p.b_kind);
    const bool any_gather = !a_tma || !b_tma, any_tma = a_tma || b_tma;
    const int tiles_total = p.grid_m * p.grid_n * p.splits;

    // TMEM columns: two accumulator buffers of BN columns, power of two >= 32
    uint32_t ncols = 32;
    while ((int)ncols < 2 * BN) ncols <<= 1;

    if (tid == 0) {
        const uint32_t full_count = (any_gather ? PRODUCER_THREADS : 0) + (any_tma ? 1 : 0);
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], full_count); mbar_init(&empty_bar[s], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&accum_full[i], 1); mbar_init(&accum_empty[i], PRODUCER_THREADS * epi_groups);
            mbar_init(&halo_full[i], 1); mbar_init(&halo_empty[i], 1);
        }
        fence_barrier_init();
        if (a_tma) tma_prefetch_desc(&mapA);
        if (b_tma) tma_prefetch_desc(&mapB);
    }
    if (warp == 4) tmem_alloc(tmem_slot, ncols);
    if (p.a_kind == OP_CONV_TAPS || p.b_kind == OP_CONV_TAPS_MN) {
        for (int tap = tid; tap < 256; tap += blockDim.x) {
            int t = tap;
            const int kw = t % p.g.KW; t /= p.g.KW; const int kh = t % p.g.KH; const int kt = t / p.g.KH;
            tapofs[tap] = kt | (kh << 8) | (kw << 16);
        }
    }
    // MN-major TMA tiles whose k extent (conv tile rows) is not a multiple of 16: the tail rows are never written by TMA
    // and must read as zeros -> clear all stages once (generic proxy), then hand the buffers to the async proxy
    if (p.a_kind == OP_TMA_CONV_MN || p.b_kind == OP_TMA_CONV_MN) {
        uint4* z = reinterpret_cast<uint4*>(smem);
        const int n16 = p.stages * stage_bytes / 16;
        for (int i = tid; i < n16; i += blockDim.x) z[i] = make_uint4(0u, 0u, 0u, 0u);
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // everything above touched only this CTA's shared memory / TMEM and the kernel parameters: with a programmatic dependent
    // launch it overlapped the tail of the preceding kernel.  From here on global memory is read and written.
    pdl_wait();
    if (tid == 0) AVEC_TS(1);   // setup done (barriers, TMEM alloc)

    if (warp < 4 || warp >= 6) {
        // ===================== gather producers (when an operand is not TMA-fed), then epilogue =====================
        // warps 0-3: producers + epilogue warpgroup 0; warps 6-9 (wide launches): epilogue warpgroup 1 (odd 64-column slabs)
        const int grp = warp >= 6 ? 1 : 0, qw = warp & 3;   // qw = TMEM lane quarter this warp may access
        const int nepi = PRODUCER_THREADS * epi_groups;
        if (p.rm_on && grp == 0) {
            // tile row r = ((image il) * ct_H' + row i) * ct_W + column j  ->  relative output row on the strided grid
            const int per_img = p.ct_BI == 1 ? 0x7fffffff : p.ct_H * p.ct_W;
            const int il = tid / per_img, rem = tid - il * per_img;
            const int i = rem / p.ct_W, j = rem - i * p.ct_W;
            reinterpret_cast<int*>(rinfo)[tid] = (il * p.rm_Hi + p.rm_sy * i) * p.rm_Wi + p.rm_sx * j;
            asm volatile("bar.sync 3, 128;" ::: "memory");   // (warpgroup 1 sees the table after the first refresh barrier)
        }
        int it = 0;   // k-blocks pushed through the ring so far
        int j = 0;    // local tile counter
        int cur_n0 = -1;
        bool cur_bias_on = false;
        for (int t = blockIdx.x; t < tiles_total; t += gridDim.x, ++j) {
            const TileInfo ti = decode_tile(p, t);
            const int n0 = ti.n0, nkb = ti.nkb;
            // per-tile shared state (bias slice, conv row decode) is rebuilt only when it changes: persistent CTAs walk the
            // m tiles of one column block back to back, so most tiles skip both barriers
            const bool refresh = j == 0 || n0 != cur_n0 || (ti.z == 0) != cur_bias_on;
            cur_n0 = n0; cur_bias_on = ti.z == 0;
            if (refresh) {
            asm volatile("bar.sync 1, %0;" ::"r"(nepi) : "memory");   // previous tile's epilogues have released bias_s / rinfo
            if (grp == 0) {
            for (int c = tid; c < 256; c += PRODUCER_THREADS)
                bias_s[c] = (p.ep.bias && ti.z == 0 && c < BN && n0 + c < p.N) ? p.ep.bias[n0 + c] : 0.0f;
            }
            if (grp == 0 && (p.a_kind == OP_CONV_FWD || p.a_kind == OP_CONV_DGRAD || p.a_kind == OP_CONV_TAPS)) {
                // per-row site decode for the conv gathers (A operand rows are fixed for the whole tile)
                RowInfo ri; ri.n = -1; ri.t = ri.h = ri.w = 0;
                long long m = (long long)ti.m0 + tid;
                if (m < p.M) {
                    if (p.a_kind != OP_CONV_DGRAD) {
                        int wo = (int)(m % p.g.Wo); m /= p.g.Wo; int ho = (int)(m % p.g.Ho); m /= p.g.Ho; int to = (int)(m % p.g.To);
                        ri.n = (int)(m / p.g.To);
                        ri.t = to * p.g.st - p.g.pt; ri.h = ho * p.g.sh - p.g.ph; ri.w = wo * p.g.sw - p.g.pw;
                    } else {
                        int wi = (int)(m % p.g.Wi); m /= p.g.Wi; int hi = (int)(m % p.g.Hi); m /= p.g.Hi; int ti_ = (int)(m % p.g.Ti);
                        ri.n = (int)(m / p.g.Ti);
                        ri.t = ti_ + p.g.pt; ri.h = hi + p.g.ph; ri.w = wi + p.g.pw;
                    }
                }
                rinfo[tid] = ri;
            }
            asm volatile("bar.sync 1, %0;" ::"r"(nepi) : "memory");
            }
            if (any_gather && grp == 0) {
                // cp.async groups: k-block i is published (proxy fence + mbarrier arrive) LAG iterations after it was issued,
                // so each thread keeps up to LAG+1 k-blocks of loads in flight (stages >= LAG + 1)
                constexpr int LAG = 2;
                for (int i = 0; i < nkb + LAG; ++i) {
                    if (i < nkb) {
                        const int g = it + i;
                        const int s = g % p.stages;
                        const uint32_t ph = (uint32_t)((g / p.stages) & 1);   // (one tile per CTA in gather mode: cheap enough)
                        mbar_wait(&empty_bar[s], ph ^ 1u);
                        uint8_t* a_tile = smem + (size_t)s * stage_bytes;
                        uint8_t* b_tile = a_tile + a_bytes;
                        const int kb = ti.kb_begin + i;
                        if (!a_tma) fill_operand(a_tile, p.a_rows, p.a_kind, p.A, p.a_ld, p.a_align, p.M, p.K, ti.m0, kb, p, rinfo, tapofs, tid);
                        if (!b_tma) fill_operand(b_tile, p.b_rows, p.b_kind, p.B, p.b_ld, p.b_align, p.N, p.K, n0, kb, p, rinfo, tapofs, tid);
                    }
                    cp_async_commit();
                    if (i >= LAG) {
                        cp_async_wait<LAG>();
                        fence_proxy_async();
                        mbar_arrive(&full_bar[(it + i - LAG) % p.stages]);
                    }
                }
            }
            it += nkb;
            // ---------------- epilogue of tile j (accumulator buffer j & 1) ----------------
            const int buf = j & 1;
            const int r = qw * 32 + lane;
            const bool rv = r < ti.rows_valid;
            const long long row = ti.row_base + r;
            // pull this thread's residual / auxiliary row segment towards L1 while the MMAs are still running
            if (rv && p.ep.aux && !p.out_transposed) {
                const int esz = p.ep.aux_dtype == AVEC_F32 ? 4 : 2;
                const char* a0 = reinterpret_cast<const char*>(p.ep.aux) + ((size_t)row * p.ep.ldaux + n0) * esz;
                const int nbytes = min(BN, p.N - n0) * esz;
                // (L2, not L1: with ~220 KB of the SM's array carved out as shared memory there is next to no L1 to prefetch into)
                for (int o = 0; o < nbytes; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(a0 + o));
            }
            const int slab0 = grp * 64, slab_step = 64 * epi_groups;
            // fused nn.Dropout: draw this thread's keep bits (one Philox call per row x 8 columns, <= 4 slabs x 8 rows) now, while
            // the tile's MMAs are still in flight - the epilogue proper then only shifts bits.  Word [2 * slab + g], bit u * 8 + j.
            uint32_t kbits[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
            if (p.ep.drop_rng) {
                const unsigned long long seed = p.ep.drop_rng[0], step = p.ep.drop_rng[1];
                const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
#pragma unroll
                for (int si = 0; si < 4; ++si) {
                    const int c0 = slab0 + si * slab_step;
                    if (c0 < BN) {
                        const uint32_t cg = (uint32_t)((n0 + c0 + (lane & 7) * 8) >> 3);
#pragma unroll
                        for (int gu = 0; gu < 8; ++gu) {
                            const uint32_t drow = (uint32_t)(ti.row_base + qw * 32 + gu * 4 + (lane >> 3));
                            const uint4 rb = philox4x32_10(make_uint4(drow, cg, p.ep.drop_site, (uint32_t)step), key);
                            uint32_t m = 0u;
#pragma unroll
                            for (int jj = 0; jj < 8; ++jj) m |= (bits16(rb, jj) >= p.ep.drop_thresh ? 1u : 0u) << jj;
                            kbits[si * 2 + (gu >> 2)] |= m << ((gu & 3) * 8);
                        }
                    }
                }
            }
            if (tid == 0) AVEC_TSJ(j, 0);
            if (AVEC_DBG_MODE(4)) mbar_wait_sleep(&accum_full[buf], (uint32_t)((j >> 1) & 1));
            else mbar_wait(&accum_full[buf], (uint32_t)((j >> 1) & 1));
            tc_fence_after();
            if (tid == 0 && j == 0) AVEC_TS(4);   // accumulator complete, epilogue starts
            if (tid == 0) AVEC_TSJ(j, 1);
            const uint32_t lane_addr = tmem_base + ((uint32_t)(qw * 32) << 16) + (uint32_t)(buf * BN);
            EpiParams ep = p.ep;
            ep.bias = nullptr;   // the bias slice lives in shared memory (bias_s)
            float* stg = (p.stg_dedicated ? reinterpret_cast<float*>(smem + (size_t)p.stages * stage_bytes) : reinterpret_cast<float*>(smem)) + grp * (STG_BYTES / 4) + qw * STG_WARP;
            // BatchNorm statistics go to one of AVEC_STATS_REPLICAS copies of the accumulator (by tile index): 32x fewer
            // same-address L2 reductions
            float* stats_dst = ep.colstats ? ep.colstats + (size_t)(ti.mtile % AVEC_STATS_REPLICAS) * 2 * p.N : nullptr;
            const bool bias_on = p.ep.bias != nullptr && ti.z == 0;
            if (p.epi_fast == 2) epilogue_fast_dispatch<true>(p, ep, ti, lane_addr, smem_u32(stg), smem_u32(bias_s), stats_dst, qw, lane, bias_on, smem_u32(rinfo), slab0, slab_step, kbits);
            else if (p.epi_fast == 1) epilogue_fast_dispatch<false>(p, ep, ti, lane_addr, smem_u32(stg), smem_u32(bias_s), stats_dst, qw, lane, bias_on, smem_u32(rinfo), slab0, slab_step, kbits);
            else epilogue_warp(p, ep, ti, lane_addr, stg, bias_s, stats_dst, qw, lane, slab0, slab_step);
            // all TMEM reads of this buffer are complete: hand it back to the MMA warp
            tc_fence_before();
            mbar_arrive(&accum_empty[buf]);
            if (tid == 0 && j == 0) AVEC_TS(5);   // epilogue done
            if (tid == 0) AVEC_TSJ(j, 2);
        }
        tc_fence_before();
    } else if (warp == 5) {
        // ===================== TMA producer (whole warp walks the ring, one elected lane issues) =====================
        if (any_tma) {
            // One thread feeds the whole ring: everything per k-block is kept to a handful of scalar instructions (running
            // stage / phase / tap counters, no integer divisions) - with 64-column tiles a k-block is only 128 MMA cycles.
            int j = 0, st = 0;
            uint32_t ph = 0;
            const bool halo_mode = p.a_kind == OP_TMA_CONV_HALO;
            const uint32_t tx = (uint32_t)(((a_tma && !halo_mode) ? p.a_tx : 0) + (b_tma ? p.b_tx : 0));
            const uint32_t ring0 = smem_u32(smem);
            for (int t = blockIdx.x; t < tiles_total; t += gridDim.x, ++j) {
                const TileInfo ti = decode_tile(p, t);
                int tn0 = 0, th0 = 0;
                if (p.conv_tiles && p.a_kind != OP_TMA_CONV_MN) conv_tile_origin(p, ti.mtile, tn0, th0);
                if (halo_mode) {
                    // one halo tile per 64-channel block for this output tile (double-buffered across tiles)
                    const int slot = j & 1;
                    mbar_wait(&halo_empty[slot], (uint32_t)(((j >> 1) & 1) ^ 1));
                    if (elect_one()) {
                        mbar_expect_tx(&halo_full[slot], (uint32_t)(p.cpb * p.a_tx));
                        for (int cb = 0; cb < p.cpb; ++cb)
                            tma_load_4d(smem_u32(halo + (size_t)(slot * p.cpb + cb) * p.halo_bytes), &mapA, &halo_full[slot], cb * BKE, -p.g.pw,
                                        th0 - p.g.ph, tn0);
                    }
                    __syncwarp();
                }
                int kh = 0, kw = 0, cb = 0;   // filter tap / channel block of the current k-block (conv K-major A)
                const int a_simple = (a_tma && !halo_mode) ? (p.a_kind == OP_TMA_K ? 1 : (p.a_kind == OP_TMA_CONV_K ? 2 : 3)) : 0;
                const int b_simple = b_tma ? (p.b_kind == OP_TMA_K ? 1 : 3) : 0;
                const int sgn = p.ct_dgrad ? -1 : 1;
                const int th0s = th0 * p.ct_s;
                const int kw_wrap = p.gen_taps ? 0x7fffffff : p.g.KW;
                int kcoord = ti.kb_begin * BKE;
                for (int i = 0; i < ti.nkb; ++i) {
                    mbar_wait(&empty_bar[st], ph ^ 1u);
                    if (elect_one()) {
                        const uint32_t a_dst = ring0 + (uint32_t)st * (uint32_t)stage_bytes;
                        const uint32_t b_dst = a_dst + (uint32_t)a_bytes;
                        uint64_t* bar = &full_bar[st];
                        AVEC_TSK(j, i, 200, 0);
                        if (AVEC_DBG_MODE(1)) {
                            mbar_arrive(bar);
                        } else {
                            mbar_expect_tx(bar, tx);
                            if (a_simple == 1) tma_load_2d(a_dst, &mapA, bar, kcoord, ti.m0 - p.dbg_rowofs);
                            else if (a_simple == 2 && p.gen_taps) tma_load_4d(a_dst, &mapA, bar, cb * BKE, (int)p.tap_dw[kw], th0s + (int)p.tap_dh[kw], tn0);
                            else if (a_simple == 2) tma_load_4d(a_dst, &mapA, bar, cb * BKE, sgn * (kw - p.g.pw), th0s + sgn * (kh - p.g.ph), tn0);
                            else if (a_simple == 3) tma_fill(p, p.a_kind, &mapA, smem + (size_t)st * stage_bytes, bar, p.a_rows, p.a_group_stride, ti.m0, ti.kb_begin + i, ti.mtile, true);
                            if (b_simple == 1 && p.gen_taps) tma_load_2d(b_dst, &mapB, bar, p.tap_koff[kw] + cb * BKE, ti.n0);
                            else if (b_simple == 1) tma_load_2d(b_dst, &mapB, bar, kcoord, ti.n0);
                            else if (b_simple == 3) tma_fill(p, p.b_kind, &mapB, smem + (size_t)st * stage_bytes + a_bytes, bar, p.b_rows, p.b_group_stride, ti.n0, ti.kb_begin + i, ti.mtile, false);
                        }
                        AVEC_TSK(j, i, 200, 1);
                    }
                    __syncwarp();
                    kcoord += BKE;
                    if (++cb == p.cpb) { cb = 0; if (++kw == kw_wrap) { kw = 0; ++kh; } }   // table mode: kw = running tap index
                    if (++st == p.stages) { st = 0; ph ^= 1u; }
                }
            }
        }
    } else {
        // ===================== MMA issuer =====================
        // One elected thread issues everything; per k-block it executes a few dozen scalar instructions (running stage
        // address / barrier / phase, 32-bit descriptor arithmetic: the 14-bit address field never carries out of the low
        // word), because at 64-column tiles a k-block is only 128 tensor-pipe cycles.
        const int a_mn = is_mn(p.a_kind) ? 1 : 0;
        const int b_mn = is_mn(p.b_kind) ? 1 : 0;
        const uint32_t idesc = make_idesc(BN, a_mn, b_mn);
        const bool halo_mode = p.a_kind == OP_TMA_CONV_HALO;
        const uint64_t a_desc0 = a_mn ? make_smem_desc(0, p.a_group_stride, 1024) : make_smem_desc(0, 16, 1024);
        const uint64_t b_desc0 = b_mn ? make_smem_desc(0, p.b_group_stride, 1024) : make_smem_desc(0, 16, 1024);
        const uint32_t a_hi = (uint32_t)(a_desc0 >> 32), b_hi = (uint32_t)(b_desc0 >> 32);
        const uint32_t a_lo0 = (uint32_t)a_desc0, b_lo0 = (uint32_t)b_desc0;
        const uint32_t a_kstep = a_mn ? (2048u >> 4) : (32u >> 4), b_kstep = b_mn ? (2048u >> 4) : (32u >> 4);
        const uint32_t ring0 = smem_u32(smem) >> 4, stage_step = (uint32_t)stage_bytes >> 4, a_bytes16 = (uint32_t)a_bytes >> 4;
        const uint32_t rowofs16 = ((uint32_t)p.dbg_rowofs * 128u) >> 4;
        const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
        const int ksteps = p.ksteps;
        int j = 0, st = 0;
        uint32_t ph = 0;
        uint32_t stage16 = ring0;
        for (int t = blockIdx.x; t < tiles_total; t += gridDim.x, ++j) {
            const TileInfo ti = decode_tile(p, t);
            const int buf = j & 1;
            mbar_wait(&accum_empty[buf], (uint32_t)(((j >> 1) & 1) ^ 1));   // epilogue has drained this accumulator buffer
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
            if (halo_mode) { mbar_wait(&halo_full[buf], (uint32_t)((j >> 1) & 1)); tc_fence_after(); }
            int kh = 0, kw = 0, cb = 0;
            const int nkb = ti.nkb;
            for (int i = 0; i < nkb; ++i) {
                mbar_wait_addr(full0 + 8u * (uint32_t)st, ph);
                tc_fence_after();
                const bool leader = elect_one();
                if (leader && AVEC_DBG_MODE(2)) {
                    if (i == 0) AVEC_TSJ(j, 3);
                    mbar_arrive(&empty_bar[st]);
                    if (i == nkb - 1) { mbar_arrive(&accum_full[buf]); AVEC_TSJ(j, 4); }
                } else if (leader) {
                    AVEC_TSK(j, i, 168, 0);
                    if (i == 0 && j == 0) AVEC_TS(2);   // first k-block landed in shared memory
                    if (i == 0) AVEC_TSJ(j, 3);
                    uint32_t alo = a_lo0 + stage16 + rowofs16;
                    if (halo_mode) {
                        // tap (kh, kw) of 64-channel block cb reads the resident halo tile from row kh*(W+2)+kw on
                        // (dgrad: the mirrored tap); a 128-byte row offset keeps the 128B-swizzle phase consistent
                        const int th = p.ct_dgrad ? p.g.KH - 1 - kh : kh, tw = p.ct_dgrad ? p.g.KW - 1 - kw : kw;
                        alo = a_lo0 + ((smem_u32(halo + (size_t)(buf * p.cpb + cb) * p.halo_bytes) + (uint32_t)(th * p.halo_W2 + tw) * 128u) >> 4);
                    }
                    uint32_t blo = b_lo0 + stage16 + a_bytes16;
                    if (ksteps == 4) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            umma_f16(d_tmem, ((uint64_t)a_hi << 32) | alo, ((uint64_t)b_hi << 32) | blo, idesc, (i > 0 || k > 0) ? 1u : 0u);
                            alo += a_kstep; blo += b_kstep;
                        }
                    } else {
                        for (int k = 0; k < ksteps; ++k) {
                            umma_f16(d_tmem, ((uint64_t)a_hi << 32) | alo, ((uint64_t)b_hi << 32) | blo, idesc, (i > 0 || k > 0) ? 1u : 0u);
                            alo += a_kstep; blo += b_kstep;
                        }
                    }
                    umma_commit_addr(empty0 + 8u * (uint32_t)st);
                    if (i == nkb - 1) {   // last MMA of the tile issued
                        umma_commit(&accum_full[buf]);
                        if (halo_mode) umma_commit(&halo_empty[buf]);
                        if (j == 0) AVEC_TS(3);
                        AVEC_TSJ(j, 4);
                    }
                    AVEC_TSK(j, i, 168, 1);
                }
                __syncwarp();
                if (halo_mode) { if (++cb == p.cpb) { cb = 0; if (++kw == p.g.KW) { kw = 0; ++kh; } } }
                stage16 += stage_step;
                if (++st == p.stages) { st = 0; ph ^= 1u; stage16 = ring0; }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 4) { tc_fence_after(); tmem_dealloc(tmem_base, ncols); if (lane == 0) AVEC_TS(6); }
}

// =====================================================================================================================
// Visual stem: Conv3d(1 -> 64, k (5,7,7), s (1,2,2), "same") on single-channel video as a direct tcgen05 implicit GEMM.
// (reference nnet/networks.py:459-471, layers.Conv3d).  With C = 1 an im2col row is 245 scattered bf16 values, which a TMA
// box cannot fetch and which costs 6.4 GB of HBM when materialised.  Here the im2col tile is built IN SHARED MEMORY:
//   * builder warps stage the raw input window of a tile (5 frames x 13 rows x 88 pixels, 13.5 KB, cp.async, double
//     buffered, zero padding applied once) and expand it into the canonical no-swizzle UMMA layout: one 16-byte chunk per
//     (site, kt, kh) = the 7 kw taps + 1 pad tap, read as five aligned words and funnel-shifted (window origin 2wo-3 is
//     odd); chunk block c = kt*7+kh holds [128 sites][16 B], so K-adjacent core matrices are 2048 B apart (LBO) and
//     8-site groups 128 B apart (SBO).  K = 36 chunks x 8 = 288 (245 real taps; the weights of pad taps are zero);
//   * the same tile is the K-major A operand of the forward (M = sites) and the MN-major B operand of the weight gradient
//     (N = taps, K = sites) - no transposed copy;
//   * weights (forward) stay resident in shared memory for the whole persistent CTA; the forward epilogue is the GEMM
//     kernel's fast epilogue (bias, bf16 store, BatchNorm column statistics).
constexpr int ST_CHUNKS = 36, ST_HALF = 18;
constexpr int ST_WROW = 104;                       // window row: 8 zero | <= 88 pixels | zeros
constexpr int ST_WROWS = 13;                       // input rows a 128-site tile can touch: 2*3 + 7
constexpr int ST_WIN_BYTES = 13568;                // 5 * 13 * 104 * 2 = 13520, rounded up to 128
constexpr int ST_SLOT_BYTES = ST_HALF * 2048;      // 36864: half an im2col tile (18 chunk blocks of [128][16 B])
constexpr int ST_SLOTS = 3;
constexpr int ST_KPAD = 320;                       // forward weight rows: 5 k-blocks of 64 (k = (kt*7+kh)*8 + kw)
constexpr int ST_B_BYTES = 5 * 64 * 128;           // 40960
constexpr int ST_BUILDERS = 256;             // two groups of 128 (thread = site), one per half tile

struct StemParams {
    const bf16* x;            // [Nb][T][H][W] bf16
    int Nb, T, H, W, Ho, Wo;
    int tpf;                  // tiles per frame = cdiv(Ho*Wo, 128)
    int total_tiles;
    TcParams tc;              // forward: BN = N = 64 + epilogue parameters
    float* dw;                // weight gradient [64][245] fp32 (atomic accumulation)
    int dbg;                  // ablation bits (AVEC_STEM_DBG): 1 no im2col expansion, 2 no MMAs, 4 no epilogue body, 8 no window copies
};

__device__ __forceinline__ uint64_t make_smem_desc_ns(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (sm_100); layout type 0 = no swizzle
    return d;
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t r;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(r) : "r"(a) : "memory");
    return r;
}
__device__ __forceinline__ void sts128u(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

struct StemTile { int n, t, s0, ho0; long long frame; };
__device__ __forceinline__ StemTile stem_tile(const StemParams& p, int tile) {
    StemTile st;
    st.frame = tile / p.tpf;
    const int k = tile - (int)st.frame * p.tpf;
    st.n = (int)(st.frame / p.T); st.t = (int)(st.frame - (long long)st.n * p.T);
    st.s0 = k * 128; st.ho0 = st.s0 / p.Wo;
    return st;
}
// expand one half (18 chunk blocks) of the im2col tile of this thread's site from the staged window
template <int HALF>
__device__ __forceinline__ void stem_build_half(uint32_t slot, uint32_t wsite, int site) {
#pragma unroll
    for (int cc = 0; cc < ST_HALF; ++cc) {
        const int chunk = HALF * ST_HALF + cc;
        const uint32_t dst = slot + (uint32_t)cc * 2048u + (uint32_t)site * 16u;
        if (chunk >= 35) { sts128u(dst, 0u, 0u, 0u, 0u); continue; }
        const int kt = chunk / 7, kh = chunk % 7;
        const uint32_t a = wsite + (uint32_t)((kt * ST_WROWS + kh) * (ST_WROW / 2)) * 4u;
        const uint32_t w0 = lds32(a), w1 = lds32(a + 4), w2 = lds32(a + 8), w3 = lds32(a + 12), w4 = lds32(a + 16);
        sts128u(dst, __funnelshift_r(w0, w1, 16), __funnelshift_r(w1, w2, 16), __funnelshift_r(w2, w3, 16), __funnelshift_r(w3, w4, 16));
    }
}
// Builder warps (256 threads): window prefetch + im2col expansion for every tile of this CTA.  The two groups of 128 threads
// expand the two halves of a tile concurrently (thread = site), each into its own ring slot; a single warp per scheduler
// issues one dependent instruction every ~4 cycles, so the per-tile instruction count per thread is what bounds this loop:
// window copies use per-thread offsets that are tile-invariant (computed once), 3 x 16 bytes per thread and tile.
__device__ __forceinline__ void stem_builder_loop(const StemParams& p, uint8_t* a_ring, uint8_t* wins, uint64_t* a_full, uint64_t* a_empty, int btid) {
    constexpr int SEGS = 11, N_COPY = 5 * ST_WROWS * SEGS, CPT = (N_COPY + ST_BUILDERS - 1) / ST_BUILDERS;   // 715 copies, 3 per thread
    for (int i = btid; i < 2 * ST_WIN_BYTES / 16; i += ST_BUILDERS) reinterpret_cast<uint4*>(wins)[i] = make_uint4(0u, 0u, 0u, 0u);
    asm volatile("bar.sync 2, 256;" ::: "memory");
    const uint32_t win0 = smem_u32(wins), ring0 = smem_u32(a_ring);
    const int grp = btid >> 7, site = btid & 127;
    int c_kt[CPT], c_r[CPT], c_src[CPT];
    uint32_t c_dst[CPT];
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
        const int idx = btid + k * ST_BUILDERS;
        const int seg = idx % SEGS, rr = idx / SEGS;
        c_r[k] = rr % ST_WROWS; c_kt[k] = idx < N_COPY ? rr / ST_WROWS : -100;   // -100: no copy (frame test fails)
        c_src[k] = (c_kt[k] * p.H + c_r[k]) * p.W + seg * 8;
        c_dst[k] = (uint32_t)(((rr / ST_WROWS * ST_WROWS + c_r[k]) * ST_WROW + 8 + seg * 8) * 2);
    }
    auto issue_window = [&](int tile, uint32_t win) {
        const StemTile st = stem_tile(p, tile);
        const int hi_base = 2 * st.ho0 - 3;
        const long long base = (((long long)st.n * p.T + st.t - 2) * p.H + hi_base) * p.W;
        if (!(p.dbg & 8)) {
#pragma unroll
            for (int k = 0; k < CPT; ++k) {
                const int ti = st.t - 2 + c_kt[k], hi = hi_base + c_r[k];
                const bool inside = c_kt[k] >= 0;
                const bool ok = inside && (unsigned)ti < (unsigned)p.T && (unsigned)hi < (unsigned)p.H;
                if (inside) cp_async16(win + c_dst[k], ok ? p.x + base + c_src[k] : p.x, ok ? 16 : 0);
            }
        }
        cp_async_commit();
    };
    int j = 0;
    int it = grp;   // half-tile counter of this group: slot = it % 3, use parity = (it / 3) & 1
    if ((int)blockIdx.x < p.total_tiles) issue_window(blockIdx.x, win0);
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++j, it += 2) {
        const int tn = t + gridDim.x;
        if (tn < p.total_tiles) issue_window(tn, win0 + (uint32_t)(((j + 1) & 1) * ST_WIN_BYTES));
        else cp_async_commit();
        cp_async_wait<1>();
        asm volatile("bar.sync 2, 256;" ::: "memory");
        const StemTile st = stem_tile(p, t);
        const int s = st.s0 + site;
        const int ho = s / p.Wo, wo = s - ho * p.Wo;
        // word address of element (row 2*(ho-ho0), column 2*wo + 4) of frame slice 0
        const uint32_t wsite = win0 + (uint32_t)((j & 1) * ST_WIN_BYTES) + (uint32_t)((2 * (ho - st.ho0)) * (ST_WROW / 2) + wo + 2) * 4u;
        const int slot = it % ST_SLOTS;
        mbar_wait(&a_empty[slot], (uint32_t)(((it / ST_SLOTS) & 1) ^ 1));
        const uint32_t sl = ring0 + (uint32_t)slot * ST_SLOT_BYTES;
        if (!(p.dbg & 1)) { if (grp == 0) stem_build_half<0>(sl, wsite, site); else stem_build_half<1>(sl, wsite, site); }
        fence_proxy_async();
        mbar_arrive(&a_full[slot]);
        asm volatile("bar.sync 2, 256;" ::: "memory");
    }
}

// ---- forward: warps 0-3 epilogue, warp 4 MMA issuer (+ TMEM, weight TMA), warps 5-12 builders
constexpr int ST_FWD_THREADS = 416;
constexpr size_t ST_FWD_SMEM = 1024 + ST_B_BYTES + ST_SLOTS * ST_SLOT_BYTES + 2 * ST_WIN_BYTES + STG_BYTES + 1024;

__global__ void __launch_bounds__(ST_FWD_THREADS, 1) stem3d_fwd_kernel(const __grid_constant__ StemParams p, const __grid_constant__ CUtensorMap mapW) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* b_s = smem;
    uint8_t* a_ring = b_s + ST_B_BYTES;
    uint8_t* wins = a_ring + ST_SLOTS * ST_SLOT_BYTES;
    float* stg_all = reinterpret_cast<float*>(wins + 2 * ST_WIN_BYTES);
    uint8_t* ctrl = reinterpret_cast<uint8_t*>(stg_all) + STG_BYTES;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(ctrl);        // [3]
    uint64_t* a_empty = a_full + 4;                               // [3]
    uint64_t* accum_full = a_empty + 4;                           // [2]
    uint64_t* accum_empty = accum_full + 2;                       // [2]
    uint64_t* b_full = accum_empty + 2;                           // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_full + 2);
    float* bias_s = reinterpret_cast<float*>(ctrl + 256);         // [64]
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    if (tid == 0) {
        for (int i = 0; i < ST_SLOTS; ++i) { mbar_init(&a_full[i], 128); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&accum_full[i], 1); mbar_init(&accum_empty[i], PRODUCER_THREADS); }
        mbar_init(b_full, 1);
        fence_barrier_init();
        tma_prefetch_desc(&mapW);
    }
    if (warp == 4) tmem_alloc(tmem_slot, 128);
    for (int c = tid; c < 64; c += ST_FWD_THREADS) bias_s[c] = p.tc.ep.bias ? p.tc.ep.bias[c] : 0.0f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        // ===================== epilogue =====================
        int j = 0;
        EpiParams ep = p.tc.ep;
        ep.bias = nullptr;
        const uint32_t stg_s = smem_u32(stg_all + warp * STG_WARP), bias_sa = smem_u32(bias_s);
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++j) {
            const StemTile st = stem_tile(p, t);
            TileInfo ti;
            ti.mtile = t; ti.n0 = 0; ti.z = 0; ti.kb_begin = 0; ti.nkb = 0; ti.m0 = 0;
            ti.row_base = st.frame * (long long)(p.Ho * p.Wo) + st.s0;
            ti.rows_valid = min(128, p.Ho * p.Wo - st.s0);
            const int buf = j & 1;
            mbar_wait(&accum_full[buf], (uint32_t)((j >> 1) & 1));
            tc_fence_after();
            const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * 64);
            float* stats_dst = ep.colstats ? ep.colstats + (size_t)(t % AVEC_STATS_REPLICAS) * 2 * 64 : nullptr;
            if (p.dbg & 4) { }
            else if (ep.colstats) epilogue_fast<AVEC_EPI_LINEAR, true, true>(p.tc, ep, ti, lane_addr, stg_s, bias_sa, stats_dst, warp, lane, p.tc.ep.bias != nullptr);
            else epilogue_fast<AVEC_EPI_LINEAR, false, true>(p.tc, ep, ti, lane_addr, stg_s, bias_sa, stats_dst, warp, lane, p.tc.ep.bias != nullptr);
            tc_fence_before();
            mbar_arrive(&accum_empty[buf]);
        }
        tc_fence_before();
    } else if (warp == 4) {
        // ===================== weight load + MMA issuer =====================
        if (elect_one()) {
            mbar_expect_tx(b_full, ST_B_BYTES);
            for (int kb = 0; kb < 5; ++kb) tma_load_2d(smem_u32(b_s) + (uint32_t)kb * 8192u, &mapW, b_full, kb * 64, 0);
        }
        __syncwarp();
        mbar_wait(b_full, 0);
        tc_fence_after();
        const uint32_t idesc = make_idesc(64, 0, 0);
        const uint64_t a_desc0 = make_smem_desc_ns(0, 2048, 128);
        const uint64_t b_desc0 = make_smem_desc(0, 16, 1024);
        const uint32_t a_hi = (uint32_t)(a_desc0 >> 32), b_hi = (uint32_t)(b_desc0 >> 32);
        const uint32_t ring16 = smem_u32(a_ring) >> 4, b16 = smem_u32(b_s) >> 4;
        int j = 0, slot = 0;
        uint32_t round = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++j) {
            const int buf = j & 1;
            mbar_wait(&accum_empty[buf], (uint32_t)(((j >> 1) & 1) ^ 1));
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(buf * 64);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                mbar_wait(&a_full[slot], round);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t alo0 = (uint32_t)a_desc0 + ring16 + (uint32_t)slot * (ST_SLOT_BYTES >> 4);
#pragma unroll
                    for (int sI = 0; sI < 9; ++sI) {
                        const int sg = half * 9 + sI;
                        const uint32_t alo = alo0 + (uint32_t)sI * (4096u >> 4);
                        const uint32_t blo = (uint32_t)b_desc0 + b16 + (uint32_t)(sg >> 2) * (8192u >> 4) + (uint32_t)(sg & 3) * (32u >> 4);
                        if (!(p.dbg & 2)) umma_f16(d_tmem, ((uint64_t)a_hi << 32) | alo, ((uint64_t)b_hi << 32) | blo, idesc, sg > 0 ? 1u : 0u);
                    }
                    umma_commit(&a_empty[slot]);
                    if (half == 1) umma_commit(&accum_full[buf]);
                }
                __syncwarp();
                if (++slot == ST_SLOTS) { slot = 0; round ^= 1u; }
            }
        }
        tc_fence_before();
    } else {
        stem_builder_loop(p, a_ring, wins, a_full, a_empty, tid - 160);
    }
    __syncthreads();
    if (warp == 4) { tc_fence_after(); tmem_dealloc(tmem_base, 128); }
}

// ---- weight gradient: dW[co][tap] = sum_sites dY[site][co] * col[site][tap].  A = dY tile (MN-major, 128B swizzle, TMA; the
// second 64-wide M group is a zero block so a full M = 128 instruction can be used), B = the im2col tile (MN-major, no
// swizzle: N = taps, 144 per half tile), two accumulators of 144 columns that live in TMEM for the CTA's whole tile range.
// warps 0-3 final epilogue, warp 4 MMA issuer (+ TMEM), warp 5 dY TMA, warps 6-13 builders
constexpr int ST_WG_THREADS = 448;
constexpr int ST_Y_BYTES = 2 * 128 * 128;   // dY stage: group 0 (TMA) + group 1 (zeros)
constexpr size_t ST_WG_SMEM = 1024 + 2 * ST_Y_BYTES + ST_SLOTS * ST_SLOT_BYTES + 2 * ST_WIN_BYTES + 1024;

__global__ void __launch_bounds__(ST_WG_THREADS, 1) stem3d_wgrad_kernel(const __grid_constant__ StemParams p, const __grid_constant__ CUtensorMap mapY) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* y_s = smem;                                   // [2 stages][2 groups][128 sites][128 B]
    uint8_t* a_ring = y_s + 2 * ST_Y_BYTES;
    uint8_t* wins = a_ring + ST_SLOTS * ST_SLOT_BYTES;
    uint8_t* ctrl = wins + 2 * ST_WIN_BYTES;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(ctrl);
    uint64_t* a_empty = a_full + 4;
    uint64_t* y_full = a_empty + 4;      // [2]
    uint64_t* y_empty = y_full + 2;      // [2]
    uint64_t* done = y_empty + 2;        // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 2);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    if (tid == 0) {
        for (int i = 0; i < ST_SLOTS; ++i) { mbar_init(&a_full[i], 128); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&y_full[i], 1); mbar_init(&y_empty[i], 1); }
        mbar_init(done, 1);
        fence_barrier_init();
        tma_prefetch_desc(&mapY);
    }
    if (warp == 4) tmem_alloc(tmem_slot, 512);
    // zero M groups (generic proxy writes, then handed to the async proxy)
    for (int i = tid; i < 2 * ST_Y_BYTES / 16; i += ST_WG_THREADS) reinterpret_cast<uint4*>(y_s)[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const bool any_tile = (int)blockIdx.x < p.total_tiles;

    if (warp < 4) {
        // ===================== final epilogue: rows = co (lanes 0-63), columns = (chunk, kw) =====================
        if (any_tile && warp < 2) {
            mbar_wait(done, 0);
            tc_fence_after();
            const int co = warp * 32 + lane;
            for (int c0 = 0; c0 < 288; c0 += 16) {
                float v[16];
                tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int n = c0 + i, chunk = n >> 3, kw = n & 7;
                    if (kw < 7 && chunk < 35) atomicAdd(p.dw + (size_t)co * 245 + chunk * 7 + kw, v[i]);
                }
            }
        }
        tc_fence_before();
    } else if (warp == 4) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = make_idesc(144, 1, 1);
        const uint64_t a_desc0 = make_smem_desc(0, 128 * 128, 1024);       // MN-major, 128B swizzle: LBO = 64-wide group stride
        const uint64_t b_desc0 = make_smem_desc_ns(0, 128, 2048);          // MN-major, no swizzle: LBO = next 8 sites, SBO = next 8 taps
        const uint32_t a_hi = (uint32_t)(a_desc0 >> 32), b_hi = (uint32_t)(b_desc0 >> 32);
        const uint32_t ring16 = smem_u32(a_ring) >> 4, y16 = smem_u32(y_s) >> 4;
        int j = 0, slot = 0;
        uint32_t round = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++j) {
            const int ys = j & 1;
            mbar_wait(&y_full[ys], (uint32_t)((j >> 1) & 1));
            tc_fence_after();
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                mbar_wait(&a_full[slot], round);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t d_tmem = tmem_base + (uint32_t)(half * 144);
                    const uint32_t alo0 = (uint32_t)a_desc0 + y16 + (uint32_t)ys * (ST_Y_BYTES >> 4);
                    const uint32_t blo0 = (uint32_t)b_desc0 + ring16 + (uint32_t)slot * (ST_SLOT_BYTES >> 4);
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks)
                        if (!(p.dbg & 2)) umma_f16(d_tmem, ((uint64_t)a_hi << 32) | (alo0 + (uint32_t)ks * (2048u >> 4)), ((uint64_t)b_hi << 32) | (blo0 + (uint32_t)ks * (256u >> 4)),
                                 idesc, (j > 0 || ks > 0) ? 1u : 0u);
                    umma_commit(&a_empty[slot]);
                    if (half == 1) umma_commit(&y_empty[ys]);
                }
                __syncwarp();
                if (++slot == ST_SLOTS) { slot = 0; round ^= 1u; }
            }
        }
        if (any_tile && elect_one()) umma_commit(done);
        __syncwarp();
        tc_fence_before();
    } else if (warp == 5) {
        // ===================== dY TMA producer =====================
        int j = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++j) {
            const int ys = j & 1;
            mbar_wait(&y_empty[ys], (uint32_t)(((j >> 1) & 1) ^ 1));
            if (elect_one()) {
                const StemTile st = stem_tile(p, t);
                mbar_expect_tx(&y_full[ys], 128 * 128);
                tma_load_3d(smem_u32(y_s) + (uint32_t)ys * ST_Y_BYTES, &mapY, &y_full[ys], 0, st.s0, (int)st.frame);
            }
            __syncwarp();
        }
    } else {
        stem_builder_loop(p, a_ring, wins, a_full, a_empty, tid - 192);
    }
    __syncthreads();
    if (warp == 4) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// =====================================================================================================================
// Weight gradient of the 64 -> 64 channel 3x3 stride-1 convolutions (ResNet stage 1: 3.1 M sites at B = 64, the largest
// reduction of the model).  As a generic 128 x 128-tile GEMM this launch re-reads X once per filter tap and dY once per column
// block from L2 (1.3 ms, L2 -> SMEM bound).  Here one persistent CTA keeps the COMPLETE dW (64 x 576 fp32) in tensor memory and
// streams every site exactly once: per tile of BH image rows a TMA box of dY on the (W+2)-wide halo grid (columns W, W+1 and
// rows past the image are zero-filled by the TMA unit) and ONE halo tile of X ((BH+2) x (W+2) sites x 64 channels).  Filter
// tap (kh, kw) is the halo tile read from row kh*(W+2)+kw on (128-byte row offsets keep the 128B-swizzle phase), so the 9
// taps cost no extra traffic.  MMA shape: A = X halo, MN-major, M = 128 = two taps (kh, kh+1) x 64 input channels (the second
// 64-wide group starts (W+2) rows further: LBO = (W+2)*128 bytes, a multiple of the 1024-byte swizzle atom because
// (W+2) % 8 == 0), B = dY, MN-major, N = 64 output channels, K = 128 halo sites per tile; six accumulators of 64 columns:
// (kh 0|1, kw 0..2) and (kh 1|2, kw 0..2) - the kh = 1 taps are computed twice so that no operand ever reads past the tile.
// warps 0-3: final flush (fp32 atomics into dW[co][tap*64+ci]); warp 4: MMA issuer + TMEM; warp 5: TMA producer.
constexpr int WH_THREADS = 192;
constexpr int WH_Y_BYTES = 128 * 128;      // dY stage: 128 K rows (rows >= BH*(W+2) stay zero)
constexpr int WH_X_BYTES = 184 * 128;      // X halo stage: <= 168 rows from TMA + zero tail read by the last taps' K padding
constexpr int WH_STAGE_BYTES = 40960;
constexpr int WH_STAGES = 4;
constexpr size_t WH_SMEM = 1024 + (size_t)WH_STAGES * WH_STAGE_BYTES + 1024;

struct WgHaloParams {
    int N, H, W, W2, BH, tph, total_tiles;
    int y_tx, x_tx;
    float alpha;
    float* dw;    // [64][576] fp32
};

__global__ void __launch_bounds__(WH_THREADS, 1) wgrad_halo64_kernel(const __grid_constant__ WgHaloParams p, const __grid_constant__ CUtensorMap mapY,
                                                                    const __grid_constant__ CUtensorMap mapX) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* ctrl = smem + (size_t)WH_STAGES * WH_STAGE_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(ctrl);   // [4]
    uint64_t* empty = full + WH_STAGES;                     // [4]
    uint64_t* done = empty + WH_STAGES;                     // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 2);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    if (tid == 0) {
        for (int i = 0; i < WH_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 2); }   // two MMA-issuing threads release a stage
        mbar_init(done, 2);
        fence_barrier_init();
        tma_prefetch_desc(&mapY);
        tma_prefetch_desc(&mapX);
    }
    if (warp == 4) tmem_alloc(tmem_slot, 512);
    // rows the TMA boxes never write must read as zeros (K padding of dY, tail of the halo tile)
    for (int i = tid; i < WH_STAGES * WH_STAGE_BYTES / 16; i += WH_THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const bool any_tile = (int)blockIdx.x < p.total_tiles;
    // ===================== MMA issue loop for accumulators [q_lo, q_lo + 3): two issuing threads (warp 4 and warp 0) share the
    // six accumulators - one thread needs ~36 cycles of scalar work per 32-cycle 128x64x16 MMA =====================
    auto issue_loop = [&](int q_lo) {
        const uint32_t idesc = make_idesc(64, 1, 1);
        const uint64_t a_desc0 = make_smem_desc(0, (uint32_t)p.W2 * 128u, 1024);   // X halo: second 64-row M group = next filter row
        const uint64_t b_desc0 = make_smem_desc(0, WH_Y_BYTES, 1024);              // dY: single 64-wide N group
        const uint32_t a_hi = (uint32_t)(a_desc0 >> 32), b_hi = (uint32_t)(b_desc0 >> 32);
        const uint32_t base16 = smem_u32(smem) >> 4;
        int j = 0, st = 0;
        uint32_t ph = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++j) {
            mbar_wait(&full[st], ph);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t y16 = base16 + (uint32_t)st * (WH_STAGE_BYTES >> 4);
                const uint32_t x16 = y16 + (WH_Y_BYTES >> 4);
#pragma unroll
                for (int qq = 0; qq < 3; ++qq) {
                    const int q = q_lo + qq;
                    const int set = q / 3, kw = q - set * 3;
                    const uint32_t a0 = (uint32_t)a_desc0 + x16 + (uint32_t)((set * p.W2 + kw) * 128 >> 4);
                    const uint32_t d_tmem = tmem_base + (uint32_t)(q * 64);
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks)
                        umma_f16(d_tmem, ((uint64_t)a_hi << 32) | (a0 + (uint32_t)ks * (2048u >> 4)),
                                 ((uint64_t)b_hi << 32) | ((uint32_t)b_desc0 + y16 + (uint32_t)ks * (2048u >> 4)), idesc, (j > 0 || ks > 0) ? 1u : 0u);
                }
                umma_commit(&empty[st]);
            }
            __syncwarp();
            if (++st == WH_STAGES) { st = 0; ph ^= 1u; }
        }
        if (any_tile && elect_one()) umma_commit(done);
        __syncwarp();
        tc_fence_before();
    };


    if (warp < 4) {
        // ===================== final flush: lane = (tap row g, input channel c), columns = output channels =====================
        if (warp == 0) { issue_loop(3); tc_fence_after(); }
        if (any_tile) {
            mbar_wait(done, 0);
            tc_fence_after();
            const int row = warp * 32 + lane, g = row >> 6, c = row & 63;
            for (int q = 0; q < 6; ++q) {
                const int set = q / 3, kw = q - set * 3, kh = set + g;
                if (set == 1 && g == 0) continue;   // kh = 1 duplicate (warp-uniform: g is fixed per warp)
                float* dst = p.dw + (size_t)(kh * 3 + kw) * 64 + c;
                for (int c0 = 0; c0 < 64; c0 += 16) {
                    float v[16];
                    tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(q * 64 + c0), v);
#pragma unroll
                    for (int i = 0; i < 16; ++i) atomicAdd(dst + (size_t)(c0 + i) * 576, p.alpha * v[i]);
                }
            }
        }
        tc_fence_before();
    } else if (warp == 4) {
        issue_loop(0);
    } else {
        // ===================== TMA producer =====================
        int st = 0;
        uint32_t ph = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
            mbar_wait(&empty[st], ph ^ 1u);
            if (elect_one()) {
                const int n = t / p.tph, h0 = (t - n * p.tph) * p.BH;
                const uint32_t ydst = smem_u32(smem) + (uint32_t)st * WH_STAGE_BYTES;
                mbar_expect_tx(&full[st], (uint32_t)(p.y_tx + p.x_tx));
                tma_load_4d(ydst, &mapY, &full[st], 0, 0, h0, n);
                tma_load_4d(ydst + WH_Y_BYTES, &mapX, &full[st], 0, -1, h0 - 1, n);
            }
            __syncwarp();
            if (++st == WH_STAGES) { st = 0; ph ^= 1u; }
        }
    }
    __syncthreads();
    if (warp == 4) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// =====================================================================================================================
// Forward / data gradient of the 64 -> 64 channel 3x3 stride-1 convolutions (ResNet stage 1) on halo tiles.  The generic
// kernel loads one shifted copy of the input tile per filter tap plus the tap's weights per k-block (198 KB from L2 per
// 110 output sites - L2 -> SMEM bound) and pays one producer/consumer handshake per 128 tensor-pipe cycles.  Here the 72 KB
// of weights stay resident in shared memory for the whole persistent CTA, ONE halo tile ((BH+2) x (W+2) sites x 64 ch,
// 21 KB) is loaded per output tile and all 36 MMAs of the tile (9 taps x 4 K-steps; tap = 128-byte row offset into the halo
// tile) are issued behind a single barrier wait.  Accumulator rows live on the (W+2)-wide halo grid; the fast epilogue maps
// them to output rows through a small table and skips the two pad columns.  Two epilogue warpgroups alternate tiles (two
// TMEM accumulators each, so the MMA warp can run two tiles ahead of either group), because at 64 columns a tile is only ~1150 tensor-pipe cycles - about one epilogue.
// warps 0-3 / 4-7: epilogue groups, warp 8: MMA issuer + TMEM + weight TMA, warp 9: halo TMA producer
constexpr int CH_THREADS = 320;
constexpr int CH_W_BYTES = 9 * 64 * 128;      // 73728
constexpr int CH_X_BYTES = 184 * 128;         // halo stage (168 rows from TMA; rows up to 177 are read by discarded accumulator rows)
constexpr int CH_STAGES = 3;
constexpr size_t CH_SMEM = 1024 + CH_W_BYTES + (size_t)CH_STAGES * CH_X_BYTES + 2 * STG_BYTES + 1024;

struct ConvHaloParams {
    int N, H, W, W2, BH, tph, total_tiles, x_tx, dgrad;
    TcParams tc;   // BN = N = 64, epilogue parameters, rm_on = 1
};

__global__ void __launch_bounds__(CH_THREADS, 1) conv3x3_halo64_kernel(const __grid_constant__ ConvHaloParams p, const __grid_constant__ CUtensorMap mapX,
                                                                      const __grid_constant__ CUtensorMap mapW) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* w_s = smem;
    uint8_t* x_s = w_s + CH_W_BYTES;
    float* stg_all = reinterpret_cast<float*>(x_s + CH_STAGES * CH_X_BYTES);
    uint8_t* ctrl = reinterpret_cast<uint8_t*>(stg_all) + 2 * STG_BYTES;
    uint64_t* x_full = reinterpret_cast<uint64_t*>(ctrl);     // [3]
    uint64_t* x_empty = x_full + 4;                             // [3]
    uint64_t* accum_full = x_empty + 4;                         // [4]: two TMEM accumulators per epilogue warpgroup
    uint64_t* accum_empty = accum_full + 4;                     // [4]
    uint64_t* w_full = accum_empty + 4;                         // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 2);
    float* bias_s = reinterpret_cast<float*>(ctrl + 192);       // [64]
    int* rowrel = reinterpret_cast<int*>(ctrl + 192 + 256);     // [128]
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    if (tid == 0) {
        for (int i = 0; i < CH_STAGES; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1); }
        for (int i = 0; i < 4; ++i) { mbar_init(&accum_full[i], 1); mbar_init(&accum_empty[i], 128); }
        mbar_init(w_full, 1);
        fence_barrier_init();
        tma_prefetch_desc(&mapX);
        tma_prefetch_desc(&mapW);
    }
    if (warp == 8) tmem_alloc(tmem_slot, 256);
    if (tid < 64) bias_s[tid] = p.tc.ep.bias ? p.tc.ep.bias[tid] : 0.0f;
    if (tid < 128) {
        const int hh = tid / p.W2, ww = tid - hh * p.W2;
        rowrel[tid] = (ww < p.W && hh < p.BH) ? hh * p.W + ww : (int)0x80000000;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 8) {
        // ===================== epilogue group grp handles tiles j with (j & 1) == grp =====================
        const int grp = warp >> 2, w4 = warp & 3;
        EpiParams ep = p.tc.ep;
        ep.bias = nullptr;
        const uint32_t stg_s = smem_u32(stg_all + (size_t)grp * (STG_BYTES / 4) + w4 * STG_WARP), bias_sa = smem_u32(bias_s), rowrel_sa = smem_u32(rowrel);
        int j = grp;
        for (int t = blockIdx.x + grp * gridDim.x; t < p.total_tiles; t += 2 * gridDim.x, j += 2) {
            const int n = t / p.tph, h0 = (t - n * p.tph) * p.BH;
            TileInfo ti;
            ti.mtile = t; ti.n0 = 0; ti.z = 0; ti.kb_begin = 0; ti.nkb = 0; ti.m0 = 0;
            ti.row_base = ((long long)n * p.H + h0) * p.W;
            ti.rows_valid = min(p.BH, p.H - h0) * p.W2;
            const int buf = grp * 2 + ((j >> 1) & 1);   // this group's k-th tile (k = j >> 1) uses its accumulator k & 1
            mbar_wait(&accum_full[buf], (uint32_t)((j >> 2) & 1));
            tc_fence_after();
            const uint32_t lane_addr = tmem_base + ((uint32_t)(w4 * 32) << 16) + (uint32_t)(buf * 64);
            float* stats_dst = ep.colstats ? ep.colstats + (size_t)(t % AVEC_STATS_REPLICAS) * 2 * 64 : nullptr;
            const bool bias_on = p.tc.ep.bias != nullptr;
            if (ep.kind == AVEC_EPI_RESIDUAL) epilogue_fast<AVEC_EPI_RESIDUAL, false, true>(p.tc, ep, ti, lane_addr, stg_s, bias_sa, stats_dst, w4, lane, bias_on, rowrel_sa);
            else if (ep.colstats) epilogue_fast<AVEC_EPI_LINEAR, true, true>(p.tc, ep, ti, lane_addr, stg_s, bias_sa, stats_dst, w4, lane, bias_on, rowrel_sa);
            else epilogue_fast<AVEC_EPI_LINEAR, false, true>(p.tc, ep, ti, lane_addr, stg_s, bias_sa, stats_dst, w4, lane, bias_on, rowrel_sa);
            tc_fence_before();
            mbar_arrive(&accum_empty[buf]);
        }
        tc_fence_before();
    } else if (warp == 8) {
        // ===================== weight load + MMA issuer =====================
        if (elect_one()) {
            mbar_expect_tx(w_full, CH_W_BYTES);
            for (int tap = 0; tap < 9; ++tap) tma_load_2d(smem_u32(w_s) + (uint32_t)tap * 8192u, &mapW, w_full, tap * 64, 0);
        }
        __syncwarp();
        mbar_wait(w_full, 0);
        tc_fence_after();
        const uint32_t idesc = make_idesc(64, 0, 0);
        const uint64_t desc0 = make_smem_desc(0, 16, 1024);
        const uint32_t d_hi = (uint32_t)(desc0 >> 32), d_lo = (uint32_t)desc0;
        const uint32_t w16 = smem_u32(w_s) >> 4, x16_0 = smem_u32(x_s) >> 4;
        int j = 0, st = 0;
        uint32_t ph = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++j) {
            const int buf = (j & 1) * 2 + ((j >> 1) & 1);
            mbar_wait(&accum_empty[buf], (uint32_t)(((j >> 2) & 1) ^ 1));
            mbar_wait(&x_full[st], ph);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * 64);
                const uint32_t x16 = x16_0 + (uint32_t)st * (CH_X_BYTES >> 4);
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                    const int kh = tap / 3, kw = tap - kh * 3;
                    const int th = p.dgrad ? 2 - kh : kh, tw = p.dgrad ? 2 - kw : kw;
                    const uint32_t a0 = d_lo + x16 + (uint32_t)((th * p.W2 + tw) * 128 >> 4);
                    const uint32_t b0 = d_lo + w16 + (uint32_t)tap * (8192u >> 4);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        umma_f16(d_tmem, ((uint64_t)d_hi << 32) | (a0 + (uint32_t)ks * 2u), ((uint64_t)d_hi << 32) | (b0 + (uint32_t)ks * 2u), idesc,
                                 (tap > 0 || ks > 0) ? 1u : 0u);
                }
                umma_commit(&x_empty[st]);
                umma_commit(&accum_full[buf]);
            }
            __syncwarp();
            if (++st == CH_STAGES) { st = 0; ph ^= 1u; }
        }
        tc_fence_before();
    } else {
        // ===================== halo TMA producer =====================
        int st = 0;
        uint32_t ph = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
            mbar_wait(&x_empty[st], ph ^ 1u);
            if (elect_one()) {
                const int n = t / p.tph, h0 = (t - n * p.tph) * p.BH;
                mbar_expect_tx(&x_full[st], (uint32_t)p.x_tx);
                tma_load_4d(smem_u32(x_s) + (uint32_t)st * CH_X_BYTES, &mapX, &x_full[st], 0, -1, h0 - 1, n);
            }
            __syncwarp();
            if (++st == CH_STAGES) { st = 0; ph ^= 1u; }
        }
    }
    __syncthreads();
    if (warp == 8) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
}

// ---- the same resident-dW scheme for the wider 3x3 stride-1 stages whose images fit one tile (ResNet stage 2: 11 x 11 x 128,
// stage 3: 6 x 6 x 256).  A tile is BI whole images, each laid out on a (H+2) x G grid (G = W+2 rounded up to a multiple of 8 so
// that the second filter-row M group starts on a 1024-byte swizzle atom); dY and the X halo use the same per-image pitch, so
// a filter tap is again a constant row offset kh*G+kw for the whole tile.  dW (Co x 9C fp32) exceeds tensor memory, so each
// persistent CTA owns one (64 input channels) x (64 output channels) block of it (6 accumulators x 64 columns) and streams
// its share of the images; CTA i works on block i % nblocks.
constexpr int WI_THREADS = 192;
constexpr size_t WI_SMEM_MAX = 225 * 1024;

struct WgImgParams {
    int N, H, W, G, BI, C, Co, ksteps, k_rows, x_rows, stage_bytes, stages, ncb, ncob, tiles;
    int y_tx, x_tx;
    float alpha;
    float* dw;    // [Co][9 * C] fp32
};

__global__ void __launch_bounds__(WI_THREADS, 1) wgrad_img_kernel(const __grid_constant__ WgImgParams p, const __grid_constant__ CUtensorMap mapY,
                                                                 const __grid_constant__ CUtensorMap mapX) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* ctrl = smem + (size_t)p.stages * p.stage_bytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(ctrl);   // [<= 8]
    uint64_t* empty = full + 8;
    uint64_t* done = empty + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 2);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    if (tid == 0) {
        for (int i = 0; i < p.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 2); }   // two MMA-issuing threads release a stage
        mbar_init(done, 2);
        fence_barrier_init();
        tma_prefetch_desc(&mapY);
        tma_prefetch_desc(&mapX);
    }
    if (warp == 4) tmem_alloc(tmem_slot, 512);
    for (int i = tid; i < p.stages * p.stage_bytes / 16; i += WI_THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int nblocks = p.ncb * p.ncob;
    const int blk = blockIdx.x % nblocks, part = blockIdx.x / nblocks, nparts = gridDim.x / nblocks;   // gridDim.x is a multiple of nblocks
    const int c0 = (blk % p.ncb) * 64, co0 = (blk / p.ncb) * 64;
    const bool any_tile = part < p.tiles;
    const int y_bytes = p.k_rows * 128;
    // MMA issue loop for accumulators [q_lo, q_hi).  One thread needs ~36 cycles of scalar work per 32-cycle 128x64x16 MMA, so
    // the six accumulators are split over TWO issuing threads (warp 4 and warp 0, which only has the final flush to do)
    auto issue_loop = [&](int q_lo, int q_hi) {
        const uint32_t idesc = make_idesc(64, 1, 1);
        const uint64_t a_desc0 = make_smem_desc(0, (uint32_t)p.G * 128u, 1024);
        const uint64_t b_desc0 = make_smem_desc(0, 16384, 1024);
        const uint32_t a_hi = (uint32_t)(a_desc0 >> 32), b_hi = (uint32_t)(b_desc0 >> 32);
        const uint32_t base16 = smem_u32(smem) >> 4;
        int j = 0, st = 0;
        uint32_t ph = 0;
        for (int t = part; t < p.tiles; t += nparts, ++j) {
            mbar_wait(&full[st], ph);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t y16 = base16 + (uint32_t)st * ((uint32_t)p.stage_bytes >> 4);
                const uint32_t x16 = y16 + ((uint32_t)y_bytes >> 4);
                auto issue = [&](auto KS) {
                    constexpr int NKS = decltype(KS)::value;
                    const int nks = NKS > 0 ? NKS : p.ksteps;
#pragma unroll
                    for (int qq = 0; qq < 3; ++qq) {
                        const int q = q_lo + qq;
                        const int set = q / 3, kw = q - set * 3;
                        const uint32_t a0 = (uint32_t)a_desc0 + x16 + (uint32_t)((set * p.G + kw) * 128 >> 4);
                        const uint32_t b0 = (uint32_t)b_desc0 + y16;
                        const uint32_t d_tmem = tmem_base + (uint32_t)(q * 64);
                        if (NKS > 0) {
#pragma unroll
                            for (int ks = 0; ks < (NKS > 0 ? NKS : 1); ++ks)
                                umma_f16(d_tmem, ((uint64_t)a_hi << 32) | (a0 + (uint32_t)ks * (2048u >> 4)), ((uint64_t)b_hi << 32) | (b0 + (uint32_t)ks * (2048u >> 4)),
                                         idesc, (j > 0 || ks > 0) ? 1u : 0u);
                        } else {
                            for (int ks = 0; ks < nks; ++ks)
                                umma_f16(d_tmem, ((uint64_t)a_hi << 32) | (a0 + (uint32_t)ks * (2048u >> 4)), ((uint64_t)b_hi << 32) | (b0 + (uint32_t)ks * (2048u >> 4)),
                                         idesc, (j > 0 || ks > 0) ? 1u : 0u);
                        }
                    }
                };
                if (p.ksteps == 8) issue(std::integral_constant<int, 8>{});
                else if (p.ksteps == 13) issue(std::integral_constant<int, 13>{});
                else issue(std::integral_constant<int, 0>{});
                umma_commit(&empty[st]);
            }
            __syncwarp();
            if (++st == p.stages) { st = 0; ph ^= 1u; }
        }
        (void)q_hi;
        if (any_tile && elect_one()) umma_commit(done);
        __syncwarp();
        tc_fence_before();
    };


    if (warp < 4) {
        if (warp == 0) { issue_loop(3, 6); tc_fence_after(); }
        if (any_tile) {
            mbar_wait(done, 0);
            tc_fence_after();
            const int row = warp * 32 + lane, g = row >> 6, c = row & 63;
            const size_t ldw = (size_t)9 * p.C;
            for (int q = 0; q < 6; ++q) {
                const int set = q / 3, kw = q - set * 3, kh = set + g;
                if (set == 1 && g == 0) continue;   // kh = 1 duplicate (g is warp-uniform)
                float* dst = p.dw + (size_t)co0 * ldw + (size_t)(kh * 3 + kw) * p.C + c0 + c;
                for (int cc = 0; cc < 64; cc += 16) {
                    float v[16];
                    tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(q * 64 + cc), v);
#pragma unroll
                    for (int i = 0; i < 16; ++i) atomicAdd(dst + (size_t)(cc + i) * ldw, p.alpha * v[i]);
                }
            }
        }
        tc_fence_before();
    } else if (warp == 4) {
        issue_loop(0, 3);
        } else {
        int st = 0;
        uint32_t ph = 0;
        for (int t = part; t < p.tiles; t += nparts) {
            mbar_wait(&empty[st], ph ^ 1u);
            if (elect_one()) {
                const int n0 = t * p.BI;
                const uint32_t ydst = smem_u32(smem) + (uint32_t)st * (uint32_t)p.stage_bytes;
                mbar_expect_tx(&full[st], (uint32_t)(p.y_tx + p.x_tx));
                tma_load_4d(ydst, &mapY, &full[st], co0, 0, 0, n0);
                tma_load_4d(ydst + (uint32_t)y_bytes, &mapX, &full[st], c0, -1, -1, n0);
            }
            __syncwarp();
            if (++st == p.stages) { st = 0; ph ^= 1u; }
        }
    }
    __syncthreads();
    if (warp == 4) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// tile width: as wide as possible (<= 256, multiple of 16, awkward N such as 180 / 720 / 1080 split evenly), but narrow
// enough that small problems still put >= ~100 CTAs on the 148 SMs (never below 64 columns)
int pick_bn(int N, int mtiles) {
    int ntiles = cdiv(N, 256);
    while (mtiles * ntiles < 100 && cdiv(cdiv(N, ntiles + 1), 16) * 16 >= 64) ++ntiles;
    int bn = cdiv(cdiv(N, ntiles), 16) * 16;
    return bn < 16 ? 16 : bn;
}
int ptr_align(const void* p, long long ld_elems) {
    uintptr_t a = reinterpret_cast<uintptr_t>(p);
    long long ldb = ld_elems * 2;
    int al = 16;
    while (al > 2 && ((a % al) != 0 || (ldb % al) != 0)) al >>= 1;
    return al;
}

// ---- tensor maps (driver entry point resolved through the runtime: no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(f);
    }
    return fn;
}
// rank-d bf16 tensor map, 128B swizzle, zero OOB fill.  dims/box innermost first; strides (bytes) for dims 1..d-1.
bool encode_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes, const cuuint32_t* box,
                const cuuint32_t* elem_strides = nullptr) {
    EncodeTiledFn fn = get_encode();
    if (!fn) return false;
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    if (elem_strides) for (int i = 0; i < rank; ++i) es[i] = elem_strides[i];
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}
bool tma_ok_2d(const void* base, long long ld) { return (reinterpret_cast<uintptr_t>(base) % 16) == 0 && (ld * 2) % 16 == 0 && get_encode() != nullptr; }

// row-aligned conv tiles on the OUTPUT grid: whole images (BI per tile) when an image has <= 128 sites, else BH rows of one image
bool conv_tiling(const ConvGeom& g, int& BH, int& BI, int& tph) {
    const int hw = g.Ho * g.Wo;
    if (g.Wo > 128) return false;
    if (hw <= 128) { BI = 128 / hw; BH = g.Ho; tph = 1; }
    else { BI = 1; BH = 128 / g.Wo; tph = cdiv(g.Ho, BH); }
    return BH >= 1 && (BH - 1) * g.sh + 1 <= 256 && BI <= 256 && (g.Wo - 1) * g.sw + 1 <= 256;
}
// 2-d "same" convolutions with odd square filters and stride 1 or 2 (strided ones via tensor-map element strides)
bool conv_tma_geom_ok(const ConvGeom& g, bool dgrad) {
    const bool base = g.KT == 1 && g.Ti == 1 && g.st == 1 && g.sh == g.sw && (g.sh == 1 || g.sh == 2) && g.KH == g.KW && (g.KH % 2) == 1 &&
                      g.ph == (g.KH - 1) / 2 && g.pw == (g.KW - 1) / 2 && g.Ho == (g.Hi - 1) / g.sh + 1 && g.Wo == (g.Wi - 1) / g.sw + 1;
    return base && (!dgrad || g.sh == 1);
}

// per-DEVICE caches: cudaFuncSetAttribute and the SM count belong to the current device, and one process may drive several
constexpr int MAX_DEVICES = 64;
int current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return dev >= 0 && dev < MAX_DEVICES ? dev : 0;
}
int num_sms_cached() {
    static int num_sms[MAX_DEVICES] = {0};
    const int dev = current_device();
    if (num_sms[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        num_sms[dev] = n;
    }
    return num_sms[dev];
}
// opt-in dynamic shared memory of kernel `slot`, set once per device
bool smem_attr_once(int slot, const void* fn, int bytes) {
    static bool done[8][MAX_DEVICES] = {{false}};
    const int dev = current_device();
    if (done[slot][dev]) return true;
    if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) return false;
    done[slot][dev] = true;
    return true;
}

}  // namespace

static int g_tma_enabled = 1;
static bool g_tma_enabled_flag() { return g_tma_enabled != 0; }
static unsigned long long* g_dbg_ts = nullptr;
extern "C" void avec_set_tma(int enabled) { g_tma_enabled = enabled; }
extern "C" void avec_set_debug_timestamps(void* dev_buf_8_u64) { g_dbg_ts = reinterpret_cast<unsigned long long*>(dev_buf_8_u64); }

bool avec_gemm_tc_supported(const avec_gemm_args* a) {
    if (a->ab_dtype != AVEC_BF16) return false;
    const avec_conv_geom& g = a->g;
    switch (a->mode) {
    case AVEC_GEMM_PLAIN: {
        bool ak = a->sak == 1, am = a->sam == 1;
        bool bk = a->sbk == 1, bm = a->sbn == 1;
        if (!(ak || am) || !(bk || bm)) return false;
        // tiny problems are not worth a 128-row tile
        if ((long long)a->M * a->N * a->K < (1LL << 18)) return false;
        return true;
    }
    case AVEC_GEMM_CONV_FWD: return g.C % 64 == 0 || (g.C == 1 && g.KT * g.KH * g.KW <= 256 && g.KT < 256 && g.KH < 256);
    case AVEC_GEMM_CONV_DGRAD: return g.Co % 64 == 0;
    case AVEC_GEMM_CONV_WGRAD: return g.C % 64 == 0 || (g.C == 1 && g.KT * g.KH * g.KW <= 256 && g.KT < 256 && g.KH < 256);
    default: return false;
    }
}

namespace {
struct DgradClass { int a, b; };   // output parity class (rows 2i + a, columns 2j + b) of a stride-2 3x3 dgrad
}
static int gemm_tc_launch(const avec_gemm_args* a, cudaStream_t st, const DgradClass* dc);

// Stride-2 3x3 dgrad, exact: dx[2i+a, 2j+b] only receives the filter taps of matching parity (1, 2, 2 or 4 of the 9), so each
// of the 4 parity classes is a stride-1 convolution over the dY grid with a sub-filter, written to a strided sub-grid of dx
// (together the classes cover dx exactly once).  9 tap-GEMMs instead of the 36 of the zero-insertion formulation.
static bool dgrad_classes_ok(const avec_gemm_args* a) {
    const avec_conv_geom& g = a->g;
    static int on = -1;
    if (on < 0) { const char* e = getenv("AVEC_DGRAD_CLASSES"); on = e ? atoi(e) : 1; }
    if (!on || a->mode != AVEC_GEMM_CONV_DGRAD || !g_tma_enabled_flag() || get_encode() == nullptr) return false;
    if (!(g.KT == 1 && g.Ti == 1 && g.st == 1 && g.sh == 2 && g.sw == 2 && g.KH == 3 && g.KW == 3 && g.ph == 1 && g.pw == 1)) return false;
    if (g.C % 64 != 0 || g.Co % 64 != 0 || g.Ho != (g.Hi - 1) / 2 + 1 || g.Wo != (g.Wi - 1) / 2 + 1 || (g.Wi + 1) / 2 > 128) return false;
    if (a->ab_dtype != AVEC_BF16 || a->out_dtype != AVEC_BF16 || (a->epi != AVEC_EPI_LINEAR && a->epi != AVEC_EPI_RESIDUAL)) return false;
    if (a->bias || a->colstats || a->out2 || (a->epi == AVEC_EPI_RESIDUAL && (!a->aux || a->aux_dtype != AVEC_BF16))) return false;
    if (ptr_align(a->out, a->ldo) < 8 || (a->aux && ptr_align(a->aux, a->ldaux) < 8)) return false;
    return (reinterpret_cast<uintptr_t>(a->A) % 16) == 0 && (reinterpret_cast<uintptr_t>(a->B) % 16) == 0;
}

// ---- 64-channel 3x3 stride-1 wgrad on halo tiles (see wgrad_halo64_kernel)
static bool wgrad_halo_ok(const avec_gemm_args* a) {
    const avec_conv_geom& g = a->g;
    static int on = -1;
    if (on < 0) { const char* e = getenv("AVEC_WGRAD_HALO"); on = e ? atoi(e) : 1; }
    if (!on || a->mode != AVEC_GEMM_CONV_WGRAD || !g_tma_enabled || get_encode() == nullptr || a->ab_dtype != AVEC_BF16) return false;
    if (!(g.KT == 1 && g.Ti == 1 && g.st == 1 && g.sh == 1 && g.sw == 1 && g.KH == 3 && g.KW == 3 && g.ph == 1 && g.pw == 1)) return false;
    if (g.C != 64 || g.Co != 64 || g.Ho != g.Hi || g.Wo != g.Wi) return false;
    const int W2 = g.Wi + 2;
    if (W2 % 8 != 0 || W2 > 64 || g.Hi < 1) return false;
    const int BH = 128 / W2;
    if (BH < 1 || (BH + 2) * W2 > 168) return false;
    if (a->epi != AVEC_EPI_ACCUM || a->out_dtype != AVEC_F32 || a->ldo != 576 || a->bias || a->colstats) return false;
    return (reinterpret_cast<uintptr_t>(a->A) % 16) == 0 && (reinterpret_cast<uintptr_t>(a->B) % 16) == 0;
}

static int wgrad_halo_launch(const avec_gemm_args* a, cudaStream_t st) {
    const avec_conv_geom& g = a->g;
    WgHaloParams p;
    memset(&p, 0, sizeof(p));
    p.N = g.N; p.H = g.Hi; p.W = g.Wi; p.W2 = g.Wi + 2; p.BH = 128 / p.W2; p.tph = cdiv(p.H, p.BH);
    const long long tiles = (long long)p.N * p.tph;
    if (tiles > 0x7fffffffLL) return AVEC_ERR_INVALID;
    p.total_tiles = (int)tiles;
    p.y_tx = p.BH * p.W2 * 128; p.x_tx = (p.BH + 2) * p.W2 * 128;
    p.alpha = a->alpha; p.dw = reinterpret_cast<float*>(a->out);
    CUtensorMap mapY, mapX;
    cuuint64_t d[4] = {64, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.N};
    cuuint64_t sb[3] = {128, (cuuint64_t)p.W * 128, (cuuint64_t)p.W * p.H * 128};
    cuuint32_t boxY[4] = {64, (cuuint32_t)p.W2, (cuuint32_t)p.BH, 1}, boxX[4] = {64, (cuuint32_t)p.W2, (cuuint32_t)(p.BH + 2), 1};
    if (!encode_map(&mapY, a->A, 4, d, sb, boxY) || !encode_map(&mapX, a->B, 4, d, sb, boxX)) return AVEC_ERR_DRIVER;
    if (!smem_attr_once(0, reinterpret_cast<const void*>(wgrad_halo64_kernel), (int)WH_SMEM)) return AVEC_ERR_LAUNCH;
    const int ctas = std::min(p.total_tiles, num_sms_cached());
    wgrad_halo64_kernel<<<ctas, WH_THREADS, WH_SMEM, st>>>(p, mapY, mapX);
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}

// ---- 64-channel 3x3 stride-1 forward / dgrad on halo tiles (see conv3x3_halo64_kernel)
static bool conv_halo_ok(const avec_gemm_args* a) {
    const avec_conv_geom& g = a->g;
    static int on = -1;
    if (on < 0) { const char* e = getenv("AVEC_CONV_HALO64"); on = e ? atoi(e) : 1; }
    if (!on || (a->mode != AVEC_GEMM_CONV_FWD && a->mode != AVEC_GEMM_CONV_DGRAD) || !g_tma_enabled || get_encode() == nullptr) return false;
    if (!(g.KT == 1 && g.Ti == 1 && g.st == 1 && g.sh == 1 && g.sw == 1 && g.KH == 3 && g.KW == 3 && g.ph == 1 && g.pw == 1)) return false;
    if (g.C != 64 || g.Co != 64 || g.Ho != g.Hi || g.Wo != g.Wi || a->ab_dtype != AVEC_BF16 || a->out_dtype != AVEC_BF16) return false;
    const int W2 = g.Wi + 2;
    if (W2 > 64 || g.Hi * g.Wi <= 128) return false;   // small images: whole-image tiles of the generic kernel are better
    const int BH = 128 / W2;
    if (BH < 1 || (BH + 2) * W2 > 168) return false;
    if (a->epi != AVEC_EPI_LINEAR && a->epi != AVEC_EPI_RESIDUAL) return false;
    if (a->out2 || a->alpha != 1.0f || (a->epi == AVEC_EPI_RESIDUAL && (!a->aux || a->aux_dtype != AVEC_BF16 || a->colstats))) return false;
    if (ptr_align(a->out, a->ldo) < 16 || (a->aux && ptr_align(a->aux, a->ldaux) < 16) || (a->colstats && (reinterpret_cast<uintptr_t>(a->colstats) % 16) != 0)) return false;
    return (reinterpret_cast<uintptr_t>(a->A) % 16) == 0 && (reinterpret_cast<uintptr_t>(a->B) % 16) == 0 && a->K == 576;
}

static int conv_halo_launch(const avec_gemm_args* a, cudaStream_t st) {
    const avec_conv_geom& g = a->g;
    ConvHaloParams p;
    memset(&p, 0, sizeof(p));
    p.N = g.N; p.H = g.Hi; p.W = g.Wi; p.W2 = g.Wi + 2; p.BH = 128 / p.W2; p.tph = cdiv(p.H, p.BH);
    const long long tiles = (long long)p.N * p.tph;
    if (tiles > 0x7fffffffLL) return AVEC_ERR_INVALID;
    p.total_tiles = (int)tiles;
    p.x_tx = (p.BH + 2) * p.W2 * 128;
    p.dgrad = a->mode == AVEC_GEMM_CONV_DGRAD ? 1 : 0;
    p.tc.BN = 64; p.tc.N = 64; p.tc.M = a->M; p.tc.epi_fast = 2; p.tc.rm_on = 1;
    p.tc.ep = make_epi(a);
    CUtensorMap mapX, mapW;
    cuuint64_t d[4] = {64, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.N};
    cuuint64_t sb[3] = {128, (cuuint64_t)p.W * 128, (cuuint64_t)p.W * p.H * 128};
    cuuint32_t boxX[4] = {64, (cuuint32_t)p.W2, (cuuint32_t)(p.BH + 2), 1};
    cuuint64_t dw_[2] = {576, 64}; cuuint64_t sw_[1] = {576 * 2}; cuuint32_t boxW[2] = {64, 64};
    if (!encode_map(&mapX, a->A, 4, d, sb, boxX) || !encode_map(&mapW, a->B, 2, dw_, sw_, boxW)) return AVEC_ERR_DRIVER;
    if (!smem_attr_once(1, reinterpret_cast<const void*>(conv3x3_halo64_kernel), (int)CH_SMEM)) return AVEC_ERR_LAUNCH;
    const int ctas = std::min(p.total_tiles, num_sms_cached());
    conv3x3_halo64_kernel<<<ctas, CH_THREADS, CH_SMEM, st>>>(p, mapX, mapW);
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}

// ---- whole-image resident-dW wgrad (see wgrad_img_kernel)
static bool wgrad_img_plan(const avec_gemm_args* a, WgImgParams& p) {
    const avec_conv_geom& g = a->g;
    static int on = -1;
    if (on < 0) { const char* e = getenv("AVEC_WGRAD_IMG"); on = e ? atoi(e) : 1; }
    if (!on || a->mode != AVEC_GEMM_CONV_WGRAD || !g_tma_enabled || get_encode() == nullptr || a->ab_dtype != AVEC_BF16) return false;
    if (!(g.KT == 1 && g.Ti == 1 && g.st == 1 && g.sh == 1 && g.sw == 1 && g.KH == 3 && g.KW == 3 && g.ph == 1 && g.pw == 1)) return false;
    if (g.C % 64 != 0 || g.Co % 64 != 0 || g.Ho != g.Hi || g.Wo != g.Wi) return false;
    if (a->epi != AVEC_EPI_ACCUM || a->out_dtype != AVEC_F32 || a->ldo != 9LL * g.C || a->bias || a->colstats) return false;
    if ((reinterpret_cast<uintptr_t>(a->A) % 16) != 0 || (reinterpret_cast<uintptr_t>(a->B) % 16) != 0) return false;
    memset(&p, 0, sizeof(p));
    p.N = g.N; p.H = g.Hi; p.W = g.Wi; p.C = g.C; p.Co = g.Co;
    p.G = (g.Wi + 2 + 7) / 8 * 8;
    const int P = (g.Hi + 2) * p.G;                     // rows per image (dY grid and X halo share the pitch)
    if (p.G > 32 || P > 256 || g.Hi * g.Wi * 5 < P * 2) return false;   // at least 40 % of the grid must be real sites
    p.BI = 1;
    while ((p.BI + 1) * P <= 128) ++p.BI;
    while ((p.BI * P) % 16 != 0) ++p.BI;                 // P is a multiple of 8
    if (p.BI * P > 256 || p.BI > 256) return false;
    p.k_rows = p.BI * P; p.ksteps = p.k_rows / 16;
    p.x_rows = p.k_rows + 2 * p.G + 8;
    p.stage_bytes = ((p.k_rows + p.x_rows) * 128 + 1023) / 1024 * 1024;
    p.stages = (int)std::min<size_t>(6, (WI_SMEM_MAX - 2048) / p.stage_bytes);
    if (p.stages < 2) return false;
    p.ncb = g.C / 64; p.ncob = g.Co / 64;
    if (p.ncb * p.ncob > num_sms_cached()) return false;
    p.tiles = cdiv(g.N, p.BI);
    p.y_tx = p.k_rows * 128; p.x_tx = p.k_rows * 128;
    p.alpha = a->alpha; p.dw = reinterpret_cast<float*>(a->out);
    return true;
}

static int wgrad_img_launch(const avec_gemm_args* a, const WgImgParams& p, cudaStream_t st) {
    CUtensorMap mapY, mapX;
    cuuint64_t dY[4] = {(cuuint64_t)p.Co, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.N};
    cuuint64_t sY[3] = {(cuuint64_t)p.Co * 2, (cuuint64_t)p.W * p.Co * 2, (cuuint64_t)p.W * p.H * p.Co * 2};
    cuuint64_t dX[4] = {(cuuint64_t)p.C, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.N};
    cuuint64_t sX[3] = {(cuuint64_t)p.C * 2, (cuuint64_t)p.W * p.C * 2, (cuuint64_t)p.W * p.H * p.C * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)p.G, (cuuint32_t)(p.H + 2), (cuuint32_t)p.BI};
    if (!encode_map(&mapY, a->A, 4, dY, sY, box) || !encode_map(&mapX, a->B, 4, dX, sX, box)) return AVEC_ERR_DRIVER;
    if (!smem_attr_once(2, reinterpret_cast<const void*>(wgrad_img_kernel), (int)WI_SMEM_MAX)) return AVEC_ERR_LAUNCH;
    const int nblocks = p.ncb * p.ncob;
    const int parts = std::max(1, std::min(num_sms_cached() / nblocks, p.tiles));
    const size_t smem = 1024 + (size_t)p.stages * p.stage_bytes + 1024;
    wgrad_img_kernel<<<nblocks * parts, WI_THREADS, smem, st>>>(p, mapY, mapX);
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}

int avec_gemm_tc(const avec_gemm_args* a, cudaStream_t st) {
    if (wgrad_halo_ok(a)) return wgrad_halo_launch(a, st);
    {
        WgImgParams wp;
        if (wgrad_img_plan(a, wp)) return wgrad_img_launch(a, wp, st);
    }
    if (conv_halo_ok(a)) return conv_halo_launch(a, st);
    if (dgrad_classes_ok(a)) {
        for (int ca = 0; ca < 2; ++ca)
            for (int cb = 0; cb < 2; ++cb) {
                if ((a->g.Hi - ca + 1) / 2 <= 0 || (a->g.Wi - cb + 1) / 2 <= 0) continue;
                DgradClass dc{ca, cb};
                const int rc = gemm_tc_launch(a, st, &dc);
                if (rc != AVEC_OK) return rc;
            }
        return AVEC_OK;
    }
    return gemm_tc_launch(a, st, nullptr);
}

static int gemm_tc_launch(const avec_gemm_args* a, cudaStream_t st, const DgradClass* dc) {
    TcParams p;
    memset(&p, 0, sizeof(p));
    CUtensorMap mapA, mapB;
    memset(&mapA, 0, sizeof(mapA));
    memset(&mapB, 0, sizeof(mapB));
    p.M = a->M; p.N = a->N; p.K = a->K;
    p.A = reinterpret_cast<const bf16*>(a->A);
    p.B = reinterpret_cast<const bf16*>(a->B);
    p.g = make_geom(a->g);
    p.ep = make_epi(a);
    p.out_transposed = 0;
    p.ksteps = 4;
    p.dbg_ts = g_dbg_ts;
    p.a_group_stride = p.b_group_stride = 8192;
    const int taps = p.g.KT * p.g.KH * p.g.KW;
    int grid_m = cdiv(a->M, BM);
    bool wgrad_bn_fixed = false;
    int BH = 0, BI = 0, tph = 0;
    const bool conv_tma = !dc && g_tma_enabled && get_encode() != nullptr && (a->mode == AVEC_GEMM_CONV_FWD || a->mode == AVEC_GEMM_CONV_DGRAD ||
                          a->mode == AVEC_GEMM_CONV_WGRAD) && conv_tma_geom_ok(p.g, a->mode == AVEC_GEMM_CONV_DGRAD) && p.g.C % 64 == 0 && p.g.Co % 64 == 0 &&
                          conv_tiling(p.g, BH, BI, tph) && (reinterpret_cast<uintptr_t>(a->A) % 16) == 0 && (reinterpret_cast<uintptr_t>(a->B) % 16) == 0;
    switch (a->mode) {
    case AVEC_GEMM_PLAIN:
        if (a->sak == 1) { p.a_kind = OP_PLAIN_K; p.a_ld = a->sam; } else { p.a_kind = OP_PLAIN_MN; p.a_ld = a->sak; }
        if (a->sbk == 1) { p.b_kind = OP_PLAIN_K; p.b_ld = a->sbn; } else { p.b_kind = OP_PLAIN_MN; p.b_ld = a->sbk; }
        p.num_kb = cdiv(a->K, BKE);
        break;
    case AVEC_GEMM_CONV_FWD:
        AVEC_CHECK_ARG(a->K == taps * p.g.C && a->N == p.g.Co);
        p.b_kind = OP_PLAIN_K; p.b_ld = a->K;
        if (p.g.C == 1) { p.a_kind = OP_CONV_TAPS; p.a_ld = 1; p.cpb = 1; p.num_kb = cdiv(taps, BKE); }
        else { p.a_kind = conv_tma ? OP_TMA_CONV_K : OP_CONV_FWD; p.a_ld = p.g.C; p.cpb = p.g.C / 64; p.num_kb = taps * p.cpb; }
        break;
    case AVEC_GEMM_CONV_DGRAD:
        AVEC_CHECK_ARG(a->K == taps * p.g.Co && a->N == p.g.C);
        p.a_kind = conv_tma ? OP_TMA_CONV_K : OP_CONV_DGRAD; p.a_ld = p.g.Co;
        p.b_kind = OP_PLAIN_K; p.b_ld = a->K;
        p.cpb = p.g.Co / 64; p.num_kb = taps * p.cpb;
        p.ct_dgrad = 1;
        break;
    case AVEC_GEMM_CONV_WGRAD:
        AVEC_CHECK_ARG(a->M == p.g.Co && a->N == taps * p.g.C);
        p.a_kind = conv_tma ? OP_TMA_CONV_MN : OP_PLAIN_MN; p.a_ld = p.g.Co;
        p.b_kind = p.g.C == 1 ? OP_CONV_TAPS_MN : (conv_tma ? OP_TMA_CONV_MN : OP_CONV_WGRAD_X); p.b_ld = p.g.C;
        p.cpb = p.g.C == 1 ? 1 : p.g.C / 64; p.num_kb = cdiv(a->K, BKE);
        wgrad_bn_fixed = true;
        break;
    default: return AVEC_ERR_INVALID;
    }
    int class_tile_rows = 0;
    if (dc) {
        const ConvGeom& g = p.g;
        const int Hs = (g.Hi - dc->a + 1) / 2, Ws = (g.Wi - dc->b + 1) / 2;
        ConvGeom gs = g;
        gs.Ho = Hs; gs.Wo = Ws; gs.sh = gs.sw = 1;
        if (!conv_tiling(gs, BH, BI, tph)) return AVEC_ERR_UNSUPPORTED;
        // taps of matching parity: (kh, dh) in {(1, 0)} for even rows, {(0, +1), (2, 0)} for odd rows; same for columns
        const int kh_a[2][2] = {{1, -1}, {0, 2}}, dh_a[2][2] = {{0, 0}, {1, 0}}, n_a[2] = {1, 2};
        int nt = 0;
        for (int i = 0; i < n_a[dc->a]; ++i)
            for (int j = 0; j < n_a[dc->b]; ++j) {
                p.tap_dh[nt] = (signed char)dh_a[dc->a][i]; p.tap_dw[nt] = (signed char)dh_a[dc->b][j];
                p.tap_koff[nt] = (kh_a[dc->a][i] * 3 + kh_a[dc->b][j]) * g.Co;
                ++nt;
            }
        p.gen_taps = nt;
        p.a_kind = OP_TMA_CONV_K; p.cpb = g.Co / 64; p.num_kb = nt * p.cpb; p.ct_dgrad = 0;
        p.conv_tiles = 1; p.ct_BH = BH; p.ct_BI = BI; p.ct_tph = tph; p.ct_H = Hs; p.ct_W = Ws; p.ct_s = 1;
        grid_m = BI == 1 ? g.N * tph : cdiv(g.N, BI);
        class_tile_rows = BI == 1 ? BH * Ws : BI * Hs * Ws;
        p.a_tx = class_tile_rows * 128;
        p.rm_on = 1; p.rm_Hi = g.Hi; p.rm_Wi = g.Wi; p.rm_sy = 2; p.rm_sx = 2; p.rm_a = dc->a; p.rm_b = dc->b;
        cuuint64_t dA[4] = {(cuuint64_t)g.Co, (cuuint64_t)g.Wo, (cuuint64_t)g.Ho, (cuuint64_t)g.N};
        cuuint64_t sA[3] = {(cuuint64_t)g.Co * 2, (cuuint64_t)g.Co * g.Wo * 2, (cuuint64_t)g.Co * g.Wo * g.Ho * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)Ws, (cuuint32_t)BH, (cuuint32_t)BI};
        if (!encode_map(&mapA, a->A, 4, dA, sA, box)) return AVEC_ERR_DRIVER;
    }
    p.a_align = ptr_align(p.A, p.a_ld);
    p.b_align = ptr_align(p.B, p.b_ld);
    p.BN = pick_bn(a->N, (a->mode == AVEC_GEMM_PLAIN && a->epi != AVEC_EPI_ACCUM) ? cdiv(a->M, BM) : 1000);
    {
        static int max_bn = -1;
        if (max_bn < 0) { const char* e = getenv("AVEC_MAX_BN"); max_bn = e ? atoi(e) : 256; }
        if (p.BN > max_bn) p.BN = cdiv(cdiv(a->N, cdiv(a->N, max_bn)), 16) * 16;
    }
    if (wgrad_bn_fixed) p.BN = a->N >= 256 ? 256 : (a->N >= 192 ? 192 : (a->N >= 128 ? 128 : 64));

    // ---- halo mode: stride-1 3x3 fwd / dgrad on images larger than one tile (ResNet stage 1)
    bool halo = false;
    {
        static int halo_on = -1;
        if (halo_on < 0) { const char* e = getenv("AVEC_HALO"); halo_on = e ? atoi(e) : 0; }
        const ConvGeom& g = p.g;
        if (conv_tma && halo_on && (a->mode == AVEC_GEMM_CONV_FWD || a->mode == AVEC_GEMM_CONV_DGRAD) && g.sh == 1 && g.KH == 3 && g.KW == 3 &&
            g.Ho * g.Wo > 128 && g.Wo + 2 <= 128) {
            const int W2 = g.Wo + 2, bh = 128 / W2;
            const int cin_blocks = (a->mode == AVEC_GEMM_CONV_FWD ? g.C : g.Co) / 64;
            const int hb = cdiv((bh + 2) * W2 * 128, 1024) * 1024;
            if (bh >= 1 && bh * g.Wo * 10 >= BH * g.Wo * 9 && 2 * cin_blocks * hb <= 96 * 1024) {
                halo = true;
                BH = bh; BI = 1; tph = cdiv(g.Ho, bh);
                p.halo_bytes = hb; p.halo_W2 = W2;
            }
        }
    }
    // ---- conv TMA geometry
    if (conv_tma) {
        p.conv_tiles = 1; p.ct_BH = BH; p.ct_BI = BI; p.ct_tph = tph;
        p.ct_H = p.g.Ho; p.ct_W = p.g.Wo; p.ct_s = p.g.sh;
        const int ntiles = BI == 1 ? p.g.N * tph : cdiv(p.g.N, BI);
        const int tile_rows = BI == 1 ? BH * p.g.Wo : BI * p.g.Ho * p.g.Wo;   // <= 128
        const cuuint32_t sbox[4] = {64, (cuuint32_t)(halo ? p.halo_W2 : (p.g.Wo - 1) * p.g.sw + 1),
                                    (cuuint32_t)(halo ? BH + 2 : (BH - 1) * p.g.sh + 1), (cuuint32_t)BI};
        const cuuint32_t sstr[4] = {1, (cuuint32_t)p.g.sw, (cuuint32_t)p.g.sh, 1};
        // 4-d maps over the NHWC tensors
        const ConvGeom& g = p.g;
        if (a->mode == AVEC_GEMM_CONV_WGRAD) {
            if (p.BN > 128) p.BN = 128;   // 2 groups of B per stage keeps 3 stages in shared memory
            p.ksteps = cdiv(tile_rows, 16);
            const int krows = p.ksteps * 16;
            p.a_rows = 2 * krows; p.b_rows = (p.BN / 64) * krows;
            p.a_group_stride = p.b_group_stride = krows * 128;
            p.num_kb = ntiles;
            p.a_tx = 2 * tile_rows * 128; p.b_tx = (p.BN / 64) * tile_rows * 128;
            cuuint64_t dA[4] = {(cuuint64_t)g.Co, (cuuint64_t)g.Wo, (cuuint64_t)g.Ho, (cuuint64_t)g.N};
            cuuint64_t sA[3] = {(cuuint64_t)g.Co * 2, (cuuint64_t)g.Co * g.Wo * 2, (cuuint64_t)g.Co * g.Wo * g.Ho * 2};
            cuuint64_t dB[4] = {(cuuint64_t)g.C, (cuuint64_t)g.Wi, (cuuint64_t)g.Hi, (cuuint64_t)g.N};
            cuuint64_t sB[3] = {(cuuint64_t)g.C * 2, (cuuint64_t)g.C * g.Wi * 2, (cuuint64_t)g.C * g.Wi * g.Hi * 2};
            cuuint32_t box[4] = {64, (cuuint32_t)g.Wo, (cuuint32_t)BH, (cuuint32_t)BI};
            if (!encode_map(&mapA, a->A, 4, dA, sA, box) || !encode_map(&mapB, a->B, 4, dB, sB, sbox, sstr)) return AVEC_ERR_DRIVER;
        } else {
            const int Cin = a->mode == AVEC_GEMM_CONV_FWD ? g.C : g.Co;   // channels of the gathered tensor
            grid_m = ntiles;
            p.a_tx = halo ? (BH + 2) * p.halo_W2 * 128 : tile_rows * 128;
            if (halo) p.a_kind = OP_TMA_CONV_HALO;
            cuuint64_t dA[4] = {(cuuint64_t)Cin, (cuuint64_t)g.Wi, (cuuint64_t)g.Hi, (cuuint64_t)g.N};
            cuuint64_t sA[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)Cin * g.Wi * 2, (cuuint64_t)Cin * g.Wi * g.Hi * 2};
            if (!encode_map(&mapA, a->A, 4, dA, sA, sbox, sstr)) return AVEC_ERR_DRIVER;
        }
    }
    // ---- plain 2-d TMA operands where strides allow it
    if (g_tma_enabled && p.a_kind == OP_PLAIN_K && tma_ok_2d(p.A, p.a_ld)) {
        cuuint64_t d[2] = {(cuuint64_t)a->K, (cuuint64_t)a->M}; cuuint64_t s1[1] = {(cuuint64_t)p.a_ld * 2}; cuuint32_t box[2] = {64, 128};
        if (encode_map(&mapA, p.A, 2, d, s1, box)) { p.a_kind = OP_TMA_K; p.a_tx = 128 * 128; }
    } else if (g_tma_enabled && p.a_kind == OP_PLAIN_MN && tma_ok_2d(p.A, p.a_ld)) {
        cuuint64_t d[2] = {(cuuint64_t)a->M, (cuuint64_t)a->K}; cuuint64_t s1[1] = {(cuuint64_t)p.a_ld * 2}; cuuint32_t box[2] = {64, 64};
        if (encode_map(&mapA, p.A, 2, d, s1, box)) { p.a_kind = OP_TMA_MN; p.a_tx = 2 * 64 * 128; }
    }
    if (g_tma_enabled && p.b_kind == OP_PLAIN_K && tma_ok_2d(p.B, p.b_ld)) {
        cuuint64_t d[2] = {(cuuint64_t)a->K, (cuuint64_t)a->N}; cuuint64_t s1[1] = {(cuuint64_t)p.b_ld * 2}; cuuint32_t box[2] = {64, (cuuint32_t)p.BN};
        if (encode_map(&mapB, p.B, 2, d, s1, box)) { p.b_kind = OP_TMA_K; p.b_tx = p.BN * 128; }
    } else if (g_tma_enabled && p.b_kind == OP_PLAIN_MN && tma_ok_2d(p.B, p.b_ld)) {
        cuuint64_t d[2] = {(cuuint64_t)a->N, (cuuint64_t)a->K}; cuuint64_t s1[1] = {(cuuint64_t)p.b_ld * 2}; cuuint32_t box[2] = {64, 64};
        if (encode_map(&mapB, p.B, 2, d, s1, box)) { p.b_kind = OP_TMA_MN; p.b_tx = cdiv(p.BN, 64) * 64 * 128; }
    }
    {
        static int dbg = -1;
        if (dbg < 0) { const char* e = getenv("AVEC_DEBUG_ROWOFS"); dbg = e ? atoi(e) : 0; }
        if (dbg > 0 && p.a_kind == OP_TMA_K && a->mode == AVEC_GEMM_PLAIN) p.dbg_rowofs = dbg;
        static int dbg_mode = -1;
        if (dbg_mode < 0) { const char* e = getenv("AVEC_DEBUG_MODE"); dbg_mode = e ? atoi(e) : 0; }
        p.dbg_mode = dbg_mode;
    }
    if (p.a_rows == 0) p.a_rows = halo ? 0 : BM;
    if (p.b_rows == 0) p.b_rows = is_mn(p.b_kind) ? cdiv(p.BN, 64) * 64 : p.BN;

    int split = (a->epi == AVEC_EPI_ACCUM && a->split_k > 1) ? a->split_k : 1;
    if (split > p.num_kb) split = p.num_kb;
    p.kb_per_split = cdiv(p.num_kb, split);
    split = cdiv(p.num_kb, p.kb_per_split);
    const int stage_bytes = p.a_rows * 128 + cdiv(p.b_rows * 128, 1024) * 1024;
    const bool any_gather = !is_tma(p.a_kind) || !is_tma(p.b_kind);
    const size_t ctrl_bytes = 256 + BM * sizeof(RowInfo) + 256 * sizeof(int) + 256 * sizeof(float) + 1024;
    size_t smem;
    int ctas_per_sm = 1;
    // second epilogue warpgroup for tiles wider than one 64-column slab (the epilogue of a 128 x 256 tile is 4 slabs of ~1.3 us)
    static int epi2 = -1;
    if (epi2 < 0) { const char* e = getenv("AVEC_EPI2"); epi2 = e ? atoi(e) : 1; }
    int epi_groups = (epi2 && p.BN > 64) ? 2 : 1;
    const long long tiles_for_cta2 = (long long)grid_m * cdiv(a->N, p.BN) * split;
    if (any_gather) {
        // one tile per CTA; the epilogue staging aliases the drained ring; <= 32 KB stages allow 2 CTAs / SM
        p.stages = stage_bytes <= 32 * 1024 ? 3 : 4;   // >= LAG + 1 = 3 (cp.async run-ahead)
        while (p.stages > 3 && (size_t)p.stages * stage_bytes + ctrl_bytes > 227 * 1024) --p.stages;
        p.stg_dedicated = 0;
        smem = (size_t)p.stages * stage_bytes + ctrl_bytes;
        if ((size_t)p.stages * stage_bytes < (size_t)STG_BYTES * epi_groups) epi_groups = 1;
        if (smem > 227 * 1024 || (size_t)p.stages * stage_bytes < STG_BYTES) return AVEC_ERR_UNSUPPORTED;
    } else {
        // persistent: one CTA per SM, as many ring stages as fit beside the dedicated epilogue staging
        p.stg_dedicated = 1;
        p.stages = 8;
        const size_t stg_total = (size_t)STG_BYTES * epi_groups;
        while (p.stages > 2 && (size_t)p.stages * stage_bytes + stg_total + ctrl_bytes > 227 * 1024) --p.stages;
        const size_t halo_total = 2 * (size_t)p.cpb * p.halo_bytes;
        while (p.stages > 2 && halo_total + (size_t)p.stages * stage_bytes + stg_total + ctrl_bytes > 227 * 1024) --p.stages;
        smem = halo_total + (size_t)p.stages * stage_bytes + stg_total + ctrl_bytes;
        if (smem > 227 * 1024) return AVEC_ERR_UNSUPPORTED;
        // narrow tiles (BN <= 64: a k-block is only 128 tensor-pipe cycles, the epilogue and the per-k-block handshakes are
        // latency bound): two co-resident CTAs per SM interleave their MMA streams and run two epilogues at a time
        static int cta2 = -1;
        if (cta2 < 0) { const char* e = getenv("AVEC_CTA2"); cta2 = e ? atoi(e) : 1; }
        if (cta2 && !halo && p.BN <= 64) {
            int st2 = 6;
            while (st2 > 3 && (size_t)st2 * stage_bytes + STG_BYTES + ctrl_bytes > 112 * 1024) --st2;
            if ((size_t)st2 * stage_bytes + STG_BYTES + ctrl_bytes <= 112 * 1024 && tiles_for_cta2 > num_sms_cached()) {
                p.stages = st2; ctas_per_sm = 2;
                smem = (size_t)p.stages * stage_bytes + STG_BYTES + ctrl_bytes;
            }
        }
    }
    if (!smem_attr_once(3, reinterpret_cast<const void*>(gemm_tc_kernel), 227 * 1024)) return AVEC_ERR_LAUNCH;
    // ---- fast epilogue: launch-constant eligibility (see epilogue_fast)
    {
        static int fast_on = -1;
        if (fast_on < 0) { const char* e = getenv("AVEC_EPI_FAST"); fast_on = e ? atoi(e) : 1; }
        const EpiParams& ep = p.ep;
        int al = 16;
        // widths of 4 mod 8 (D = 180) take the fast path too: their last column group is handled with 8-byte accesses
        bool ok = fast_on && !halo && a->N % 4 == 0 && !p.out_transposed;
        if (ep.kind == AVEC_EPI_ACCUM) {
            ok = ok && ep.out_dtype == AVEC_F32 && (reinterpret_cast<uintptr_t>(ep.out) % 16) == 0 && ep.ldo % 4 == 0 && !ep.colstats;
        } else {
            ok = ok && ep.out_dtype == AVEC_BF16 && ep.kind >= AVEC_EPI_LINEAR && ep.kind <= AVEC_EPI_RELU;
            al = std::min(al, ptr_align(ep.out, ep.ldo));
            if (ep.aux) { ok = ok && ep.aux_dtype == AVEC_BF16; al = std::min(al, ptr_align(ep.aux, ep.ldaux)); }
            if (ep.out2) { ok = ok && ep.out2_dtype == AVEC_BF16 && ep.kind == AVEC_EPI_SWISH; al = std::min(al, ptr_align(ep.out2, ep.ldo2)); }
            if (ep.colstats) ok = ok && a->N % 8 == 0 && ep.kind == AVEC_EPI_LINEAR && (reinterpret_cast<uintptr_t>(ep.colstats) % 16) == 0;
            if ((ep.kind == AVEC_EPI_RESIDUAL || ep.kind == AVEC_EPI_DSWISH) && !ep.aux) ok = false;
        }
        p.epi_fast = ok && al >= 8 ? (al >= 16 ? 2 : 1) : 0;
    }
    if (dc && (!p.epi_fast || p.b_kind != OP_TMA_K)) return AVEC_ERR_UNSUPPORTED;
    if (p.ep.drop_rng && (!p.epi_fast || p.rm_on)) return AVEC_ERR_UNSUPPORTED;   // fused dropout lives in the fast epilogue only
    p.grid_m = grid_m; p.grid_n = cdiv(a->N, p.BN); p.splits = split;
    const long long tiles = (long long)p.grid_m * p.grid_n * p.splits;
    if (tiles > 0x7fffffffLL) return AVEC_ERR_INVALID;
    const int num_sms = num_sms_cached();
    // persistent launch when both operands are TMA-fed (the gather producers double as epilogue warps, so gather kinds keep
    // one tile per CTA); two CTAs per SM when shared memory and TMEM (2 x BN columns each) allow it
    long long ctas = any_gather ? tiles : std::min<long long>(tiles, (long long)num_sms * ctas_per_sm);
    // programmatic dependent launch (AVEC_PDL_SMALL_ONLY does not apply: a parked GEMM grid is one CTA per SM): this grid may be
    // scheduled while the preceding kernel of the stream drains; its CTAs set up barriers / TMEM / descriptors, then pdl_wait()
    if (avec_pdl_for_stream(st)) {
        avec_launch_pdl(gemm_tc_kernel, dim3((unsigned)ctas), dim3(epi_groups == 2 ? TC_THREADS_WIDE : TC_THREADS), smem, st, true, p, mapA, mapB);
        AVEC_LAUNCH_CHECK();
        return AVEC_OK;
    }
    gemm_tc_kernel<<<(unsigned)ctas, epi_groups == 2 ? TC_THREADS_WIDE : TC_THREADS, smem, st>>>(p, mapA, mapB);
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}

// ---- visual stem entry points -----------------------------------------------------------------------------------------
static bool stem_geom_ok(int Nb, int T, int H, int W) {
    return Nb > 0 && T > 0 && H >= 8 && W == 88 && H % 2 == 0 && (long long)Nb * T * cdiv((H / 2) * (W / 2), 128) < 0x7fffffffLL;
}
static void stem_fill(StemParams& p, const void* x, int Nb, int T, int H, int W) {
    memset(&p, 0, sizeof(p));
    p.x = reinterpret_cast<const bf16*>(x);
    p.Nb = Nb; p.T = T; p.H = H; p.W = W; p.Ho = (H - 1) / 2 + 1; p.Wo = (W - 1) / 2 + 1;
    p.tpf = cdiv(p.Ho * p.Wo, 128);
    p.total_tiles = Nb * T * p.tpf;
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("AVEC_STEM_DBG"); dbg = e ? atoi(e) : 0; }
    p.dbg = dbg;
}

extern "C" int avec_stem3d_fwd(const void* x, const void* wp, const float* bias, void* out, float* colstats, int Nb, int T, int H, int W,
                               avec_stream_t stream) {
    AVEC_CHECK_ARG(x && wp && out && stem_geom_ok(Nb, T, H, W));
    AVEC_CHECK_ARG((reinterpret_cast<uintptr_t>(x) % 16) == 0 && (reinterpret_cast<uintptr_t>(wp) % 16) == 0 && (reinterpret_cast<uintptr_t>(out) % 16) == 0);
    AVEC_CHECK_ARG(!colstats || (reinterpret_cast<uintptr_t>(colstats) % 16) == 0);
    StemParams p;
    stem_fill(p, x, Nb, T, H, W);
    p.tc.BN = 64; p.tc.N = 64; p.tc.M = (int)std::min<long long>((long long)Nb * T * p.Ho * p.Wo, 0x7fffffffLL); p.tc.epi_fast = 2;
    EpiParams& ep = p.tc.ep;
    ep.M = p.tc.M; ep.N = 64; ep.kind = AVEC_EPI_LINEAR; ep.alpha = 1.0f; ep.bias = bias;
    ep.out = out; ep.out_dtype = AVEC_BF16; ep.ldo = 64; ep.colstats = colstats;
    CUtensorMap mapW;
    cuuint64_t d[2] = {(cuuint64_t)ST_KPAD, 64}; cuuint64_t s1[1] = {(cuuint64_t)ST_KPAD * 2}; cuuint32_t box[2] = {64, 64};
    if (!encode_map(&mapW, wp, 2, d, s1, box)) return AVEC_ERR_DRIVER;
    if (!smem_attr_once(4, reinterpret_cast<const void*>(stem3d_fwd_kernel), (int)ST_FWD_SMEM)) return AVEC_ERR_LAUNCH;
    const int ctas = std::min(p.total_tiles, num_sms_cached());
    stem3d_fwd_kernel<<<ctas, ST_FWD_THREADS, ST_FWD_SMEM, as_stream(stream)>>>(p, mapW);
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}

extern "C" int avec_stem3d_wgrad(const void* x, const void* dy, float* dw, int Nb, int T, int H, int W, avec_stream_t stream) {
    AVEC_CHECK_ARG(x && dy && dw && stem_geom_ok(Nb, T, H, W));
    AVEC_CHECK_ARG((reinterpret_cast<uintptr_t>(x) % 16) == 0 && (reinterpret_cast<uintptr_t>(dy) % 16) == 0);
    StemParams p;
    stem_fill(p, x, Nb, T, H, W);
    p.dw = dw;
    CUtensorMap mapY;
    const cuuint64_t sites = (cuuint64_t)p.Ho * p.Wo;
    cuuint64_t d[3] = {64, sites, (cuuint64_t)Nb * T}; cuuint64_t st[2] = {128, sites * 128}; cuuint32_t box[3] = {64, 128, 1};
    if (!encode_map(&mapY, dy, 3, d, st, box)) return AVEC_ERR_DRIVER;
    if (!smem_attr_once(5, reinterpret_cast<const void*>(stem3d_wgrad_kernel), (int)ST_WG_SMEM)) return AVEC_ERR_LAUNCH;
    const int ctas = std::min(p.total_tiles, num_sms_cached());
    stem3d_wgrad_kernel<<<ctas, ST_WG_THREADS, ST_WG_SMEM, as_stream(stream)>>>(p, mapY);
    AVEC_LAUNCH_CHECK();
    return AVEC_OK;
}
