// Helpers shared by the SIMT relative-position attention kernels (attention.cu: whole head resident in shared memory;
// attention_long.cu: key-tiled variant for long sequences).
#pragma once
#include "common.cuh"

namespace {

// Grouped attention (GroupedRelPosMultiHeadSelfAttention.forwardQKV, reference nnet/attentions.py:579-650): a token is G
// consecutive frames concatenated (G * D1 wide, split into H heads of d = G * D1 / H channels); frames past the real
// length Tf are zero rows.  Element e = h * d + c of token `tok` therefore lives in frame tok * G + e / D1, column e % D1.
// G = 1 is the plain layout.  `which` selects q (0), k (1) or v (2) inside the [frames, 3 * D1] matrix.
template <typename T>
__device__ __forceinline__ float fetch_qkv(const T* __restrict__ qkv_b, int tok, int e, int which, int G, int D1, int Tf) {
    const int i = e / D1, col = e - i * D1, frame = tok * G + i;
    return frame < Tf ? ldf(qkv_b + (size_t)frame * 3 * D1 + which * D1 + col) : 0.0f;
}

constexpr int ATT_THREADS = 256;
constexpr int ATT_WARPS = ATT_THREADS / 32;
constexpr int MAX_KPL = 13;  // keys per lane  -> T <= 416
constexpr int MAX_CPL = 5;   // head channels per lane -> d <= 160

// Row stride (elements) of the K / V / E / Q / dO tiles in shared memory.  The tiles are kept in the tensor dtype (bf16 tiles
// are a lossless copy of bf16 tensors and halve the footprint: T = 400 keys of a 64-channel head fit in 227 KB); lanes read
// different rows at the same column, so the stride in 32-bit words must be odd.
template <typename T> __host__ __device__ constexpr int att_ds(int d) { return d + 1; }
template <> __host__ __device__ constexpr int att_ds<bf16>(int d) { return (((d + 1) / 2) % 2 == 1) ? (d + 1) / 2 * 2 : (d + 1) / 2 * 2 + 2; }

}  // namespace
