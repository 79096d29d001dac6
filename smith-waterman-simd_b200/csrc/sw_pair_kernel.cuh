// sw_pair_kernel.cuh -- the LATENCY kernel behind swb200_score_pair and small batches.
//
// The reference's per-pair call (SmithWaterman_simdN, source.cpp:462-466) takes 2-4 us on a CPU core.  The
// throughput kernel (sw_kernel.cuh) gives one thread two whole pairs -- 16 384 cells in sequence, ~65 us however few
// pairs there are.  For a handful of pairs the matrix is swept the other way round: ONE WARP PER PAIR, the
// anti-diagonal wavefront across the lanes, the band's boundary row handed from lane to lane by warp shuffle:
//   * lane l owns rows 4l .. 4l+3; at step t it computes the four cells (4l+k, t-4l-k), k = 0..3, which are mutually
//     independent; 255 steps cover the matrix;
//   * row 4l's upper neighbours come from lane l-1's row 4l-1 of the previous step: one __shfl_up per step (the diagonal
//     neighbour is the value shuffled one step earlier);
//   * the recurrence is the reference's scalar one (source.cpp:50-53) in plain int32 -- exact on the whole parameter
//     domain, no offset frame: a cell is max(up, left) - g, then max(diag + s, that, 0) as one VIADDMNMX.RELU;
//   * the substitution score is one PRMT: each row keeps S[a][0..3] as four bytes, the column's selector (from a
//     per-warp table in shared memory, built once from the target) picks and sign-extends one.  Columns outside
//     [0,128) select a fifth byte, -128: such a cell can never exceed a real neighbour, stays 0 left of the matrix,
//     and so needs no predicate anywhere.
// The score goes to (mapped, pinned) host memory as one 8-byte store tagged with the call's sequence number; the host
// spins on that word instead of synchronising a stream.  A single pair travels inside the kernel's launch
// parameters, so the call needs no copy at all.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace swb {

constexpr int PAIR_WARPS = 4;            // warps (= pairs) per block
constexpr int PAIR_PAD = 124;            // columns left of the matrix that lane 31 sweeps before it reaches column 0
constexpr int PAIR_SEL_WORDS = 384;      // PAIR_PAD + 128 + 127 padded up

struct PairArgs {
    const uint8_t* seq1;                 // [n][128] byte codes (mapped host memory or device memory), 4-byte aligned
    const uint8_t* seq2;
    unsigned long long* out;             // [n]: (seq << 32) | (uint32_t)score
    uint32_t n;
    uint32_t seq;                        // this call's tag
    uint32_t t4[4];                      // t4[a] = bytes S[a][0..3]
    int32_t gap;
    uint32_t seq2_stride;                // 128, or 0: every pair against the one target at seq2 (SmithWaterman_8b111x32mark1, source.cpp:1227-1234)
    uint32_t inline_pair;                // 1: n == 1 and the pair is in inl1 / inl2
    uint32_t inl1[32], inl2[32];
};

__device__ __forceinline__ uint32_t pair_prmt(uint32_t a, uint32_t b, uint32_t s)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(s));
    return d;
}

__global__ void __launch_bounds__(32 * PAIR_WARPS)
sw_pair_kernel(const PairArgs pa)
{
    __shared__ uint16_t sel_s[PAIR_WARPS][PAIR_SEL_WORDS];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t p = blockIdx.x * PAIR_WARPS + wib;
    if (p >= pa.n) return;                                   // a whole warp leaves together
    uint16_t* sel = sel_s[wib];

    uint32_t aw, bw;                                         // this lane's four query bases / four target bases
    if (pa.inline_pair) { aw = pa.inl1[lane]; bw = pa.inl2[lane]; }
    else {
        aw = reinterpret_cast<const uint32_t*>(pa.seq1 + (size_t)p * 128)[lane];
        bw = reinterpret_cast<const uint32_t*>(pa.seq2 + (size_t)p * pa.seq2_stride)[lane];
    }
    for (int i = lane; i < PAIR_SEL_WORDS; i += 32) sel[i] = 0xCCC4u;          // byte 4 of {profile, 0x80808080}: -128
    __syncwarp();
#pragma unroll
    for (int b = 0; b < 4; ++b) sel[PAIR_PAD + 4 * lane + b] = (uint16_t)(((bw >> (8 * b)) & 3u) * 0x1111u + 0x8880u);
    __syncwarp();

    uint32_t prof[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) prof[k] = pa.t4[(aw >> (8 * k)) & 3u];
    const int g = pa.gap;
    int h1[4] = {0, 0, 0, 0}, h2[4] = {0, 0, 0, 0};
    int up0 = 0, dg0 = 0, best = 0;
    uint32_t w[4] = {0xCCC4u, 0xCCC4u, 0xCCC4u, 0xCCC4u};    // selectors of columns c, c-1, c-2, c-3
    const uint16_t* sp = sel + PAIR_PAD - 4 * lane;          // sp[t] = selector of this lane's column t - 4*lane

    uint32_t nxt = sp[0];
#pragma unroll 4
    for (int t = 0; t < 256; ++t) {                          // 255 steps cover the matrix; the 256th only touches padding
        w[3] = w[2]; w[2] = w[1]; w[1] = w[0]; w[0] = nxt;
        nxt = sp[t + 1];                                     // read one step ahead (t + 1 <= 256 < PAIR_SEL_WORDS - PAIR_PAD)
        dg0 = up0;
        up0 = __shfl_up_sync(0xffffffffu, h1[3], 1);
        if (lane == 0) up0 = 0;                              // row -1: H[0][*] = 0 (source.cpp:44)
        int hn[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int s = (int)pair_prmt(prof[k], 0x80808080u, w[k]);
            const int up = k ? h1[k - 1] : up0;
            const int dg = k ? h2[k - 1] : dg0;
            const int u = max(up, h1[k]) - g;
            hn[k] = __viaddmax_s32_relu(dg, s, u);           // max(dg + s, u, 0)
        }
        best = __vimax3_s32(best, hn[0], hn[1]);
        best = __vimax3_s32(best, hn[2], hn[3]);
#pragma unroll
        for (int k = 0; k < 4; ++k) { h2[k] = h1[k]; h1[k] = hn[k]; }
    }
    best = __reduce_max_sync(0xffffffffu, best);
    if (lane == 0) pa.out[p] = ((unsigned long long)pa.seq << 32) | (unsigned long long)(uint32_t)best;
}

} // namespace swb
