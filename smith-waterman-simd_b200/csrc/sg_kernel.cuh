// sg_kernel.cuh -- the adaptive-banded X-drop semi-global aligner on sm_100a.
//
// What it computes: exactly the result of the reference's
//   SemiGlobal_AdaptiveBanded_XDrop_111_32_70(seq1, seq2)      (/root/reference/source.cpp:1836-1976)
// and therefore of its AVX2 forms _simd, _simd_mark2/3/4 (source.cpp:1978-2725), which the
// reference asserts equal to it (TestSemiGlobal, source.cpp:2774-2784): match/mismatch/gap =
// 1/1/1, a band of 32 cells on an anti-diagonal that moves right or down each round
// (right iff result[0] < result[31], source.cpp:1883-1906), X-drop threshold 70 applied per cell
// against the best score so far (source.cpp:1933-1936), alignment anchored at (0,0), free end at
// the FIRST round that reaches the best score and, within it, the upper-right-most such cell
// (source.cpp:1928-1931, 1953-1954), traceback preferring diagonal > up > left (source.cpp:1958-1971).
//
// This is not a translation of the AVX2 code.  The mapping is the one the band suggests on a GPU:
//   * THE BAND IN ONE LANE (or two), as packed int16x2 registers (sg2_core.cuh): a warp advances 32 (16) pairs per
//     round.  The reference's byte shifts across a 256-bit register (alignr/permute2x128, source.cpp:2622,2632)
//     are funnel shifts inside a lane; its five-step horizontal max (source.cpp:2656-2660) is a chain of packed
//     three-input max.
//   * Values live in the X-drop frame (value - max(best - 70, 1)), as in the reference's 8-bit AVX2 forms
//     (offset_diff, source.cpp:2661-2665), so int16 holds at any length; a dropped cell is a sentinel that one
//     unsigned minimum produces.
//   * The reference keeps the whole band history (1 MB of uint8 per pair, source.cpp:2591) and re-derives each
//     traceback step by comparing scores.  Here the forward pass records, per round and cell, the OUTCOME of those
//     comparisons -- "this cell's value came from the diagonal" / "... from above", evaluated with the reference's
//     own comparisons and preference order (a 2-bit tag per cell) -- plus two bits that say which way the band moved:
//     16 bytes per round.
//   * The traceback is a second kernel, ONE THREAD PER PAIR: the walk is serial (as in the reference), so a warp
//     spent on it would execute every instruction for one live lane.  The records of 32 pairs are interleaved
//     round by round, and the 32 walkers of a warp march down the ROUNDS together (a walker acts in the rounds its
//     path visits), so every record fetch of the warp is one contiguous 512-byte read.  One op per step:
//     0 = diagonal, 1 = down (y+1), 2 = right (x+1).  The ops land right-aligned in the pair's output row in
//     forward order and a third kernel shifts each row to the left edge.
//
// HBM layout: seq1, seq2 [n][len] byte codes 0..3 (the reference's std::array<uint8_t,16384>, len = 16384);
// scratch: per GROUP of 32 pairs of a launch the round records [round][32 pairs][16 bytes] (sg_group_bytes), plus
// one spare group; outputs
// score/end_y/end_x/n_ops [n] int32 and ops [n][2*len] bytes (optional).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "sg2_core.cuh"

namespace swb {

// rounds are numbered 0 .. 2*len (MAX_ROUND = (len+1)*2-1, source.cpp:1872); rounded up to a chunk of 32
__host__ __device__ inline uint32_t sg_rounds_cap(int len) { return ((uint32_t)(2 * len + 1) + 31u) & ~31u; }
// per pair of a launch: [rounds_cap] uint4 records (four lane words, sg2_core.cuh); record 0 = where the traceback starts
__host__ __device__ inline size_t sg_trace_bytes(int len) { return (size_t)sg_rounds_cap(len) * 16; }
// the records of pairs 32 g .. 32 g + 31 form one group, interleaved: [round][pair & 31] uint4
constexpr int SG_GROUP = 32;
__host__ __device__ inline size_t sg_group_bytes(int len) { return sg_trace_bytes(len) * SG_GROUP; }
__host__ __device__ inline size_t sg_groups_for(unsigned long long pairs) { return (size_t)((pairs + SG_GROUP - 1) / SG_GROUP) + 1; }   // + the spare group

struct SgOut {
    int32_t* score;     // [n]  best score (offset removed)
    int32_t* end_y;     // [n]  best cell, 0..len
    int32_t* end_x;     // [n]
    int32_t* n_ops;     // [n]  traceback length; nullable together with ops
    uint8_t* ops;       // [n][2*len], nullable: score and end cell only
};

// ---------------------------------------------------------------------------------------------
// Forward kernel (sg2_core.cuh): four lanes per pair and eight pairs per warp (NW = 4), or two lanes per pair and
// sixteen pairs per warp (NW = 8).  The lane groups of a warp run their pairs side by side and pick up the next ones
// together (pairs of similar length finish together; the bench's and the reference's pairs all run the full 2*len rounds).
constexpr int SG2_THREADS = 32;
// Words per lane of the forward kernel, chosen per launch by the batch size (sg_words_for): 16 -- the whole band in one lane,
// 32 pairs per warp, no shuffles, 360 instructions per warp-round = 11.3 per pair -- once the batch gives every scheduler
// a warp of 32 pairs; below that 8 -- two lanes per pair, sixteen pairs per warp, 224 = 14 per pair -- which has twice
// the warps.  Measured (alignments/s, 16384-mers): 8192 pairs 599 k / 784 k (16 / 8 words); 18 944 pairs 1.24 M / 1.21 M;
// 37 888 pairs 1.42 M with 16 words.  Four lanes per pair (4 words, 158 = 19.8 per pair) never won: 759 k and 1.06 M.
__host__ __device__ inline int sg_words_for(unsigned long long pairs, int sm_count) { return pairs >= (unsigned long long)sm_count * 128ull ? 16 : 8; }

template <int LANES>
struct Sg2DevEnv {
    int lane_; uint32_t one_;
    __device__ __forceinline__ int q() const { return lane_; }
    __device__ __forceinline__ uint32_t one() const { return one_; }
    __device__ __forceinline__ uint32_t shfl(uint32_t v, int src) const { return __shfl_sync(0xffffffffu, v, src, LANES); }
    __device__ __forceinline__ uint32_t shfl_xor(uint32_t v, int m) const { return __shfl_xor_sync(0xffffffffu, v, m, LANES); }
};

template <bool RECORD, int NW>         // RECORD false: score and end cell only, no round records are written; NW: words per lane
__global__ void __launch_bounds__(SG2_THREADS)
sg2_xdrop_kernel(const uint8_t* __restrict__ seq1, const uint8_t* __restrict__ seq2, const int len, const unsigned long long n,
                 uint4* __restrict__ traces, const SgOut out, const uint32_t one)
{
    constexpr int LANES = Sg2State<NW>::kLanes, PPW = 32 / LANES;      // lanes per pair, pairs per warp
    const unsigned lane = threadIdx.x & 31u;
    Sg2DevEnv<LANES> env{(int)(lane & (LANES - 1)), one};
    const unsigned long long warp = ((unsigned long long)blockIdx.x * SG2_THREADS + threadIdx.x) >> 5;
    const unsigned long long n_warps = ((unsigned long long)gridDim.x * SG2_THREADS) >> 5;
    const uint32_t rounds_cap = sg_rounds_cap(len);
    const int max_round = 2 * len + 1;          // rounds run while round < MAX_ROUND (source.cpp:1872,1886)
    // The warp stays converged (full-mask shuffles): its lane groups take PPW consecutive pairs, run their rounds
    // together until the last of them is done, then take the next PPW.  A group beyond the batch shadows the last
    // pair and writes its records to the spare group of the scratch.
    for (unsigned long long base = warp * PPW; base < n; base += n_warps * PPW) {
        const unsigned long long want = base + lane / LANES;
        const bool live = want < n;
        const unsigned long long p = live ? want : n - 1ull;
        const uint8_t* const s1 = seq1 + p * (unsigned long long)len;
        const uint8_t* const s2 = seq2 + p * (unsigned long long)len;
        const unsigned long long slot = live ? want : ((n + SG_GROUP - 1) / SG_GROUP) * SG_GROUP + (want & (SG_GROUP - 1));
        uint32_t* const rec_row = reinterpret_cast<uint32_t*>(traces + (slot / SG_GROUP) * rounds_cap * SG_GROUP + (slot & (SG_GROUP - 1)));
        Sg2State<NW> s;
        sg2_init(s, env, s1, s2, len);
        const uint8_t* const role = env.q() == 0 ? s1 : s2;
        for (int round = 1; round < max_round; ++round) {
            const bool go = sg2_round<RECORD>(s, env, role, len, round, rec_row, 4 * SG_GROUP, s2);
            if ((round & 3) == 0 && !__any_sync(0xffffffffu, go)) break;      // a finished pair stays finished: asking every fourth round is enough
        }
        int32_t score, end_y, end_x, best_round, loc;
        sg2_finish(s, env, score, end_y, end_x, best_round, loc);
        if (live && env.q() == 0) {
            if (RECORD) *reinterpret_cast<uint4*>(rec_row) = make_uint4((uint32_t)best_round, (uint32_t)loc, (uint32_t)end_y, (uint32_t)end_x);   // record 0: where the traceback starts
            out.score[p] = score; out.end_y[p] = end_y; out.end_x[p] = end_x;
        }
    }
}

// Traceback (source.cpp:1956-1973) over the recorded tags, one thread per pair, the 32 walkers of a warp in step BY
// ROUND: the warp marches from the highest best round of its 32 pairs down to round 1, and a walker acts in round r
// when its path stands in round r (a diagonal step skips one round, a finished or not yet started walker idles).
// With the records of the warp's 32 pairs interleaved round by round, lane i's record of round r sits at
// group[r][i]: the warp's fetch of a round is ONE contiguous 512-byte read, streamed through a shared-memory ring
// SG_TB_DEPTH rounds ahead with cp.async (eight rounds per commit group).  Each lane reads back only what it wrote
// into the ring itself, so no barrier is needed.
// Step k (k = 0 is the LAST move) is written to row[2*len - 1 - k], so the row ends up holding the forward-ordered op
// string right-aligned; sg_left_align_kernel then moves it to the left edge.
constexpr int SG_TB_THREADS = 64;
constexpr int SG_TB_DEPTH = 32;
constexpr size_t SG_TB_SMEM = (size_t)SG_TB_DEPTH * SG_TB_THREADS * sizeof(uint4);    // 32 KiB: seven blocks per SM (64 KiB allowed three, and a batch of 37888 pairs ran in two waves)

__device__ __forceinline__ void sg_cp_async16(uint4* smem_dst, const uint4* gmem_src)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(s), "l"(gmem_src) : "memory");
}

// The moves leave the walker EIGHT AT A TIME (WIDE: rows of a multiple of 8 bytes at an 8-byte aligned base, i.e. every
// length that is a multiple of 4).  A walker's moves go to descending addresses of its own row, so a byte store per step
// is 32 lanes writing one byte each into 32 different sectors: 680 M partial-sector writes for the 37 888-pair batch, as
// many L2 requests as the record stream itself has sectors -- the stores, not the 20 GB of records, were what the kernel
// waited for.  The last eight moves now ride in a 64-bit register (shift in at the bottom: the first of them ends up in the
// top byte = the highest address) and go out as one aligned 8-byte store; the row's last partial word is flushed bytewise.
template <int NW, bool WIDE>    // NW: the forward kernel's words per lane (the record layout)
__global__ void __launch_bounds__(SG_TB_THREADS)
sg_traceback_kernel(const uint4* __restrict__ traces, const int len, const unsigned long long n, const SgOut out)
{
    extern __shared__ uint4 sg_ring[];          // [warp of the block][round & (DEPTH-1)][lane]
    const unsigned lane = threadIdx.x & 31u;
    const unsigned long long g = ((unsigned long long)blockIdx.x * SG_TB_THREADS + threadIdx.x) >> 5;      // group = warp
    if (g * SG_GROUP >= n) return;              // the whole warp
    const unsigned long long p = g * SG_GROUP + lane;
    const bool live = p < n;
    const uint32_t rounds_cap = sg_rounds_cap(len);
    const uint32_t cap = 2u * (uint32_t)len;
    const uint4* const grp = traces + g * rounds_cap * SG_GROUP + lane;                 // round r: grp[r * 32]
    uint4* const ring = sg_ring + (threadIdx.x >> 5) * (SG_TB_DEPTH * 32) + lane;       // round r: ring[(r & 63) * 32]
    uint8_t* const row = out.ops + (live ? p : 0ull) * (unsigned long long)cap;

    const uint4 start = live ? grp[0] : make_uint4(0u, 31u, 0u, 0u);      // {best round, band element of the end cell, end_y, end_x}
    int rw = (int)start.x, o = (int)start.y;    // the walker stands on element o of round rw
    int rtop = rw;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) rtop = max(rtop, __shfl_xor_sync(0xffffffffu, rtop, d));
    rtop |= 7;                                  // blocks of eight rounds, aligned: rtop, rtop - 8, ...
#pragma unroll 1
    for (int k = 0; k < SG_TB_DEPTH; k += 8) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int r = rtop - k - q;
            if (r >= 1) sg_cp_async16(ring + (r & (SG_TB_DEPTH - 1)) * 32, grp + (size_t)r * 32);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    uint32_t n_ops = 0;
    unsigned long long acc = 0ull;              // WIDE: the moves since the last 8-byte store, newest in the low byte
#pragma unroll 1
    for (int rb = rtop; rb >= 1; rb -= 8) {
        asm volatile("cp.async.wait_group %0;" :: "n"(SG_TB_DEPTH / 8 - 1) : "memory");      // this block of rounds has landed
        uint4 rec[8];                           // the eight records first (their addresses do not depend on the walk) ...
#pragma unroll
        for (int q = 0; q < 8; ++q) rec[q] = ring[((rb - q) & (SG_TB_DEPTH - 1)) * 32];
#pragma unroll
        for (int q = 0; q < 8; ++q) {           // ... then the dependent chain: element -> lane word -> tag -> next element,
            const int r = rb - q;               //     branch-free (selects and one predicated store)
            const bool act = rw == r && r >= 1 && n_ops < cap;
            int o2 = o, r2 = rw;
            const uint32_t op = sg2_tb_step<NW>(rec[q].x, rec[q].y, rec[q].z, rec[q].w, o2, r2);                                 // 0 = diagonal, 1 = down, 2 = right
            if (WIDE) {
                acc = act ? ((acc << 8) | (unsigned long long)op) : acc;
                if (act && (n_ops & 7u) == 7u) *reinterpret_cast<unsigned long long*>(row + (cap - 1u - n_ops)) = acc;   // moves n_ops-7 .. n_ops: bytes cap-1-n_ops .. +7
            } else {
                if (act) row[cap - 1u - n_ops] = (uint8_t)op;
            }
            o = act ? o2 : o;
            rw = act ? r2 : rw;
            n_ops += act ? 1u : 0u;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {           // refill the slots just read with the rounds DEPTH further down
            const int r = rb - SG_TB_DEPTH - q;
            if (r >= 1) sg_cp_async16(ring + (r & (SG_TB_DEPTH - 1)) * 32, grp + (size_t)r * 32);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (WIDE && live) {                          // the last n_ops % 8 moves: the low bytes of acc, oldest on top
        const uint32_t r = n_ops & 7u;
        for (uint32_t k = 0; k < r; ++k) row[cap - n_ops + k] = (uint8_t)(acc >> (8u * k));
    }
    if (live) out.n_ops[p] = (int32_t)n_ops;
}

// Moves each row's op string from the right edge (where the traceback left it) to the left edge.  One block per
// row; a step reads 256 x 16 bytes into registers, synchronises, and writes them `shift` bytes lower: the
// destination of a step never reaches the source of a later one, so an overlapping move is safe.
// Rows of a multiple of 16 bytes move as aligned words: five 32-bit loads, four byte-permutes (the shift modulo 4)
// and one 16-byte store per thread and step; up to three bytes past the string may be overwritten (they are
// unspecified: the string's length is n_ops).  Other rows move byte by byte.
__global__ void __launch_bounds__(256)
sg_left_align_kernel(const int len, const unsigned long long n, const SgOut out)
{
    const unsigned long long p = blockIdx.x;
    if (p >= n) return;
    const uint32_t cap = 2u * (uint32_t)len;
    const uint32_t n_ops = (uint32_t)out.n_ops[p];
    if (n_ops == 0u || n_ops >= cap) return;
    uint8_t* const row = out.ops + p * (unsigned long long)cap;
    const uint32_t shift = cap - n_ops;
    if ((cap & 15u) == 0u && (reinterpret_cast<uintptr_t>(out.ops) & 15u) == 0u) {
        uint32_t* const row32 = reinterpret_cast<uint32_t*>(row);
        const uint32_t s4 = shift >> 2, sel = 0x3210u + 0x1111u * (shift & 3u), last = cap / 4u - 1u;
        const uint32_t n_words = (n_ops + 3u) >> 2;
        for (uint32_t base = 0; base < n_words; base += 256u * 4u) {
            const uint32_t k0 = base + threadIdx.x * 4u;
            uint32_t src[5];
#pragma unroll
            for (int q = 0; q < 5; ++q) src[q] = (k0 < n_words) ? row32[min(s4 + k0 + q, last)] : 0u;
            __syncthreads();
            if (k0 < n_words)
                *reinterpret_cast<uint4*>(row32 + k0) = make_uint4(__byte_perm(src[0], src[1], sel), __byte_perm(src[1], src[2], sel),
                                                                  __byte_perm(src[2], src[3], sel), __byte_perm(src[3], src[4], sel));
            __syncthreads();
        }
        return;
    }
    for (uint32_t base = 0; base < n_ops; base += 256u * 16u) {
        const uint32_t i0 = base + threadIdx.x * 16u;
        uint8_t b[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) b[q] = (i0 + q < n_ops) ? row[shift + i0 + q] : (uint8_t)0;
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 16; ++q) if (i0 + q < n_ops) row[i0 + q] = b[q];
        __syncthreads();
    }
}

} // namespace swb
