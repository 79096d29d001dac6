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
//   * ONE WARP PER PAIR, lane i = band element i (31 = upper-right end).  The reference's byte
//     shifts across a 256-bit register (alignr/permute2x128, source.cpp:2622,2632) are SHFL.UP/DOWN,
//     its five-step horizontal max (source.cpp:2656-2660) is one REDUX.MAX, and the direction
//     decision reads lanes 0 and 31.
//   * Values are plain int32 with the reference's +70 offset (0 = dropped / never reached), so the
//     8-bit renormalisation of the AVX2 forms (offset_diff, source.cpp:2661-2665) does not exist.
//   * The reference keeps the whole band history (1 MB of uint8 per pair, source.cpp:2591) and
//     re-derives each traceback step by comparing scores.  Here the forward pass records, per round,
//     two 32-bit masks -- "this cell's value came from the diagonal" / "... from above", evaluated
//     with the reference's own comparisons and preference order -- plus one bit for the band move:
//     8.1 bytes per round instead of 32, written as one coalesced 256-byte store per 32 rounds.
//   * The traceback walks those records backwards, the warp loading them 32 rounds at a time
//     (the walker itself is serial, as in the reference), and emits one op per step:
//     0 = diagonal, 1 = down (y+1), 2 = right (x+1), in forward order from (0,0).
//
// HBM layout: seq1, seq2 [n][len] byte codes 0..3 (the reference's std::array<uint8_t,16384>, len = 16384);
// per resident warp a scratch slot of sg_slot_bytes(len); outputs score/end_y/end_x/n_ops [n] int32
// and ops [n][2*len] bytes (optional).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace swb {

constexpr int SG_BAND = 32;         // BANDWIDTH, source.cpp:1848
constexpr int SG_X = 70;            // X_THRESHOLD, source.cpp:1848
constexpr int SG_WARPS_PER_BLOCK = 4;

// rounds are numbered 0 .. 2*len (MAX_ROUND = (len+1)*2-1, source.cpp:1872); rounded up to a chunk of 32
__host__ __device__ inline uint32_t sg_rounds_cap(int len) { return ((uint32_t)(2 * len + 1) + 31u) & ~31u; }
// per-warp scratch: [rounds_cap] uint2 records, then [rounds_cap/32] move words
__host__ __device__ inline size_t sg_slot_bytes(int len)
{
    const size_t cap = sg_rounds_cap(len);
    return (cap * 8 + cap / 8 + 255) & ~(size_t)255;
}

struct SgOut {
    int32_t* score;     // [n]  best score (offset removed)
    int32_t* end_y;     // [n]  best cell, 0..len
    int32_t* end_x;     // [n]
    int32_t* n_ops;     // [n]  traceback length (= end_y + end_x - number of diagonal steps); nullable with ops
    uint8_t* ops;       // [n][2*len], nullable: score and end cell only
};

__global__ void __launch_bounds__(SG_WARPS_PER_BLOCK * 32)
sg_xdrop_kernel(const uint8_t* __restrict__ seq1, const uint8_t* __restrict__ seq2, const int len, const unsigned long long n,
                uint8_t* __restrict__ scratch, const SgOut out)
{
    const unsigned lane = threadIdx.x & 31u;
    const unsigned long long warp = (unsigned long long)blockIdx.x * SG_WARPS_PER_BLOCK + (threadIdx.x >> 5);
    const unsigned long long n_warps = (unsigned long long)gridDim.x * SG_WARPS_PER_BLOCK;
    const uint32_t rounds_cap = sg_rounds_cap(len);
    uint2* const trace = reinterpret_cast<uint2*>(scratch + warp * sg_slot_bytes(len));
    uint32_t* const moves = reinterpret_cast<uint32_t*>(trace + rounds_cap);
    const int max_round = 2 * len + 1;          // rounds run while round < MAX_ROUND (source.cpp:1872,1886)
    const unsigned FULL = 0xffffffffu;

    for (unsigned long long p = warp; p < n; p += n_warps) {
        const uint8_t* const s1 = seq1 + p * (unsigned long long)len;
        const uint8_t* const s2 = seq2 + p * (unsigned long long)len;
        // seq1p[k] = seq1[k-1], seq2p[k] = seq2[k-32]; everything else is padding (source.cpp:1859-1870).
        // The two pads differ so that pad never "matches" pad: one compare gives the score.
        auto ld_a = [&](int k) -> int { k -= 1;  return ((unsigned)k < (unsigned)len) ? (int)__ldg(s1 + k) : 0xF0; };
        auto ld_b = [&](int k) -> int { k -= 32; return ((unsigned)k < (unsigned)len) ? (int)__ldg(s2 + k) : 0xF1; };

        int res = (lane == 31u) ? SG_X : 0;     // dp[31] = X_THRESHOLD (source.cpp:1877)
        int hor = 0, ver = 0;
        int now_y = 0, now_x = 31;
        int best = SG_X, best_round = 0, best_py = 0, best_lane = 31;
        int ca = ld_a(now_y + 31 - (int)lane), cb = ld_b(now_x - 31 + (int)lane);
        int ca_dn = ld_a(now_y + 1 + 31 - (int)lane);      // this lane's seq1 character if the next move is down
        int cb_rt = ld_b(now_x + 1 - 31 + (int)lane);      // this lane's seq2 character if the next move is right
        uint32_t rec_d = 0, rec_u = 0, movebits = 0;

        int round = 1;
        for (; round < max_round; ++round) {
            // ---- direction (source.cpp:1883-1911)
            const int r0 = __shfl_sync(FULL, res, 0), r31 = __shfl_sync(FULL, res, 31);
            int from_above = __shfl_down_sync(FULL, res, 1);   // result[i+1]
            int from_below = __shfl_up_sync(FULL, res, 1);     // result[i-1]
            if (lane == 31u) from_above = 0;
            if (lane == 0u) from_below = 0;
            const bool right = r0 < r31;
            int diag;
            if (right) {
                diag = ver; hor = res; ver = from_above; cb = cb_rt;
                if (32 + len + 31 < ++now_x) break;
            } else {
                diag = hor; ver = res; hor = from_below; ca = ca_dn;
                if (1 + len < ++now_y) break;
            }
            ca_dn = ld_a(now_y + 1 + 31 - (int)lane);
            cb_rt = ld_b(now_x + 1 - 31 + (int)lane);

            // ---- the 32 cells of the round (source.cpp:1916-1926)
            const int s = (ca == cb) ? 1 : -1;
            const int d = diag ? diag + s : 0;
            const int v = max(max(d, max(hor, ver) - 1), 0);      // a gap from a dropped (0) neighbour gives -1 -> 0
            const int rmax = __reduce_max_sync(FULL, v);
            // what the reference's traceback will find for this cell (source.cpp:1960-1969), diagonal first
            const bool d_ok = (diag != 0) && (v == d);
            const bool u_ok = (ver != 0) && (v == ver - 1);
            const uint32_t dmask = __ballot_sync(FULL, d_ok), umask = __ballot_sync(FULL, u_ok);
            if (best < rmax) {                                    // strict: the FIRST round that reaches the best (source.cpp:1928-1931)
                best = rmax; best_round = round; best_py = now_y;
                best_lane = 31 - __clz(__ballot_sync(FULL, v == rmax));   // upper-right-most cell (source.cpp:1953-1954)
            }
            res = (v < best - SG_X) ? 0 : v;                      // X-drop (source.cpp:1933-1936)

            // ---- record the round
            const unsigned k = (unsigned)round & 31u;
            if (lane == k) { rec_d = dmask; rec_u = umask; }
            movebits |= (right ? 1u : 0u) << k;
            if (k == 31u) {
                trace[round - 31 + (int)lane] = make_uint2(rec_d, rec_u);
                if (lane == 0u) moves[round >> 5] = movebits;
                movebits = 0;
            }
            if (rmax == 0) { ++round; break; }                    // everything dropped (source.cpp:1938-1941)
        }
        // flush the partial chunk that holds round-1 (rounds that were never run hold garbage, never read)
        {
            const int last = round - 1;
            if (last >= 0 && (last & 31) != 31) {
                trace[(last & ~31) + (int)lane] = make_uint2(rec_d, rec_u);
                if (lane == 0u) moves[last >> 5] = movebits;
            }
        }
        __syncwarp();

        const int end_y = best_py + 31 - best_lane;
        const int end_x = (best_round - best_py) - 31 + best_lane;     // pos_x - 31 = round - pos_y (pos_x = 31 + #right moves)
        if (lane == 0u) { out.score[p] = best - SG_X; out.end_y[p] = end_y; out.end_x[p] = end_x; }
        if (!out.ops) continue;

        // ---- traceback (source.cpp:1956-1973) over the recorded masks
        uint8_t* const ops = out.ops + p * 2ull * (unsigned long long)len;
        int r = best_round, y = end_y, x = end_x, py = best_py;
        int c_cur = r >> 5;
        uint2 recA = trace[c_cur * 32 + (int)lane];
        uint32_t mvA = moves[c_cur];
        uint2 recB = make_uint2(0u, 0u);
        uint32_t mvB = 0;
        if (c_cur >= 1) { recB = trace[(c_cur - 1) * 32 + (int)lane]; mvB = moves[c_cur - 1]; }
        uint32_t n_ops = 0, opreg = 0;
        const uint32_t cap = 2u * (uint32_t)len;
        while ((y | x) != 0 && n_ops < cap) {
            if ((r >> 5) != c_cur) {                 // r only ever drops into the previous chunk
                --c_cur;
                recA = recB; mvA = mvB;
                if (c_cur >= 1) { recB = trace[(c_cur - 1) * 32 + (int)lane]; mvB = moves[c_cur - 1]; }
            }
            const int k = r & 31;
            const uint32_t dm = __shfl_sync(FULL, recA.x, k), um = __shfl_sync(FULL, recA.y, k);
            const int o = 31 - (y - py);             // band element of (y,x) in round r (source.cpp:1947)
            const int down_r = ((mvA >> k) & 1u) ? 0 : 1;                                       // round r moved down: pos_y[r] = pos_y[r-1] + 1
            const int down_r1 = (k > 0) ? (((mvA >> (k - 1)) & 1u) ? 0 : 1) : (((mvB >> 31) & 1u) ? 0 : 1);
            uint32_t op;
            if ((dm >> o) & 1u)      { op = 0; --y; --x; py -= down_r + down_r1; r -= 2; }
            else if ((um >> o) & 1u) { op = 1; --y;      py -= down_r;           r -= 1; }
            else                     { op = 2;      --x; py -= down_r;           r -= 1; }
            if (lane == (n_ops & 31u)) opreg = op;
            ++n_ops;
            if ((n_ops & 31u) == 0) ops[n_ops - 32u + lane] = (uint8_t)opreg;      // reversed order for now
        }
        if ((n_ops & 31u) != 0 && lane < (n_ops & 31u)) ops[(n_ops & ~31u) + lane] = (uint8_t)opreg;
        __syncwarp();
        // reverse in place: forward order from (0,0)
        for (uint32_t i = lane; i < n_ops / 2; i += 32) {
            const uint8_t a = ops[i], b = ops[n_ops - 1 - i];
            ops[i] = b; ops[n_ops - 1 - i] = a;
        }
        if (lane == 0u) out.n_ops[p] = (int32_t)n_ops;
        __syncwarp();
    }
}

} // namespace swb
