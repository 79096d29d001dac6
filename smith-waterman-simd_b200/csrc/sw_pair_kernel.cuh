// sw_pair_kernel.cuh -- the LATENCY kernels behind swb200_score_pair and small batches.
//
// The reference's per-pair call (SmithWaterman_simdN, source.cpp:462-466) takes 2-4 us on a CPU core.  The
// throughput kernel (sw_kernel.cuh) gives one thread two whole pairs -- 16 384 cells in sequence, ~65 us however few
// pairs there are.  For a handful of pairs the matrix is swept the other way round: ONE WARP PER PAIR, the
// anti-diagonal wavefront across the lanes, the band's boundary row handed from lane to lane by warp shuffle:
//   * lane l owns rows 4l .. 4l+3; at step t it computes the four cells (4l+k, t-4l-k), k = 0..3, which are mutually
//     independent; 255 steps cover the matrix;
//   * row 4l's upper neighbours come from lane l-1's row 4l-1 of the previous step: one __shfl_up per step (the diagonal
//     neighbour is the value shuffled one step earlier).  Row 4l+3 -- the one that is shuffled -- is computed FIRST in a
//     step and shuffled at once, row 4l -- the one that needs the shuffled value -- LAST: a single warp has nobody to hide
//     the shuffle's latency behind but its own other three rows;
//   * the recurrence is the reference's scalar one (source.cpp:50-53) in int32, in the anti-diagonal offset frame of the
//     throughput kernel (exact on the whole parameter domain here: int32 never overflows): two ALU-pipe instructions per
//     cell, see pair_sweep;
//   * the substitution scores come ready-made from a per-warp table in shared memory, S[a][target[c]] as int32 for the
//     four bases a and every column c, built once per pair; a row reads four columns of its base's table row with one
//     LDS.128, one group ahead.  Columns outside [0,128) hold -128: such a cell can never exceed a real neighbour, stays 0
//     left of the matrix, and so needs no predicate anywhere.
//
// Two kernels share that sweep:
//   sw_pair_kernel    one launch per call, one warp per pair (up to 2048 pairs).  The scores go to (mapped, pinned) host
//                     memory as 8-byte stores tagged with the call's sequence number; the host spins on those words
//                     instead of synchronising a stream.
//   sw_pair_server    the per-pair call proper (swb200_score_pair, the shape the reference's SpeedTest times: one pair,
//                     a million calls, source.cpp:3036-3054).  A launch costs ~7 us on this platform whatever the kernel
//                     does, so this one STAYS: a single warp that polls a 320-byte DOORBELL in mapped pinned host memory,
//                     scores the pair it finds there, stores the tagged result to a mapped MAILBOX word and polls again.
//                     It leaves by itself when no call has come for `linger_ns` (it says so in the mailbox; the next call
//                     launches a new one), so it never holds up cudaFree, a device synchronisation or another stream's
//                     work for longer than that -- and the library DISMISSES it (a header-only doorbell message) before
//                     it enqueues anything else on the device.  A call in a run of calls is then: 36 host stores, one PCIe read by the
//                     GPU, the sweep, one PCIe write -- no launch.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace swb {

constexpr int PAIR_WARPS = 4;            // sw_pair_kernel: warps (= pairs) per block
constexpr int PAIR_PAD = 124;            // columns left of the matrix that lane 31 sweeps before it reaches column 0
constexpr int PAIR_COLS = 384;           // PAIR_PAD + 128 + 127 padded up to 4-column groups (+ one group read ahead)
constexpr int PAIR_OUTSIDE = -128;       // substitution score of a column outside the matrix

// The warp's score table: row a holds S[a][target[c]] for every column c as int32, PAIR_OUTSIDE outside the matrix.
struct __align__(16) PairTable { int sc[4][PAIR_COLS]; };

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

// S[a][b] out of t4[a] = bytes S[a][0..3], sign-extended
__device__ __forceinline__ int pair_score_of(uint32_t t4a, uint32_t b) { return (int)(int8_t)(t4a >> (8u * b)); }

// One cell in the offset frame: max3(max(diag^ + s'', up^), left^, Z) -- two ALU-pipe instructions, nothing else
__device__ __forceinline__ int pair_cell(int s, int dg, int up, int left, int z)
{
    return __vimax3_s32(__viaddmax_s32(dg, s, up), left, z);
}

// a * b + c as an IMAD whatever the compiler knows about the operands (the FMA pipe is idle, the ALU pipe is the bound)
__device__ __forceinline__ int pair_imad(int a, int b, int c)
{
    int d;
    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// The sweep of one pair by one warp.  aw / bw: this lane's four query / target bases (bytes; only bits 0-1 are used);
// tab: the warp's score table; the caller has put PAIR_OUTSIDE + 2g into the columns outside [PAIR_PAD, PAIR_PAD + 128)
// (pair_table_init).  Returns the pair's score in every lane.
//
// What bounds it: ONE warp lives on one of the SM's four schedulers, whose ALU pipe takes a warp instruction every second
// cycle, and every cell hangs on the cell above it one step earlier.  So both the count of ALU-pipe instructions per step
// and the length of a cell's dependent chain are the sweep's time:
//   * int32 values in the anti-diagonal OFFSET frame H^ = H + g*t (t = row + column = the step; exact on the whole
//     parameter domain: nothing ever leaves [-2*127, 127*128 + 127*256]).  A gap step costs nothing in that frame, a
//     diagonal step s'' = s + 2g, a true zero is Z = g*t:  H^ = max3(max(diag^ + s'', up^), left^, Z) -- a VIADDMNMX and a
//     VIMNMX3, each waiting only for the one before (the plain recurrence needs max, subtract, add-max-relu);
//   * the table holds s'' ready-made as int32 (an LDS.128 per row and four steps, on the load/store pipe) -- no
//     byte-permute per cell;
//   * everything that is bookkeeping is an IMAD on the FMA pipe: Z += g, best^ += g, the lane-0 masks.
// Per step: 8 (cells) + 2 (running best) ALU-pipe instructions.  tests/test_pair_schedule.py restates this sweep in
// numpy against the oracle.
__device__ __forceinline__ int pair_sweep(uint32_t aw, uint32_t bw, const uint32_t (&t4)[4], int g, PairTable* tab, int lane)
{
    __syncwarp();                                            // the previous pair's table has been read by every lane
    const int g2 = 2 * g;
#pragma unroll
    for (int a = 0; a < 4; ++a)
        *reinterpret_cast<int4*>(&tab->sc[a][PAIR_PAD + 4 * lane]) =
            make_int4(pair_score_of(t4[a], bw & 3u) + g2, pair_score_of(t4[a], (bw >> 8) & 3u) + g2,
                      pair_score_of(t4[a], (bw >> 16) & 3u) + g2, pair_score_of(t4[a], (bw >> 24) & 3u) + g2);
    __syncwarp();

    // row k reads the table row of its base, four columns at a time: group m = columns 4m - 4*lane .. +3
    const int4* row[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) row[k] = reinterpret_cast<const int4*>(&tab->sc[(aw >> (8 * k)) & 3u][PAIR_PAD - 4 * lane]);
    int not_top;                                             // 0 in lane 0, else 1 -- through an asm so that nothing below becomes a select
    asm("min.s32 %0, %1, 1;" : "=r"(not_top) : "r"(lane));
    const int g_top = g - g * not_top;                       // g in lane 0, else 0
    // step -1 and -2: true zeros in those steps' frames
    int h1[4] = {-g, -g, -g, -g}, h2[4] = {-g2, -g2, -g2, -g2};
    int up0 = -g * not_top, dg0 = -g2;
    int zm = -2 * g_top;                                     // lane 0: Z(t-1) -- what row -1 holds as the next step's diagonal neighbour
    int best = -g, z = -g;
    int4 prv[4], cur[4], nxt[4];                             // row k at step 4m + j is at column 4m + j - k: in cur, or (j < k) in prv
#pragma unroll
    for (int k = 0; k < 4; ++k) { cur[k] = make_int4(0, 0, 0, 0); nxt[k] = row[k][0]; }   // (prv of group 0: columns left of every lane's first, never real;
                                                                                           //  any score <= 2g - 128 would do, see below)
    const int out = PAIR_OUTSIDE + g2;
#pragma unroll
    for (int k = 0; k < 4; ++k) cur[k] = make_int4(out, out, out, out);

    // One step: the substitution scores s0 (row 4l) .. s3 (row 4l+3) of the four cells
    auto step = [&](int s0, int s1, int s2, int s3) {
        z = pair_imad(g, 1, z);                              // the frame moves: Z = g*t
        best = pair_imad(g, 1, best);
        zm = pair_imad(g_top, 1, zm);
        const int n3 = pair_cell(s3, h2[2], h1[2], h1[3], z);
        const int sh = __shfl_up_sync(0xffffffffu, n3, 1);   // next step's upper neighbour of the lane below: on its way early
        const int n2 = pair_cell(s2, h2[1], h1[1], h1[2], z);
        const int n1 = pair_cell(s1, h2[0], h1[0], h1[1], z);
        const int n0 = pair_cell(s0, dg0, up0, h1[0], z);
        best = __vimax3_s32(best, n3, n2);
        best = __vimax3_s32(best, n1, n0);
        dg0 = pair_imad(up0, 1, zm);                         // next step's diagonal neighbour = this step's upper one; lane 0: exactly Z(t-1)
        up0 = pair_imad(sh, not_top, 0);                     // lane 0: row -1 (source.cpp:44); as an UPPER neighbour any value <= Z will do: 0
        h2[0] = h1[0]; h2[1] = h1[1]; h2[2] = h1[2]; h2[3] = h1[3];
        h1[0] = n0; h1[1] = n1; h1[2] = n2; h1[3] = n3;
    };
#pragma unroll 2
    for (int m = 0; m < 64; ++m) {                           // 255 steps cover the matrix; the 256th only touches padding
#pragma unroll
        for (int k = 0; k < 4; ++k) { prv[k] = cur[k]; cur[k] = nxt[k]; nxt[k] = row[k][m + 1]; }   // one group ahead (inside the table for every lane)
        step(cur[0].x, prv[1].w, prv[2].z, prv[3].y);
        step(cur[0].y, cur[1].x, prv[2].w, prv[3].z);
        step(cur[0].z, cur[1].y, cur[2].x, prv[3].w);
        step(cur[0].w, cur[1].z, cur[2].y, cur[3].x);
    }
    return __reduce_max_sync(0xffffffffu, best) - z;         // out of the frame of the last step
}

// The columns outside the matrix: score PAIR_OUTSIDE (+ 2g: the table holds s'').  The matrix's own columns are
// rewritten for every pair by pair_sweep.
__device__ __forceinline__ void pair_table_init(PairTable* tab, int g, int lane)
{
    for (int i = lane; i < 4 * PAIR_COLS; i += 32) tab->sc[0][i] = PAIR_OUTSIDE + 2 * g;
}

__global__ void __launch_bounds__(32 * PAIR_WARPS)
sw_pair_kernel(const PairArgs pa)
{
    __shared__ PairTable tab_s[PAIR_WARPS];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t p = blockIdx.x * PAIR_WARPS + wib;
    if (p >= pa.n) return;                                   // a whole warp leaves together
    PairTable* tab = &tab_s[wib];

    uint32_t aw, bw;                                         // this lane's four query bases / four target bases
    if (pa.inline_pair) { aw = pa.inl1[lane]; bw = pa.inl2[lane]; }
    else {
        aw = reinterpret_cast<const uint32_t*>(pa.seq1 + (size_t)p * 128)[lane];
        bw = reinterpret_cast<const uint32_t*>(pa.seq2 + (size_t)p * pa.seq2_stride)[lane];
    }
    pair_table_init(tab, pa.gap, lane);
    const uint32_t t4[4] = {pa.t4[0], pa.t4[1], pa.t4[2], pa.t4[3]};
    const int best = pair_sweep(aw, bw, t4, pa.gap, tab, lane);
    if (lane == 0) pa.out[p] = ((unsigned long long)pa.seq << 32) | (unsigned long long)(uint32_t)best;
}

// ---------------------------------------------------------------------------------------------- the resident server
// DOORBELL (host writes, GPU polls; mapped pinned, 128-byte aligned).  Every sequence byte carries the low six bits of the
// call's sequence number above its 2-bit code, and the header line carries a checksum, so a poll that catches the host
// half-way through writing a request (or two of the three 128-byte reads of one poll on different sides of it) is simply
// not a request yet: nothing depends on the order in which PCIe reads of different lines see host memory.
struct PairDoorbell {
    uint8_t seq1[128];                   // code | (seq & 63) << 2
    uint8_t seq2[128];
    uint32_t hdr[16];                    // [0..3] t4, [4] gap, [5] checksum, [6] seq (written last); one 64-byte line
};
constexpr uint32_t PAIR_HDR_GAP = 4, PAIR_HDR_SUM = 5, PAIR_HDR_SEQ = 6;
constexpr uint32_t PAIR_SUM_SALT = 0x5bd1e995u;
constexpr uint32_t PAIR_GAP_DISMISS = 0xffffffffu;   // a header with this "gap" is no request: it tells the server to leave now

__host__ __device__ inline uint32_t pair_hdr_checksum(uint32_t t0, uint32_t t1, uint32_t t2, uint32_t t3, uint32_t gap, uint32_t seq)
{
    return t0 ^ (t1 * 3u) ^ (t2 * 5u) ^ (t3 * 7u) ^ (gap * 0x01000193u) ^ (seq * 0x9e3779b1u) ^ PAIR_SUM_SALT;
}

// MAILBOX (GPU writes, host polls; one 64-byte line of mapped pinned memory)
struct PairMailbox {
    unsigned long long result;           // (seq << 32) | (uint32_t)score of the last request served
    uint32_t exit_gen;                   // generation of the last server that has LEFT (written as its last action)
    uint32_t sweep_ns;                   // diagnostics: poll returned -> result stored, for the last request, in SM CYCLES
    uint32_t pad[12];
};

__device__ __forceinline__ uint32_t pair_ld_sys(const void* p)
{
    uint32_t v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

struct PairPoll { uint32_t w1, w2, h; };

// One poll = three coalesced reads across PCIe: the two sequences and the header line.  The loads are only ISSUED here;
// the warp waits for them where pair_handle first looks at the values.
__device__ __forceinline__ void pair_poll_issue(PairPoll& p, const PairDoorbell* db, int lane)
{
    p.w1 = pair_ld_sys(db->seq1 + 4 * lane);
    p.w2 = pair_ld_sys(db->seq2 + 4 * lane);
    p.h = pair_ld_sys(db->hdr + (lane & 15));
}

__global__ void __launch_bounds__(32)
sw_pair_server(const PairDoorbell* db, PairMailbox* mb, uint32_t last_seq, uint32_t gen, unsigned long long linger_ns, uint32_t poll_gap_ns)
{
    __shared__ PairTable tab;
    const int lane = threadIdx.x;
    uint32_t tab_gap = 0xffffffffu;                          // the gap the table's outside columns were made for (none yet)
    unsigned long long t_last;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_last));

    // What a poll brought: 0 = nothing new, 1 = a request (served), 2 = leave.
    auto handle = [&](const PairPoll& p) -> int {
        const uint32_t seq = __shfl_sync(0xffffffffu, p.h, PAIR_HDR_SEQ);
        if (seq != last_seq) {
            const uint32_t t0 = __shfl_sync(0xffffffffu, p.h, 0), t1 = __shfl_sync(0xffffffffu, p.h, 1);
            const uint32_t t2 = __shfl_sync(0xffffffffu, p.h, 2), t3 = __shfl_sync(0xffffffffu, p.h, 3);
            const uint32_t gap = __shfl_sync(0xffffffffu, p.h, PAIR_HDR_GAP), sum = __shfl_sync(0xffffffffu, p.h, PAIR_HDR_SUM);
            const uint32_t tags = (seq & 63u) * 0x04040404u;
            const bool hdr_whole = sum == pair_hdr_checksum(t0, t1, t2, t3, gap, seq);       // (uniform)
            if (hdr_whole && gap == PAIR_GAP_DISMISS) return 2;   // the host is about to use the GPU for something else
            const bool whole = ((p.w1 & 0xfcfcfcfcu) == tags) && ((p.w2 & 0xfcfcfcfcu) == tags) && hdr_whole;
            if (__all_sync(0xffffffffu, whole)) {
                const long long c0 = clock64();
                if (gap != tab_gap) {                        // (a run of calls keeps its gap: once per run)
                    __syncwarp();
                    pair_table_init(&tab, (int)gap, lane);
                    tab_gap = gap;
                }
                const uint32_t t4[4] = {t0, t1, t2, t3};
                const int best = pair_sweep(p.w1, p.w2, t4, (int)gap, &tab, lane);
                if (lane == 0) {
                    const unsigned long long r = ((unsigned long long)seq << 32) | (unsigned long long)(uint32_t)best;
                    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" :: "l"(&mb->result), "l"(r) : "memory");
                    const uint32_t cycles = (uint32_t)(clock64() - c0);          // SM cycles; the host converts with the SM clock
                    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" :: "l"(&mb->sweep_ns), "r"(cycles) : "memory");
                }
                last_seq = seq;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_last));
                return 1;
            }
        }
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        return now - t_last > linger_ns ? 2 : 0;
    };

    // Two polls are in flight at any time, half a round trip apart (poll_gap_ns after a (re)start; each is re-issued when
    // it returns, which keeps the spacing): a request is seen after ~3/4 of a PCIe round trip on average instead of one.
    PairPoll A, B;
    for (;;) {
        pair_poll_issue(A, db, lane);
        __nanosleep(poll_gap_ns);
        int r;
        for (;;) {
            pair_poll_issue(B, db, lane);
            r = handle(A);
            if (r) break;
            pair_poll_issue(A, db, lane);
            r = handle(B);
            if (r) break;
        }
        if (r == 2) break;                                   // (after a request both polls in flight are stale: start over)
    }
    // Leaving: the mailbox says so AFTER the last poll.  A request that arrives from now on finds exit_gen == gen (at once
    // or while it spins) and launches the next server, which serves whatever the doorbell then holds.
    if (lane == 0) asm volatile("st.relaxed.sys.global.u32 [%0], %1;" :: "l"(&mb->exit_gen), "r"(gen) : "memory");
}

} // namespace swb
