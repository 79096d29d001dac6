// sg2_core.cuh -- one round of the adaptive-banded X-drop semi-global aligner with the band in packed int16x2
// registers: all 32 cells in one lane (32 pairs per warp) or 16 cells in each of two lanes.
//
// What it computes: exactly SemiGlobal_AdaptiveBanded_XDrop_111_32_70 (/root/reference/source.cpp:1836-1976)
// and hence its AVX2 forms (source.cpp:1978-2725): see sg_kernel.cuh for the statement of the algorithm.
//
// Why this shape.  A round is one serial chain (direction -> shift -> 32 cells -> maximum -> X-drop), so the
// bound is instructions per round.  With a warp per pair every instruction serves ONE cell per lane; here a
// lane owns 16 or 32 cells of the band as int16x2 words, so the cell arithmetic, the shifts and the traceback
// evidence are packed two cells per instruction, and a warp advances 16 or 32 pairs per round.
//   * Values live in the X-drop frame: u = value - T with T = max(best - 70, 1) (the reference's own
//     trick for its 8-bit AVX2 forms, offset_diff at source.cpp:2661-2665), so every live cell is in 0..70
//     and int16 never overflows at any length.  Registers hold 4u; a dropped cell is the sentinel
//     F = 0xC000 (-16384):
//         t  = max(diag + sd, hor + 1, ver + 2)           VIADD.16x2 (FMA pipe) + VIMNMX3.S16x2
//         t2 = t & ~3                                     LOP3
//         R  = umin(max(t2 - 4c, F), F)                   VIADDMNMX.S16x2 + VIMNMX.U16x2
//         R + (1,1), R + (2,2)                            2 x VIADD.16x2 (FMA pipe): next round's hor / ver
//     where sd = 4 (score + 1 - (T's last increment)) + 3, c = 1 + (T's increment this round).  The two
//     low bits are a TAG that rides through the maximum: among equal values the diagonal (3) beats up (2)
//     beats left (1) -- the reference's traceback preference (source.cpp:1960-1969) -- so t & 3 IS the
//     traceback evidence of the cell and no compare is needed.  The unsigned minimum maps every negative
//     value (unsigned >= 0xC000; none is below F) to F and keeps 0..288: the X-drop and the reference's "0 = dropped" in
//     one instruction, with no compare/select.
//   * The band shift of the reference (alignr/permute2x128, source.cpp:2622,2632) is a funnel shift by
//     16 or 0 bits per word (the amount is the direction, so the pairs of a warp do not diverge) plus, with two
//     lanes per pair, ONE shuffle between the lanes that carries the boundary cell and the boundary base together.
//   * The 32 bases of either sequence under the band are bytes in registers (a byte per cell); they
//     shift by 8 or 0 bits; the match score of two cells is one PRMT through an 8-byte table indexed by
//     a XOR b (the reference's pshufb table, source.cpp:2640-2641).
//   * Traceback evidence, not band values, is stored: the 2-bit tag of every cell, gathered with multiply-adds:
//     16 bytes per pair per round, two bits of which say which way the band moved in this round and the one
//     before -- all the traceback needs to follow a cell.
//
// Every function is SWB_HD and templated on an Env that supplies the lane index and the two shuffles, so
// the same text runs on the device (real shuffles) and in tests/emu (four coroutines in lock step).
#pragma once
#include <stdint.h>
#if defined(__CUDACC__)
#include <cuda_fp16.h>
#endif
#include "sw_core.cuh"

namespace swb {

#if defined(__CUDA_ARCH__)
SWB_HD uint32_t vminu2(uint32_t a, uint32_t b) { return __vminu2(a, b); }        // VIMNMX.U16x2
// `keep` unless `take`, then the byte at p: a predicated load, no branch (the quads of a warp stay in step)
SWB_HD uint32_t ld_u8_if(const uint8_t* p, bool take, uint32_t keep)
{
    asm volatile("{ .reg .pred t; setp.ne.u32 t, %2, 0; @t ld.global.nc.u8 %0, [%1]; }" : "+r"(keep) : "l"(p), "r"((uint32_t)take));
    return keep;
}
SWB_HD uint32_t fsl(uint32_t lo, uint32_t hi, uint32_t s) { return __funnelshift_l(lo, hi, s); }   // SHF.L.W
SWB_HD uint32_t fsr(uint32_t lo, uint32_t hi, uint32_t s) { return __funnelshift_r(lo, hi, s); }   // SHF.R.W
#else
SWB_HD uint32_t ld_u8_if(const uint8_t* p, bool take, uint32_t keep) { return take ? (uint32_t)*p : keep; }
SWB_HD uint32_t vminu2(uint32_t a, uint32_t b)
{
    const uint32_t al = a & 0xffffu, bl = b & 0xffffu, ah = a >> 16, bh = b >> 16;
    return (al < bl ? al : bl) | ((ah < bh ? ah : bh) << 16);
}
SWB_HD uint32_t fsl(uint32_t lo, uint32_t hi, uint32_t s) { s &= 31u; return s ? (hi << s) | (lo >> (32u - s)) : hi; }
SWB_HD uint32_t fsr(uint32_t lo, uint32_t hi, uint32_t s) { s &= 31u; return s ? (lo >> s) | (hi << (32u - s)) : lo; }
#endif

constexpr int SG2_X = 70;                          // X_THRESHOLD, source.cpp:1848
constexpr uint32_t SG2_F = 0xC000u;                // dropped / never reached: far below every live value, and far enough
constexpr uint32_t SG2_FF = 0xC000C000u;           //   from -32768 that no sum formed from it wraps around
constexpr uint32_t SG2_UNTAG = 0xFFFCFFFCu;
constexpr uint32_t SG2_PAD_A = 0xC4u;              // base bytes: code * 0x11, seq1 side has bit 7 set; the pads (4, 5)
constexpr uint32_t SG2_PAD_B = 0x55u;              //   differ from every base and from each other (source.cpp:1913-1915)

// NW = packed words per lane: 4 (eight cells per lane, four lanes per pair, eight pairs per warp), 8 (sixteen cells
// per lane, two lanes per pair, sixteen pairs per warp) or 16 (the whole band in one lane, 32 pairs per warp, no shuffles).  The per-round overhead that does not depend on the cells
// (direction, boundary exchange, threshold, loop) is paid once per lane, so the wider lane spends fewer instructions
// per pair; the narrower one has twice the warps for the same batch.
template <int NW>
struct Sg2State {
    static constexpr int kWords = NW, kCells = 2 * NW, kLanes = 32 / (2 * NW), kLast = kLanes - 1;
    uint32_t R1[NW], R2[NW];  // this round's cells kCells q .. (word k = cells 2k | 2k+1 << 16): 4 (value - T), dropped = F,
                          //   tagged 1 (as a left neighbour) and 2 (as an upper neighbour)
    uint32_t H[NW], V[NW];  // the previous round's left / upper neighbours (views of the round before it), tagged 1 / 2
    uint32_t A[NW / 2], B[NW / 2];  // bases under the band: seq1 (byte c = cell c) and seq2
    uint32_t Rb[NW];      // t2 of the best round (to find the end cell)
    uint32_t lut_lo, lut_hi;   // sd table: index 0 = match
    uint32_t right;       // the next round moves right (else down)
    uint32_t got;         // what enters this lane from its neighbour in the next round: bits 0-15 cell, 16-23 base
    uint32_t next_raw;    // first and last lane: the next base to enter the band (code, or 4 / 5 = pad) ...
    uint32_t next2_raw;   // ... and the one after it: a base is loaded two entries before it is used
    uint32_t role_base;   // F | (0x80 << 16 in lane 0)
    int32_t cidx;         // index of next2_raw in seq1 (lane 0) / seq2 (last lane)
    uint32_t nextb_raw, nextb2_raw; int32_t cidxb;      // one lane per pair: the lane feeds both ends; these are the seq2 side
    int32_t pos_y;        // the band's upper-right cell is (pos_y, round - pos_y), source.cpp:1873-1874
    uint32_t prev_down;   // the previous round moved down
    uint32_t dead;        // 1 once a round left every cell <= 0: the reference stops there (source.cpp:1938-1941)
    int32_t best, T, best_round, best_py;
};

SWB_HD int32_t sg2_half(uint32_t w, int hi) { return (int32_t)(int16_t)(hi ? (w >> 16) : (w & 0xffffu)); }

template <int NW>
SWB_HD uint32_t sg2_max_words(const uint32_t* t)
{
    uint32_t m = vmax3(t[0], t[1], t[2]);
#pragma unroll
    for (int w = 3; w + 1 < NW; w += 2) m = vmax3(m, t[w], t[w + 1]);
    return vmax2(m, t[NW - 1]);
}

// Round 0 (source.cpp:1876-1884): cell 31 = X_THRESHOLD, everything else unreached; band at (0, 31).  Round 1
// always moves right (result[0] = 0 < result[31] = 70).
template <int NW, class Env>
SWB_HD void sg2_init(Sg2State<NW>& s, const Env& env, const uint8_t* seq1, const uint8_t* seq2, int len)
{
    constexpr int LAST = Sg2State<NW>::kLast, CELLS = Sg2State<NW>::kCells;
    const int q = env.q();
#pragma unroll
    for (int w = 0; w < NW; ++w) { s.Rb[w] = SG2_FF; s.H[w] = SG2_FF + 0x00010001u; s.V[w] = SG2_FF + 0x00020002u; }
    s.T = 1;                                         // max(70 - 70, 1)
    if (q == LAST) s.Rb[NW - 1] = ((uint32_t)(4 * (SG2_X - 1)) << 16) | SG2_F;
#pragma unroll
    for (int w = 0; w < NW; ++w) { s.R1[w] = s.Rb[w] + 0x00010001u; s.R2[w] = s.Rb[w] + 0x00020002u; }
    // cell i holds seq1p[31 - i] = seq1[30 - i] (i = 31: pad) and seq2p[i] = pad
#pragma unroll
    for (int k = 0; k < NW / 2; ++k) {
        uint32_t a = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int idx = 30 - (CELLS * q + 4 * k + c);
            const uint32_t e = ((unsigned)idx < (unsigned)len) ? ((uint32_t)seq1[idx] * 0x11u | 0x80u) : SG2_PAD_A;
            a |= e << (8 * c);
        }
        s.A[k] = a;
        s.B[k] = SG2_PAD_B * 0x01010101u;
    }
    s.lut_lo = 0x0303030Bu - 0x02020202u; s.lut_hi = 0x03030303u - 0x02020202u;      // round 1 moves right: diag carries tag 2
    s.pos_y = 0; s.prev_down = 0u; s.dead = 0u;
    s.best = SG2_X; s.best_round = 0; s.best_py = 0;
    // bases enter at cell 0 on a down move (lane 0: seq1p[pos_y + 31] = seq1[pos_y + 30]) and at cell 31 on a
    // right move (last lane: seq2p[pos_x] = seq2[pos_x - 32]); round 1 takes seq2[0]
    s.role_base = SG2_F | (q == 0 ? 0x800000u : 0u);
    s.right = 1u;
    s.got = SG2_F | (SG2_PAD_B << 16);
    s.cidx = 32;
    s.next_raw = s.next2_raw = 4u;
    if (q == 0 && 31 < len) s.next_raw = seq1[31];
    if (q == 0 && 32 < len) s.next2_raw = seq1[32];
    s.nextb_raw = s.nextb2_raw = 5u; s.cidxb = 2;
    if (LAST == 0) {                                 // one lane per pair: it keeps the seq1 side above and the seq2 side here
        s.got = SG2_F | ((0 < len ? (uint32_t)seq2[0] : 5u) * 0x110000u);
        if (1 < len) s.nextb_raw = seq2[1];
        if (2 < len) s.nextb2_raw = seq2[2];
    } else if (q == LAST) {
        s.got = SG2_F | ((0 < len ? (uint32_t)seq2[0] : 5u) * 0x110000u);
        s.cidx = 2;
        s.next_raw = (1 < len) ? (uint32_t)seq2[1] : 5u;
        s.next2_raw = (2 < len) ? (uint32_t)seq2[2] : 5u;
    }
}

// One round (source.cpp:1886-1942).  `role_seq` = seq1 in lane 0, seq2 in the last lane (unused elsewhere);
// `rec_row` = this pair's records, 16 bytes per round: round r starts at rec_row[rec_stride * r] (rec_stride = 4 when a
// pair's records are contiguous, 128 when 32 pairs are interleaved round by round).  NW = 4: one word per lane (tags in
// bytes 0 and 2, moves in byte 1).  NW = 8: two words per lane (32 tag bits; moves).  NW = 16: the lane's four words
// (64 tag bits; moves; 0).  With one lane per pair `role_seq` is seq1 and `role_seq2` seq2.  Returns false when every cell is
// <= 0 (source.cpp:1938); a pair in that state is kept inert from then on (see `dead`) -- further rounds change neither
// its best nor its cells -- so the pairs of a warp may keep running together until the last one is done.
// All communication of a round is ONE stage of independent shuffles issued right after the cells are computed (the
// maximum, two for the next direction, two for the boundary cell and base of either direction): everything that
// follows them is lane-local.  They carry t2, the cells BEFORE the X-drop; the receiver applies the drop itself, and the
// direction test is restated on t2:
//     result[0] < result[31]   <=>   t2[0] < t2[31]  and  t2[31] - c >= 0
// (with c the amount subtracted this round; dropped cells compare as the smallest value).
template <bool RECORD, int NW, class Env>
SWB_HD bool sg2_round(Sg2State<NW>& s, Env& env, const uint8_t* role_seq, int len, int round, uint32_t* rec_row, int rec_stride,
                      const uint8_t* role_seq2 = nullptr)
{
    constexpr int LAST = Sg2State<NW>::kLast, NA = NW / 2;
    const int q = env.q();
    const bool right = s.right != 0u;                      // source.cpp:1889
    const uint32_t got = s.got;
    uint32_t D[NW];
#pragma unroll
    for (int w = 0; w < NW; ++w) D[w] = right ? s.V[w] : s.H[w];       // source.cpp:1892,1903
    const uint32_t one = env.one();                        // a 1 the compiler cannot see: it keeps bookkeeping adds and shifts multiply-adds
    const uint32_t gh = got * (one << 16) + 0x00010000u, ga = got << 8, gb = got >> 16, gv = got * one + 2u;
    const uint32_t down = s.right * (0u - one) + one;      // 1 - right, as a multiply-add
    const uint32_t sR = s.right * (one << 4), sD = down * (one << 4), cR = s.right * (one << 3), cD = down * (one << 3);      // shift amounts: 16 / 8 or 0
#pragma unroll
    for (int w = NW - 1; w >= 0; --w) s.H[w] = fsl(w ? s.R1[w - 1] : gh, s.R1[w], sD);
#pragma unroll
    for (int w = 0; w < NW; ++w) s.V[w] = fsr(s.R2[w], w < NW - 1 ? s.R2[w + 1] : gv, sR);
#pragma unroll
    for (int k = NA - 1; k >= 0; --k) s.A[k] = fsl(k ? s.A[k - 1] : ga, s.A[k], cD);
#pragma unroll
    for (int k = 0; k < NA; ++k) s.B[k] = fsr(s.B[k], k < NA - 1 ? s.B[k + 1] : gb, cR);
    s.pos_y += (int32_t)down;
    const uint32_t moves = down + s.prev_down * (one << 1);      // bit 0 = this round moved down, bit 1 = the one before
    s.prev_down = down;
    // ---- scores: selector byte of a cell = (a ^ b) in the low nibble, 8 | (a ^ b) in the high one (sign replication)
    uint32_t sd[NW];
#pragma unroll
    for (int k = 0; k < NA; ++k) {
        const uint32_t x = s.A[k] ^ s.B[k];
        sd[2 * k] = prmt(s.lut_lo, s.lut_hi, x);
        sd[2 * k + 1] = prmt(s.lut_lo, s.lut_hi, x >> 16);
    }
    // ---- the cells (source.cpp:1916-1926); the tag of the winner is what the traceback will find (source.cpp:1960-1969)
    uint32_t t[NW], t2[NW];
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        t[w] = vmax3(vadd2(D[w], sd[w]), s.H[w], s.V[w]);
        t2[w] = t[w] & SG2_UNTAG;
    }
    // ---- the shuffle stage
    uint32_t m = sg2_max_words<NW>(t2);
    m = vmax2(m, prmt(m, m, 0x1032u));                     // both halves: the maximum of this lane's cells
    const uint32_t xr = prmt(t2[0], s.B[0], 0x0410u), xd = prmt(t2[NW - 1], s.A[NA - 1], 0x0732u);
    uint32_t m1 = m, m2 = m, m3 = m, e0 = t2[0], e31 = t2[NW - 1], gn = xr, gp = xd;      // one lane per pair: nothing to exchange
    if (LAST > 0) {
        m1 = env.shfl_xor(m, 1);
        if (NW == 4) { m2 = env.shfl_xor(m, 2); m3 = env.shfl_xor(m, 3); }
        e0 = env.shfl(t2[0], 0); e31 = env.shfl(t2[NW - 1], LAST);
        gn = env.shfl(xr, q + 1); gp = env.shfl(xd, q - 1);
    }
    // ---- the record (independent of the shuffles).  tag = t - t2 (no borrow between the halves: clearing bits never
    // raises a half), gathered as multiply-adds: word w's tags land at bits 2w (cell 2w) and 16 + 2w (cell 2w + 1)
    if (RECORD) {
        const uint32_t mone = 0u - one;
        uint32_t neg = t2[0] * one, acc = t[0] * one;
#pragma unroll
        for (int w = 1; w < (NW < 8 ? NW : 8); ++w) { neg = t2[w] * (one << (2 * w)) + neg; acc = t[w] * (one << (2 * w)) + acc; }
        if (NW == 16) {                                     // the sums above cover words 0-7; words 8-15 fill a second word
            uint32_t neg2 = t2[8 % NW] * one, acc2 = t[8 % NW] * one;
#pragma unroll
            for (int w = 9; w < NW; ++w) { neg2 = t2[w % NW] * (one << (2 * (w - 8))) + neg2; acc2 = t[w % NW] * (one << (2 * (w - 8))) + acc2; }
            uint32_t* const at = rec_row + rec_stride * round;                             // {tags of cells 0-15, of cells 16-31, moves, 0}
#if defined(__CUDA_ARCH__)
            *reinterpret_cast<uint4*>(at) = make_uint4(neg * mone + acc, neg2 * mone + acc2, moves, 0u);       // one 16-byte store
#else
            at[0] = neg * mone + acc; at[1] = neg2 * mone + acc2; at[2] = moves; at[3] = 0u;
#endif
        } else if (NW == 4) {
            rec_row[rec_stride * round + q] = neg * mone + acc + moves * (one << 8);       // tags in bytes 0 and 2, moves in byte 1
        } else {
            uint32_t* const at = rec_row + rec_stride * round + 2 * q;                     // {32 tag bits, moves}
            at[0] = neg * mone + acc;
            at[1] = moves;
        }
    }
    // ---- round maximum (source.cpp:1925).  In the X-drop frame the best so far sits at 69 or 70, a cell is at most
    // two above it, and the threshold moves exactly when a cell reaches 72 -- so the amount subtracted this round,
    // c = 1 + off, comes from one packed compare; the bookkeeping of the best (below) is off the critical path.
    if (LAST > 0) m = (NW == 4) ? vmax2(vmax3(m, m1, m2), m3) : vmax2(m, m1);
    const uint32_t off = vaddmax2(m, 0xFEE1FEE1u, 0u) & 1u;        // max(m - 287, 0): 1 iff the maximum is 4 * 72
    // ---- X-drop and "<= 0 is dropped" (source.cpp:1918,1933-1936) in the new frame
    // A finished pair must stay finished while its warp runs on: the round after the last one could otherwise revive a
    // cell through its diagonal (two rounds back).  For a dead pair the subtrahend is so large that every cell of this
    // round drops; from then on all its inputs are F.
    const uint32_t nc = 0xFFFCFFFCu - off * 0x00040004u - s.dead * 0x3ED03ED0u;   // (-4c, -4c); dead: a further -16080
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        const uint32_t r = vminu2(vaddmax2(t2[w], nc, SG2_FF), SG2_FF);
        s.R1[w] = vadd2(r, 0x00010001u);
        s.R2[w] = vadd2(r, 0x00020002u);
    }
    // ---- the next round's direction and what enters this lane then
    const int32_t t31 = sg2_half(e31, 1);
    const bool rn = sg2_half(e0, 0) < t31 && t31 > (int32_t)(4u * off);   // t31 - 4c >= 0 (multiples of 4)
    const bool edge = rn ? (q == LAST) : (q == 0);         // the band's end: a dropped cell and a new base come in
    uint32_t cand = rn ? gn : gp;
    if (LAST == 0) {                                        // one lane per pair: always the end; the seq2 side on a right move
        cand = (rn ? s.nextb_raw : s.next_raw) * 0x110000u + (rn ? SG2_F : s.role_base);
    } else if (edge) {
        cand = s.next_raw * 0x110000u + s.role_base;
    }
    s.got = vminu2(vaddmax2(cand, nc & 0xffffu, 0x80000000u | SG2_F), 0xFFFF0000u | SG2_F);   // drop the cell; the upper half (base, stray byte) passes
    s.right = rn ? 1u : 0u;
    // the base after next, branch-free: shift the two staged bases, load the new one under a predicate
    if (LAST == 0) {
        const bool ea = !rn, eb = rn;
        s.cidx += ea ? 1 : 0;
        s.next_raw = ea ? s.next2_raw : s.next_raw;
        s.next2_raw = ld_u8_if(role_seq + s.cidx, ea && (uint32_t)s.cidx < (uint32_t)len, ea ? 4u : s.next2_raw);
        s.cidxb += eb ? 1 : 0;
        s.nextb_raw = eb ? s.nextb2_raw : s.nextb_raw;
        s.nextb2_raw = ld_u8_if(role_seq2 + s.cidxb, eb && (uint32_t)s.cidxb < (uint32_t)len, eb ? 5u : s.nextb2_raw);
    } else {
        s.cidx += edge ? 1 : 0;
        s.next_raw = edge ? s.next2_raw : s.next_raw;
        s.next2_raw = ld_u8_if(role_seq + s.cidx, edge && (uint32_t)s.cidx < (uint32_t)len, edge ? (q == 0 ? 4u : 5u) : s.next2_raw);
    }
    // ---- the best so far (source.cpp:1928-1931: strict, the FIRST round that reaches it)
    const int32_t rmax = sg2_half(m, 0);
    const int32_t amax = (rmax >> 2) - 1 + s.T;            // with the reference's +70 offset
    if (amax > s.best) {
        s.best = amax; s.best_round = round; s.best_py = s.pos_y;
#pragma unroll
        for (int w = 0; w < NW; ++w) s.Rb[w] = t2[w];
    }
    s.T += (int32_t)off;                                   // T = max(best - 70, 1)
    s.dead |= amax > 0 ? 0u : 1u;
    // next round's diagonal inputs are one frame older and carry the tag of the view they come from (ver 2 on a
    // right move, hor 1 on a down move): sd = 4 (score + 1 - off) + 3 - that tag
    const uint32_t dtag = (rn ? 2u : 1u) * 0x01010101u;
    s.lut_lo = 0x0303030Bu - off * 0x03030404u - dtag;     // off = 0: 0B 03 03 03, off = 1: 07 FF FF FF, less the tag
    s.lut_hi = 0x03030303u - off * 0x03030304u - dtag;     //          03 03 03 03,          FF FF FF FF
    return s.dead == 0u;
}

// After the last round: the end cell is the upper-right-most cell of the best round that holds the best score
// (source.cpp:1953-1954); the traceback starts on band element `loc` of `best_round`.
template <int NW, class Env>
SWB_HD void sg2_finish(const Sg2State<NW>& s, Env& env, int32_t& score, int32_t& end_y, int32_t& end_x, int32_t& best_round, int32_t& loc)
{
    constexpr int CELLS = Sg2State<NW>::kCells;
    const int q = env.q();
    uint32_t m = sg2_max_words<NW>(s.Rb);                  // the best round's maximum, once more
    m = vmax2(m, prmt(m, m, 0x1032u));
    if (NW <= 8) m = vmax2(m, env.shfl_xor(m, 1));
    if (NW == 4) m = vmax2(m, env.shfl_xor(m, 2));
    const int32_t best_m = sg2_half(m, 0);
    loc = -1;
#pragma unroll
    for (int c = 0; c < CELLS; ++c)
        if (sg2_half(s.Rb[c >> 1], c & 1) == best_m) loc = CELLS * q + c;
    if (NW <= 8) { const int32_t o = (int32_t)env.shfl_xor((uint32_t)loc, 1); loc = loc > o ? loc : o; }
    if (NW == 4) { const int32_t o = (int32_t)env.shfl_xor((uint32_t)loc, 2); loc = loc > o ? loc : o; }
    score = s.best - SG2_X;
    end_y = s.best_py + 31 - loc;
    end_x = (s.best_round - s.best_py) - 31 + loc;         // pos_x = 31 + (number of right moves)
    best_round = s.best_round;
}

// One traceback step (source.cpp:1958-1971): the walker stands on band element o of round r; rec = that round's record
// (four words).  The tag says where the cell came from (3 diagonal, 2 up, 1 left -- the reference's preference order was
// applied by the forward pass); the move bits say how the band had shifted, which gives the element index of the
// predecessor in ITS round:  up: o + 1 - down(r),  left: o - down(r),  diagonal (two rounds back): o + 1 - down(r) - down(r-1).
// Round 0 is the cell (0,0): the walk ends at r == 0.  Returns the op: 0 = diagonal, 1 = down (y+1), 2 = right (x+1).
template <int NW>
SWB_HD uint32_t sg2_tb_step(uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3, int& o, int& r)
{
    uint32_t code, mv;
    if (NW == 4) {                                          // lane word o >> 3: tags in bytes 0 and 2, moves in byte 1
        const uint32_t w = (o & 16) ? ((o & 8) ? r3 : r2) : ((o & 8) ? r1 : r0);
        code = (w >> (((o & 1) << 4) | (o & 6))) & 3u;
        mv = w >> 8;
    } else if (NW == 8) {                                   // lane o >> 4: {32 tag bits, moves}
        const uint32_t w = (o & 16) ? r2 : r0;
        code = (w >> (((o & 1) << 4) | (o & 14))) & 3u;
        mv = r1;
    } else {                                                // {tags of cells 0-15, tags of cells 16-31, moves, -}
        const uint32_t w = (o & 16) ? r1 : r0;
        code = (w >> (((o & 1) << 4) | (o & 14))) & 3u;
        mv = r2;
    }
    const uint32_t diag = code == 3u ? 1u : 0u;
    o += (int)(code >> 1) - (int)(mv & 1u) - (int)((mv >> 1) & diag);
    r -= 1 + (int)diag;
    return 3u - code;
}

} // namespace swb
