/*
 * sg_oracle.c -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.
 *
 * A plain-C restatement of the reference's scalar adaptive-banded X-drop semi-global aligner,
 * SemiGlobal_AdaptiveBanded_XDrop_111_32_70 (/root/reference/source.cpp:1836-1976): match /
 * mismatch / gap = 1/1/1, band of 32 cells on an anti-diagonal, X-drop threshold 70, alignment
 * anchored at (0,0), free end at the best cell, traceback with the preference diagonal > up > left.
 * The reference fixes the sequence length at 16384; here it is a parameter so that tests can also
 * run short cases (the arithmetic is the same at every length).
 *
 * Parity status: PINNED -- tests/test_semiglobal_oracle.py checks this file against the reference
 * itself (scalar and its four AVX2 forms, source.cpp:1978-2725, compiled unmodified into
 * oracle/_ref/) on the inputs of TestSemiGlobal (source.cpp:2750-2771) and SpeedtestSemiGlobal
 * (source.cpp:2804-2813), and against the committed fixtures tests/golden/semiglobal.npz.
 *
 * Band geometry (source.cpp:1872-1911): round r holds the 32 cells of one anti-diagonal; element
 * i (31 = upper-right end, 0 = lower-left end) is the cell (y, x) = (pos_y + 31 - i, pos_x - 31 + i)
 * with pos_x counted in the padded target (32 pad characters in front).  Stored values carry an
 * offset of +70 (the X-drop threshold), and 0 means "dropped / never reached".  Each round the band
 * moves right if result[0] < result[31], else down (source.cpp:1883-1906).
 *
 * One deliberate deviation: at pos_y = len+1 the reference reads seq1p one byte past its end
 * (source.cpp:1913 with i = 0; the array has 1+16384+31 bytes).  Every cell of such a round lies
 * below the last row, so neither the score nor the traceback can depend on that byte; this
 * restatement treats it as padding.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SG_BAND 32
#define SG_X 70

/* ops[k] in forward order from (0,0): 0 = diagonal (y+1,x+1), 1 = down (y+1), 2 = right (x+1).
 * Returns 0, or -1 on allocation failure / bad arguments.  ops must hold 2*len entries. */
int swo_semiglobal_xdrop(const uint8_t* seq1, const uint8_t* seq2, int len,
                         int32_t* score, int32_t* end_y, int32_t* end_x, uint8_t* ops, int32_t* n_ops)
{
    if (len < 1 || !seq1 || !seq2) return -1;
    const int max_round = (len + 1) * 2 - 1;
    uint8_t* seq1p = (uint8_t*)malloc((size_t)1 + len + 31 + 1);      /* +1: the byte the reference over-reads */
    uint8_t* seq2p = (uint8_t*)malloc((size_t)32 + len + 31 + 1);
    int32_t* dp = (int32_t*)calloc((size_t)SG_BAND * max_round, sizeof(int32_t));
    int32_t* pos_y = (int32_t*)calloc((size_t)max_round, sizeof(int32_t));
    int32_t* pos_x = (int32_t*)calloc((size_t)max_round, sizeof(int32_t));
    if (!seq1p || !seq2p || !dp || !pos_y || !pos_x) { free(seq1p); free(seq2p); free(dp); free(pos_y); free(pos_x); return -1; }
    memset(seq1p, 0xF0, (size_t)1 + len + 31 + 1);                    /* source.cpp:1859-1863 */
    memcpy(seq1p + 1, seq1, (size_t)len);
    memset(seq2p, 0xF0, (size_t)32 + len + 31 + 1);                   /* source.cpp:1866-1870 */
    memcpy(seq2p + 32, seq2, (size_t)len);

    int32_t horizontal[SG_BAND] = {0}, vertical[SG_BAND] = {0}, diagonal[SG_BAND] = {0}, result[SG_BAND] = {0};
    dp[31] = SG_X;                                                    /* source.cpp:1877 */
    pos_y[0] = 0; pos_x[0] = 31;
    result[31] = SG_X;
    int now_y = 0, now_x = 31, best_round = 0, best = SG_X;
    for (int round = 1; round < max_round; ++round) {
        if (result[0] < result[31]) {                                 /* right, source.cpp:1891-1901 */
            for (int i = 0; i < SG_BAND; ++i) diagonal[i] = vertical[i];
            for (int i = 0; i < SG_BAND; ++i) horizontal[i] = result[i];
            for (int i = 0; i < SG_BAND - 1; ++i) vertical[i] = result[i + 1];
            vertical[SG_BAND - 1] = 0;
            if (32 + len + 31 < ++now_x) break;
        } else {                                                      /* down, source.cpp:1902-1911 */
            for (int i = 0; i < SG_BAND; ++i) diagonal[i] = horizontal[i];
            for (int i = 0; i < SG_BAND; ++i) vertical[i] = result[i];
            for (int i = SG_BAND - 1; i >= 1; --i) horizontal[i] = result[i - 1];
            horizontal[0] = 0;
            if (1 + len < ++now_y) break;
        }
        pos_y[round] = now_y; pos_x[round] = now_x;
        int round_best = 0;
        for (int i = 0; i < SG_BAND; ++i) {                           /* source.cpp:1916-1926 */
            const uint8_t a = seq1p[now_y + (31 - i)], b = seq2p[now_x - (31 - i)];
            const int s = (a < 4 && b < 4 && a == b) ? 1 : -1;
            int v = 0;
            if (diagonal[i] != 0 && diagonal[i] + s > v) v = diagonal[i] + s;
            if (horizontal[i] != 0 && horizontal[i] - 1 > v) v = horizontal[i] - 1;
            if (vertical[i] != 0 && vertical[i] - 1 > v) v = vertical[i] - 1;
            result[i] = v;
            if (round_best < v) round_best = v;
        }
        if (best < round_best) { best_round = round; best = round_best; }   /* source.cpp:1928-1931 */
        for (int i = 0; i < SG_BAND; ++i) {                           /* X-drop, source.cpp:1933-1936 */
            if (result[i] < best - SG_X) result[i] = 0;
            dp[(size_t)round * SG_BAND + i] = result[i];
        }
        if (round_best == 0) break;                                   /* source.cpp:1938-1941 */
    }

    /* Get(y,x), source.cpp:1944-1951; 0 stands for the reference's minus_inf */
#define SG_GET(y, x, out) do {                                                         \
        int32_t g_ = 0;                                                                \
        if ((y) >= 0 && (y) <= len && (x) >= 0 && (x) <= len) {                        \
            const int r_ = (y) + (x);                                                  \
            const int o_ = 31 - ((y) - pos_y[r_]);                                     \
            if (o_ >= 0 && o_ < SG_BAND) g_ = dp[(size_t)r_ * SG_BAND + o_];           \
        }                                                                              \
        (out) = g_;                                                                    \
    } while (0)

    int by = pos_y[best_round], bx = pos_x[best_round] - 31;          /* source.cpp:1953-1954 */
    for (;;) { int32_t g; SG_GET(by, bx, g); if (g == best) break; ++by; --bx; }
    *score = best - SG_X;
    *end_y = by; *end_x = bx;
    int n = 0, rc = 0;
    for (int i = by, j = bx; i || j;) {                               /* source.cpp:1958-1971 */
        int32_t cur, d, u, l;
        SG_GET(i, j, cur);
        SG_GET(i - 1, j - 1, d); SG_GET(i - 1, j, u); SG_GET(i, j - 1, l);
        const int s = (i && j && seq1[i - 1] == seq2[j - 1]) ? 1 : -1;
        if (n >= 2 * len) { rc = -2; break; }
        if (i && j && d != 0 && cur == d + s) { ops[n++] = 0; --i; --j; }
        else if (i && u != 0 && cur == u - 1) { ops[n++] = 1; --i; }
        else if (j && l != 0 && cur == l - 1) { ops[n++] = 2; --j; }
        else { rc = -3; break; }                                      /* the reference's assert(0) */
    }
    for (int k = 0; k < n / 2; ++k) { const uint8_t t = ops[k]; ops[k] = ops[n - 1 - k]; ops[n - 1 - k] = t; }
    *n_ops = n;
    free(seq1p); free(seq2p); free(dp); free(pos_y); free(pos_x);
    return rc;
}
