// tests/emu/emu_main.cpp -- TEST TOOL.  Compiles the kernel's per-thread algorithm
// (csrc/sw_core.cuh) for the HOST with emulated packed instructions and runs it over a
// batch, so that the algorithm (offset frames, wrap, FIFO re-basing, selectors) can be
// validated against the oracle in a container without a GPU.  Never shipped, never
// loaded by the product; the product's only compute path is the CUDA kernel.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "sw_params.h"

using namespace swb;

static int g_prefetch = 0;  // 1: exercise the read-ahead path used by the global-memory FIFO

template <int L>
struct HostFifo {
    static constexpr int kPrefetch = 0;
    uint32_t w[L];
    uint32_t pop(int c) const { return w[c & (L - 1)]; }
    void push(int c, uint32_t v) { w[c & (L - 1)] = v; }
};
template <int L, int AHEAD>
struct HostFifoAhead {          // same storage, but sw_core.cuh reads it ahead of use (Fifo::kPrefetch = 1 or 2)
    static constexpr int kPrefetch = AHEAD;
    uint32_t w[L];
    uint32_t pop(int c) const { return w[c & (L - 1)]; }
    void push(int c, uint32_t v) { w[c & (L - 1)] = v; }
};
struct HostTable {
    const uint32_t* t;
    uint32_t operator()(uint32_t byte_off) const { return t[byte_off >> 2]; }
};

static int g_variant = 0;   // variant bits of sw_core.cuh under test (SW_V_BEST_FMA)

template <int L, int V>
static void two_pairs(bool fast, const uint8_t* a, const uint8_t* b, uint32_t dqa, uint32_t dqb, HostFifo<L>& fifo, HostTable& t4,
                      const SwParams& prm, int32_t& lo, int32_t& hi)
{
    if (g_prefetch == 1) {
        HostFifoAhead<L, 1> ahead;
        for (int i = 0; i < L; ++i) ahead.w[i] = 0xdeadbeefu;   // anything read before it was written shows up as a wrong score
        if (fast) sw_two_pairs<true, L, V>(a, b, dqa, dqb, ahead, t4, prm, lo, hi);
        else      sw_two_pairs<false, L, V>(a, b, dqa, dqb, ahead, t4, prm, lo, hi);
        return;
    }
    if (g_prefetch == 2) {
        HostFifoAhead<L, 2> ahead;
        for (int i = 0; i < L; ++i) ahead.w[i] = 0xdeadbeefu;
        if (fast) sw_two_pairs<true, L, V>(a, b, dqa, dqb, ahead, t4, prm, lo, hi);
        else      sw_two_pairs<false, L, V>(a, b, dqa, dqb, ahead, t4, prm, lo, hi);
        return;
    }
    if (fast) sw_two_pairs<true, L, V>(a, b, dqa, dqb, fifo, t4, prm, lo, hi);
    else      sw_two_pairs<false, L, V>(a, b, dqa, dqb, fifo, t4, prm, lo, hi);
}

template <int L>
static int run_len(const uint8_t* seq1, const uint8_t* seq2, const int8_t* sm, int gap, int32_t* scores, uint64_t n, int force_general,
                   uint32_t stride2 = (uint32_t)L)
{
    if (!sw_len_supported(sm, L)) return -2;
    const SwParams prm = sw_make_params(sm, gap, force_general, L);
    HostTable t4{prm.t4};
    for (uint64_t p = 0; p < n; p += 2) {
        const uint64_t q = (p + 1 < n) ? p + 1 : p;
        HostFifo<L> fifo;
        int32_t lo, hi;
        const uint32_t dq = (q != p) ? (uint32_t)L : 0u;
        const uint32_t dqb = (q != p) ? stride2 : 0u;
        switch (g_variant) {
        case 0: two_pairs<L, 0>(prm.fast, seq1 + p * L, seq2 + p * stride2, dq, dqb, fifo, t4, prm, lo, hi); break;
        case 1: two_pairs<L, 1>(prm.fast, seq1 + p * L, seq2 + p * stride2, dq, dqb, fifo, t4, prm, lo, hi); break;
        case 2: two_pairs<L, 2>(prm.fast, seq1 + p * L, seq2 + p * stride2, dq, dqb, fifo, t4, prm, lo, hi); break;   // SW_V_FIFO_PREOFF
        case 3: two_pairs<L, 3>(prm.fast, seq1 + p * L, seq2 + p * stride2, dq, dqb, fifo, t4, prm, lo, hi); break;
        default: two_pairs<L, 1>(prm.fast, seq1 + p * L, seq2 + p * stride2, dq, dqb, fifo, t4, prm, lo, hi); break;
        }
        scores[p] = lo;
        if (q != p) scores[q] = hi;
    }
    return prm.fast;
}

// returns 1 = fast kernel's algorithm ran, 0 = general, -1 = outside the reference domain, -2 = length unsupported
extern "C" int swemu_score_batch_len(int len, const uint8_t* seq1, const uint8_t* seq2, const int8_t* sm, int gap,
                                     int32_t* scores, uint64_t n, int force_general)
{
    if (sw_check_domain(sm, gap) != SW_DOMAIN_OK) return -1;
    switch (len) {
    case 128: return run_len<128>(seq1, seq2, sm, gap, scores, n, force_general);
    case 256: return run_len<256>(seq1, seq2, sm, gap, scores, n, force_general);
    case 512: return run_len<512>(seq1, seq2, sm, gap, scores, n, force_general);
    default: return -2;
    }
}

extern "C" int swemu_score_batch(const uint8_t* seq1, const uint8_t* seq2, const int8_t* sm, int gap,
                                 int32_t* scores, uint64_t n, int force_general)
{
    return swemu_score_batch_len(128, seq1, seq2, sm, gap, scores, n, force_general);
}

// many queries against one target (stride 0 on the target side), as the kernel does for swb200_score_one_vs_many
extern "C" int swemu_one_vs_many(const uint8_t* seq1s, const uint8_t* seq2, const int8_t* sm, int gap,
                                 int32_t* scores, uint64_t n, int force_general)
{
    if (sw_check_domain(sm, gap) != SW_DOMAIN_OK) return -1;
    return run_len<128>(seq1s, seq2, sm, gap, scores, n, force_general, 0u);
}

extern "C" void swemu_set_variant(int v) { g_variant = v; }
extern "C" void swemu_set_prefetch(int on) { g_prefetch = on; }
