// tests/emu/sg2_emu.cpp -- TEST TOOL.  Runs the semi-global aligner's per-lane code (csrc/sg2_core.cuh) on the
// HOST: the lanes of a pair (four, or two with sixteen cells per lane) are coroutines that advance in lock step, a
// shuffle is "publish, yield, read the other lane's slot".  This validates the frame arithmetic, the sentinel, the
// funnel-shift band moves, the PRMT tables and the record layout against the oracle in a container without a GPU.
// Never shipped, never loaded by the product; the product's only compute path is the CUDA kernel.
#include <ucontext.h>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "sg2_core.cuh"

using namespace swb;

namespace {

struct Quad {
    ucontext_t main_ctx, lane_ctx[4];
    uint32_t slot[2][4];
    int phase[4] = {0, 0, 0, 0};
    bool done[4] = {false, false, false, false};
    int lanes;
    // job
    const uint8_t* seq1; const uint8_t* seq2; int len;
    std::vector<uint32_t> rec;      // [rounds][4]
    int32_t score, end_y, end_x, best_round, loc;
};

struct HostEnv {
    Quad* quad; int lane;
    int q() const { return lane; }
    uint32_t one() const { return 1u; }
    uint32_t shfl(uint32_t v, int src)
    {
        const int ph = quad->phase[lane]++ & 1;
        quad->slot[ph][lane] = v;
        swapcontext(&quad->lane_ctx[lane], &quad->main_ctx);
        return quad->slot[ph][src & (quad->lanes - 1)];
    }
    uint32_t shfl_xor(uint32_t v, int m) { return shfl(v, lane ^ m); }
};

Quad* g_quad;

template <int NW>
void lane_main(int lane)
{
    Quad& Q = *g_quad;
    HostEnv env{&Q, lane};
    Sg2State<NW> s;
    sg2_init(s, env, Q.seq1, Q.seq2, Q.len);
    const uint8_t* role = lane == 0 ? Q.seq1 : Q.seq2;
    const int max_round = 2 * Q.len + 1;
    // On the device a finished pair keeps running rounds beside the live pairs of its warp: run EVERY round here, so
    // that a finished pair that is not inert (a best, a threshold or a record that still moves) shows up as a mismatch.
    if (Sg2State<NW>::kLanes == 1) role = Q.seq1;
    for (int round = 1; round < max_round; ++round) sg2_round<true>(s, env, role, Q.len, round, Q.rec.data(), 4, Q.seq2);
    int32_t sc, ey, ex, br, loc;
    sg2_finish(s, env, sc, ey, ex, br, loc);
    if (lane == 0) { Q.score = sc; Q.end_y = ey; Q.end_x = ex; Q.best_round = br; Q.loc = loc; }
    Q.done[lane] = true;
    swapcontext(&Q.lane_ctx[lane], &Q.main_ctx);
}

template <int NW>
int run(const uint8_t* seq1, const uint8_t* seq2, int len, int32_t* score, int32_t* end_y, int32_t* end_x, uint8_t* ops, int32_t* n_ops)
{
    constexpr int LANES = Sg2State<NW>::kLanes;
    Quad Q;
    Q.lanes = LANES;
    Q.seq1 = seq1; Q.seq2 = seq2; Q.len = len;
    Q.rec.assign((size_t)4 * (2 * (size_t)len + 64), 0xdeadbeefu);
    g_quad = &Q;
    const size_t stack_bytes = 256 * 1024;
    std::vector<char> stacks(LANES * stack_bytes);
    for (int l = 0; l < LANES; ++l) {
        getcontext(&Q.lane_ctx[l]);
        Q.lane_ctx[l].uc_stack.ss_sp = stacks.data() + l * stack_bytes;
        Q.lane_ctx[l].uc_stack.ss_size = stack_bytes;
        Q.lane_ctx[l].uc_link = &Q.main_ctx;
        makecontext(&Q.lane_ctx[l], (void (*)())lane_main<NW>, 1, l);
    }
    for (;;) {
        bool any = false;
        for (int l = 0; l < LANES; ++l)
            if (!Q.done[l]) { any = true; swapcontext(&Q.main_ctx, &Q.lane_ctx[l]); }
        if (!any) break;
    }
    *score = Q.score; *end_y = Q.end_y; *end_x = Q.end_x;
    // traceback over the records, as the device's traceback kernel does
    const uint32_t* rec = Q.rec.data();
    int r = Q.best_round, o = Q.loc;
    std::vector<uint8_t> rev;
    while (r > 0) {
        if ((int)rev.size() >= 2 * len) return -2;
        if (o < 0 || o > 31) return -3;
        const uint32_t* w = rec + 4 * (size_t)r;
        rev.push_back((uint8_t)sg2_tb_step<NW>(w[0], w[1], w[2], w[3], o, r));
    }
    if (r != 0 || o != 31) return -5;       // (0,0) is element 31 of round 0
    *n_ops = (int32_t)rev.size();
    for (size_t k = 0; k < rev.size(); ++k) ops[k] = rev[rev.size() - 1 - k];
    return 0;
}

} // namespace

// words_per_lane: 4 (four lanes per pair), 8 (two lanes per pair) or 16 (one lane per pair)
extern "C" int swemu_sg2(const uint8_t* seq1, const uint8_t* seq2, int len, int32_t* score, int32_t* end_y, int32_t* end_x,
                         uint8_t* ops, int32_t* n_ops, int words_per_lane)
{
    if (words_per_lane == 16) return run<16>(seq1, seq2, len, score, end_y, end_x, ops, n_ops);
    if (words_per_lane == 8) return run<8>(seq1, seq2, len, score, end_y, end_x, ops, n_ops);
    return run<4>(seq1, seq2, len, score, end_y, end_x, ops, n_ops);
}
