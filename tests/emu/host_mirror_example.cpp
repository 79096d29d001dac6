// Compiles the C++ host mirror (smith_waterman_b200.hpp) the way the reference's own test
// would use it (cf. TestSimdSmithWaterman, source.cpp:2943-2982): same generator, same
// matrix, same call shape, the new function next to a scalar check.  On a machine without a
// B200 the first call throws (no CPU fallback); on the GPU box it prints the first scores.
#include <cstdio>
#include <random>
#include "../../smith-waterman-simd_b200/host/smith_waterman_b200.hpp"

int main()
{
    std::mt19937_64 rnd(10000);
    std::vector<std::array<uint8_t, 128>> as(16), bs(16);
    for (int it = 0; it < 16; ++it)
        for (int i = 0; i < 128; ++i) {
            as[it][i] = (uint8_t)(rnd() >> 62);
            bs[it][i] = (uint8_t)(rnd() >> 62);
        }
    const std::array<int8_t, 16> score_matrix = {10, -30, -30, -30, -30, 10, -30, -30, -30, -30, 10, -30, -30, -30, -30, 10};
    const int8_t gap_penalty = 15;
    try {
        std::vector<int> dest;
        SmithWaterman_b200_batch(as, bs, score_matrix, gap_penalty, dest);
        const int expect[16] = {80, 80, 70, 95, 70, 80, 80, 75, 80, 70, 75, 65, 65, 80, 100, 70};
        int bad = 0;
        for (int it = 0; it < 16; ++it) {
            const int one = SmithWaterman_b200(as[it], bs[it], score_matrix, gap_penalty);
            bad += (one != expect[it]) + (dest[it] != expect[it]);
        }
        // the x32 shape (source.cpp:1227-1234): 32 queries vs one target, fixed +1/-1/1
        std::array<uint8_t, 128 * 32> q32;
        for (int p = 0; p < 32; ++p) for (int i = 0; i < 128; ++i) q32[p * 128 + i] = as[p % 16][i];
        std::array<int, 32> dest32;
        const std::array<int8_t, 16> m111 = {1, -1, -1, -1, -1, 1, -1, -1, -1, -1, 1, -1, -1, -1, -1, 1};
        const int first = SmithWaterman_b200_x32(q32, bs[0], dest32);
        for (int p = 0; p < 32; ++p) bad += (dest32[p] != SmithWaterman_b200(as[p % 16], bs[0], m111, 1));
        bad += (first != dest32[0]);
        // the fixed 1/1/1 call (source.cpp:1073-1076) against the general one with the same scoring
        for (int it = 0; it < 16; ++it) bad += (SmithWaterman_111_b200(as[it], bs[it]) != SmithWaterman_b200(as[it], bs[it], m111, 1));
        const int expect111[4] = {18, 20, 14, 24};   // tests/golden/golden.json, x32_1_-1_1 first16
        for (int it = 0; it < 4; ++it) bad += (SmithWaterman_111_b200(as[it], bs[it]) != expect111[it]);
        std::printf("host mirror: %s\n", bad ? "MISMATCH" : "16/16 scores equal the reference's known answers");
        return bad ? 1 : 0;
    } catch (const std::exception& e) {
        std::printf("host mirror: %s\n", e.what());
        return 2;
    }
}
