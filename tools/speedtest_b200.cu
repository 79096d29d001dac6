// speedtest_b200: the reference's own metric shape on the drop-in (SpeedTest, /root/reference/source.cpp:3036-3054):
// ONE fixed pair -- the first of the reference's seeded stream -- scored again and again through the per-pair call,
// "ms / 1M calls", from a plain C++ loop (no Python in the timed region).  Beside it, on the same box:
//   * the floor of ANY per-call GPU path here: an empty kernel that only writes a tagged word to mapped pinned memory,
//     launched and spun on the same way (launch + PCIe write + poll), and
//   * the same 10 000 pairs as ONE batch call, for scale.
// Links against libswb200.so only (the product); prints one JSON line.  Built by __graft_entry__.build().
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/speedtest_b200 tools/speedtest_b200.cu \
//        -Lsmith-waterman-simd_b200 -lswb200 -Xlinker -rpath -Xlinker '$ORIGIN/../smith-waterman-simd_b200'
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "../include/swb200.h"

__global__ void floor_kernel(unsigned long long* out, unsigned seq) { out[0] = ((unsigned long long)seq << 32) | 80ull; }

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char** argv)
{
    const int calls = argc > 1 ? atoi(argv[1]) : 20000;
    swb200_ctx* ctx = nullptr;
    if (swb200_init(&ctx, nullptr, 1) != SWB200_OK) { printf("{\"error\": \"%s\"}\n", swb200_last_error(nullptr)); return 1; }
    std::vector<uint8_t> a(128 * 10000), b(128 * 10000);
    swb200_gen_reference_stream(10000, 10000, a.data(), b.data());
    const int8_t sm[16] = {10, -30, -30, -30, -30, 10, -30, -30, -30, -30, 10, -30, -30, -30, -30, 10};   // source.cpp:3041-3045
    int32_t score = -1;
    for (int i = 0; i < 500; ++i) swb200_score_pair(ctx, a.data(), b.data(), sm, 15, &score);
    double t0 = now_s();
    for (int i = 0; i < calls; ++i) swb200_score_pair(ctx, a.data(), b.data(), sm, 15, &score);
    const double per_pair_us = (now_s() - t0) / calls * 1e6;

    // the floor: empty kernel, tagged mapped word, spin
    unsigned long long *h = nullptr, *d = nullptr;
    cudaStream_t st;
    cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    cudaHostAlloc(&h, 8, cudaHostAllocMapped);
    cudaHostGetDevicePointer(&d, h, 0);
    *h = 0;
    volatile unsigned long long* vh = h;
    auto floor_call = [&](unsigned seq) {
        floor_kernel<<<1, 32, 0, st>>>(d, seq);
        while ((unsigned)(*vh >> 32) != seq) __builtin_ia32_pause();
    };
    for (unsigned i = 1; i <= 500; ++i) floor_call(i);
    t0 = now_s();
    for (int i = 0; i < calls; ++i) floor_call(1000u + (unsigned)i);
    const double floor_us = (now_s() - t0) / calls * 1e6;

    // 10 000 distinct pairs: one call per pair, and one batch call
    std::vector<int32_t> s1(10000), s2(10000);
    t0 = now_s();
    for (int i = 0; i < 10000; ++i) swb200_score_pair(ctx, a.data() + 128 * i, b.data() + 128 * i, sm, 15, &s1[i]);
    const double distinct_us = (now_s() - t0) / 10000 * 1e6;
    swb200_score_batch(ctx, a.data(), b.data(), sm, 15, s2.data(), 10000);
    t0 = now_s();
    for (int r = 0; r < 20; ++r) swb200_score_batch(ctx, a.data(), b.data(), sm, 15, s2.data(), 10000);
    const double batch_us = (now_s() - t0) / 20 / 10000 * 1e6;
    const bool same = memcmp(s1.data(), s2.data(), 10000 * sizeof(int32_t)) == 0;

    // the same per-pair call through the chunk pipeline of the throughput kernel (what round 1 did for n = 1: two H2D
    // copies, a launch, a D2H copy, an event wait), for the record
    swb200_set_latency_path(ctx, 0);
    for (int i = 0; i < 200; ++i) swb200_score_pair(ctx, a.data(), b.data(), sm, 15, &score);
    t0 = now_s();
    for (int i = 0; i < 2000; ++i) swb200_score_pair(ctx, a.data(), b.data(), sm, 15, &score);
    const double chunk_us = (now_s() - t0) / 2000 * 1e6;
    swb200_set_latency_path(ctx, 1);

    printf("{\"shape\": \"SpeedTest (source.cpp:3036-3054): one fixed pair, %d calls, C++ loop\", \"us_per_call\": %.3f, \"ms_per_1M_calls\": %.1f, "
           "\"score\": %d, \"score_expected\": 80, \"floor_us_per_call\": %.3f, \"floor\": \"empty kernel + tagged mapped word + spin (launch, PCIe write, poll)\", "
           "\"distinct_pairs_us_per_call\": %.3f, \"batch_of_10000_us_per_pair\": %.4f, \"per_pair_equals_batch\": %s, "
           "\"through_the_throughput_kernel_us_per_call\": %.3f}\n",
           calls, per_pair_us, per_pair_us * 1e3, score, floor_us, distinct_us, batch_us, same ? "true" : "false", chunk_us);
    cudaFreeHost(h);
    cudaStreamDestroy(st);
    swb200_shutdown(ctx);
    return (score == 80 && same) ? 0 : 2;
}
