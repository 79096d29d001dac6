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

// PCIe read round trip as a kernel sees it: a chain of dependent loads of one word of mapped pinned host memory
__global__ void rtt_kernel(const unsigned* host_word, unsigned long long* out_ns, int n)
{
    unsigned long long t0, t1;
    unsigned off = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (int i = 0; i < n; ++i) {
        unsigned v;
        asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(host_word + off) : "memory");
        off = v;                                   // (the host word holds 0: a true dependence the compiler cannot see through)
    }
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    out_ns[0] = (t1 - t0) / (unsigned long long)n + off;
}

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
    uint64_t server_launches = 0, doorbell_calls = 0;
    uint32_t sweep_ns = 0;
    swb200_pair_path_stats(ctx, &server_launches, &doorbell_calls, &sweep_ns);

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

    // PCIe read round trip seen by a kernel (what a poll of the doorbell costs)
    unsigned long long *h_rtt = nullptr, *d_rtt = nullptr;
    cudaHostAlloc(&h_rtt, 64, cudaHostAllocMapped);
    cudaHostGetDevicePointer(&d_rtt, h_rtt, 0);
    h_rtt[0] = 0; h_rtt[1] = 0;
    rtt_kernel<<<1, 1, 0, st>>>(reinterpret_cast<const unsigned*>(d_rtt + 1), d_rtt, 2000);
    cudaStreamSynchronize(st);
    const double rtt_us = (double)h_rtt[0] * 1e-3;
    cudaFreeHost(h_rtt);

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

    // the same call with one launch of the latency kernel per call (no resident server): what the doorbell saves
    swb200_set_latency_path(ctx, 2);
    for (int i = 0; i < 500; ++i) swb200_score_pair(ctx, a.data(), b.data(), sm, 15, &score);
    t0 = now_s();
    for (int i = 0; i < calls; ++i) swb200_score_pair(ctx, a.data(), b.data(), sm, 15, &score);
    const double launch_us = (now_s() - t0) / calls * 1e6;
    swb200_set_latency_path(ctx, 1);

    // the same per-pair call through the chunk pipeline of the throughput kernel (what round 1 did for n = 1: two H2D
    // copies, a launch, a D2H copy, an event wait), for the record
    swb200_set_latency_path(ctx, 0);
    for (int i = 0; i < 200; ++i) swb200_score_pair(ctx, a.data(), b.data(), sm, 15, &score);
    t0 = now_s();
    for (int i = 0; i < 2000; ++i) swb200_score_pair(ctx, a.data(), b.data(), sm, 15, &score);
    const double chunk_us = (now_s() - t0) / 2000 * 1e6;
    swb200_set_latency_path(ctx, 1);

    printf("{\"shape\": \"SpeedTest (source.cpp:3036-3054): one fixed pair, %d calls, C++ loop\", \"us_per_call\": %.3f, \"ms_per_1M_calls\": %.1f, "
           "\"score\": %d, \"score_expected\": 80, \"path\": \"doorbell of the resident one-warp server kernel (no launch per call)\", "
           "\"server_kernels_launched\": %llu, \"doorbell_calls\": %llu, \"sweep_us_measured_by_the_server\": %.3f, "
           "\"one_launch_per_call_us\": %.3f, \"pcie_read_round_trip_us\": %.3f, \"floor_us_per_call\": %.3f, \"floor\": \"of any launch-per-call path: empty kernel + tagged mapped word + spin (launch, PCIe write, poll)\", "
           "\"distinct_pairs_us_per_call\": %.3f, \"batch_of_10000_us_per_pair\": %.4f, \"per_pair_equals_batch\": %s, "
           "\"through_the_throughput_kernel_us_per_call\": %.3f}\n",
           calls, per_pair_us, per_pair_us * 1e3, score, (unsigned long long)server_launches, (unsigned long long)doorbell_calls, sweep_ns * 1e-3, launch_us, rtt_us, floor_us, distinct_us, batch_us, same ? "true" : "false", chunk_us);
    cudaFreeHost(h);
    cudaStreamDestroy(st);
    swb200_shutdown(ctx);
    return (score == 80 && same) ? 0 : 2;
}
