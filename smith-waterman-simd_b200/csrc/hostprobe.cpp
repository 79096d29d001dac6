// hostprobe.cpp -- swb200_host_read_bandwidth: how fast the host's cores can stream-read a buffer (bench.py's
// `host_ceiling` leg).  A byte-coded host batch has to be read from host DRAM exactly once -- by a core that packs it
// to 2 bits or by the GPU's DMA engine -- so this rate, together with the measured H2D rate, is the roofline of the
// end-to-end number for byte-coded input.  Benchmark helper only: nothing on the scoring path calls it.
#include "../../include/swb200.h"

#include <chrono>
#include <cstring>
#include <exception>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace {

#if defined(__x86_64__)
__attribute__((target("avx2")))
uint64_t read_avx2(const uint8_t* p, size_t n)
{
    __m256i a0 = _mm256_setzero_si256(), a1 = a0, a2 = a0, a3 = a0;
    size_t i = 0;
    for (; i + 128 <= n; i += 128) {
        _mm_prefetch((const char*)(p + i + 8192), _MM_HINT_T1);       // as the packer does (hostpack.cpp)
        _mm_prefetch((const char*)(p + i + 8256), _MM_HINT_T1);
        a0 = _mm256_or_si256(a0, _mm256_loadu_si256((const __m256i*)(p + i)));
        a1 = _mm256_or_si256(a1, _mm256_loadu_si256((const __m256i*)(p + i + 32)));
        a2 = _mm256_or_si256(a2, _mm256_loadu_si256((const __m256i*)(p + i + 64)));
        a3 = _mm256_or_si256(a3, _mm256_loadu_si256((const __m256i*)(p + i + 96)));
    }
    a0 = _mm256_or_si256(_mm256_or_si256(a0, a1), _mm256_or_si256(a2, a3));
    uint64_t w[4];
    _mm256_storeu_si256((__m256i*)w, a0);
    uint64_t r = w[0] | w[1] | w[2] | w[3];
    for (; i < n; ++i) r |= p[i];
    return r;
}
#endif

uint64_t read_plain(const uint8_t* p, size_t n)
{
    uint64_t r = 0;
    size_t i = 0;
    for (; i + 8 <= n; i += 8) { uint64_t x; memcpy(&x, p + i, 8); r |= x; }
    for (; i < n; ++i) r |= p[i];
    return r;
}

} // namespace

extern "C" int swb200_host_read_bandwidth(const void* buf, uint64_t bytes, int threads, int passes, double* bytes_per_s)
{
    if (!buf || !bytes_per_s || bytes == 0 || threads < 1 || passes < 1) return SWB200_ERR_ARG;
#if defined(__x86_64__)
    static const bool have_avx2 = __builtin_cpu_supports("avx2");
#else
    const bool have_avx2 = false;
#endif
    const uint8_t* base = static_cast<const uint8_t*>(buf);
    std::vector<std::thread> pool;
    std::vector<uint64_t> sink;
    const auto t0 = std::chrono::steady_clock::now();
    try {
        sink.assign((size_t)threads, 0);
        pool.reserve((size_t)threads);
        for (int k = 0; k < threads; ++k) {
            const uint64_t lo = bytes * (uint64_t)k / (uint64_t)threads, hi = bytes * (uint64_t)(k + 1) / (uint64_t)threads;
            pool.emplace_back([=, &sink] {
                uint64_t r = 0;
                for (int p = 0; p < passes; ++p) {
#if defined(__x86_64__)
                    r |= have_avx2 ? read_avx2(base + lo, hi - lo) : read_plain(base + lo, hi - lo);
#else
                    r |= read_plain(base + lo, hi - lo);
#endif
                }
                sink[(size_t)k] = r;
            });
        }
    } catch (const std::exception&) {
        for (auto& t : pool) t.join();
        return SWB200_ERR_NOMEM;
    }
    for (auto& t : pool) t.join();
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    volatile uint64_t keep = 0;
    for (uint64_t v : sink) keep = keep | v;
    (void)keep;
    *bytes_per_s = dt > 0 ? (double)bytes * passes / dt : 0.0;
    return SWB200_OK;
}
