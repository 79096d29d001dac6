// packbench: variants of the host 2-bit packing loop (csrc/hostpack.cpp) on the box it runs on -- development tool, not
// part of the product.  Each thread packs its own contiguous share of a large byte-coded buffer (first-touched by that
// thread); prints GB/s of INPUT per variant and thread count.   g++ -O3 -pthread -o tools/packbench tools/packbench.cpp
#include <immintrin.h>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

template <int PF, int HINT, bool STREAM>
__attribute__((target("avx2"))) static void pack_avx2(const uint8_t* codes, uint8_t* packed, size_t n)
{
    const __m256i m3 = _mm256_set1_epi8(3), w14 = _mm256_set1_epi16(0x0401), w116 = _mm256_set1_epi32(0x00100001);
    const __m256i order = _mm256_setr_epi32(0, 4, 1, 5, 2, 6, 3, 7);
    for (size_t i = 0; i + 128 <= n; i += 128) {
        if (PF > 0) {
            _mm_prefetch((const char*)(codes + i + PF), (_mm_hint)HINT);
            _mm_prefetch((const char*)(codes + i + PF + 64), (_mm_hint)HINT);
        }
        __m256i a = _mm256_and_si256(_mm256_loadu_si256((const __m256i*)(codes + i)), m3);
        __m256i b = _mm256_and_si256(_mm256_loadu_si256((const __m256i*)(codes + i + 32)), m3);
        __m256i c = _mm256_and_si256(_mm256_loadu_si256((const __m256i*)(codes + i + 64)), m3);
        __m256i d = _mm256_and_si256(_mm256_loadu_si256((const __m256i*)(codes + i + 96)), m3);
        a = _mm256_madd_epi16(_mm256_maddubs_epi16(a, w14), w116);
        b = _mm256_madd_epi16(_mm256_maddubs_epi16(b, w14), w116);
        c = _mm256_madd_epi16(_mm256_maddubs_epi16(c, w14), w116);
        d = _mm256_madd_epi16(_mm256_maddubs_epi16(d, w14), w116);
        __m256i r = _mm256_packus_epi16(_mm256_packus_epi32(a, b), _mm256_packus_epi32(c, d));
        r = _mm256_permutevar8x32_epi32(r, order);
        if (STREAM) _mm256_stream_si256((__m256i*)(packed + i / 4), r);
        else _mm256_storeu_si256((__m256i*)(packed + i / 4), r);
    }
    if (STREAM) _mm_sfence();
}

template <int PF, bool STREAM>
__attribute__((target("avx512f,avx512bw"))) static void pack_avx512(const uint8_t* codes, uint8_t* packed, size_t n)
{
    const __m512i m3 = _mm512_set1_epi8(3), w14 = _mm512_set1_epi16(0x0401), w116 = _mm512_set1_epi32(0x00100001);
    for (size_t i = 0; i + 256 <= n; i += 256) {
        if (PF > 0) {
            _mm_prefetch((const char*)(codes + i + PF), _MM_HINT_T0);
            _mm_prefetch((const char*)(codes + i + PF + 64), _MM_HINT_T0);
            _mm_prefetch((const char*)(codes + i + PF + 128), _MM_HINT_T0);
            _mm_prefetch((const char*)(codes + i + PF + 192), _MM_HINT_T0);
        }
        __m512i a = _mm512_and_si512(_mm512_loadu_si512(codes + i), m3);
        __m512i b = _mm512_and_si512(_mm512_loadu_si512(codes + i + 64), m3);
        __m512i c = _mm512_and_si512(_mm512_loadu_si512(codes + i + 128), m3);
        __m512i d = _mm512_and_si512(_mm512_loadu_si512(codes + i + 192), m3);
        a = _mm512_madd_epi16(_mm512_maddubs_epi16(a, w14), w116);     // 16 dwords, each one packed byte
        b = _mm512_madd_epi16(_mm512_maddubs_epi16(b, w14), w116);
        c = _mm512_madd_epi16(_mm512_maddubs_epi16(c, w14), w116);
        d = _mm512_madd_epi16(_mm512_maddubs_epi16(d, w14), w116);
        __m512i r = _mm512_castsi128_si512(_mm512_cvtepi32_epi8(a));
        r = _mm512_inserti32x4(r, _mm512_cvtepi32_epi8(b), 1);
        r = _mm512_inserti32x4(r, _mm512_cvtepi32_epi8(c), 2);
        r = _mm512_inserti32x4(r, _mm512_cvtepi32_epi8(d), 3);
        if (STREAM) _mm512_stream_si512((__m512i*)(packed + i / 4), r);
        else _mm512_storeu_si512(packed + i / 4, r);
    }
    if (STREAM) _mm_sfence();
}

typedef void (*PackFn)(const uint8_t*, uint8_t*, size_t);
struct Variant { const char* name; PackFn fn; };

int main(int argc, char** argv)
{
    const size_t per_thread = (argc > 1 ? atol(argv[1]) : 64) << 20;      // MiB of input per thread
    const Variant vs[] = {
        {"avx2 pf2K t0 nt (shipped)", pack_avx2<2048, _MM_HINT_T0, true>},
        {"avx2 no-pf nt", pack_avx2<0, _MM_HINT_T0, true>},
        {"avx2 pf1K t0 nt", pack_avx2<1024, _MM_HINT_T0, true>},
        {"avx2 pf4K t0 nt", pack_avx2<4096, _MM_HINT_T0, true>},
        {"avx2 pf8K t0 nt", pack_avx2<8192, _MM_HINT_T0, true>},
        {"avx2 pf4K nta nt", pack_avx2<4096, _MM_HINT_NTA, true>},
        {"avx2 pf4K t1 nt", pack_avx2<4096, _MM_HINT_T1, true>},
        {"avx2 pf8K t1 nt", pack_avx2<8192, _MM_HINT_T1, true>},
        {"avx2 pf16K t1 nt", pack_avx2<16384, _MM_HINT_T1, true>},
        {"avx2 pf8K t2 nt", pack_avx2<8192, _MM_HINT_T2, true>},
        {"avx2 pf2K t0 plain stores", pack_avx2<2048, _MM_HINT_T0, false>},
        {"avx512 no-pf nt", pack_avx512<0, true>},
        {"avx512 pf2K nt", pack_avx512<2048, true>},
        {"avx512 pf4K nt", pack_avx512<4096, true>},
    };
    const unsigned hc = std::thread::hardware_concurrency();
    std::vector<int> counts = {1, 4, 8};
    if (hc > 8) counts.push_back((int)hc - 1);
    for (int threads : counts) {
        std::vector<uint8_t*> in(threads), out(threads);
        std::vector<std::thread> pool;
        for (int k = 0; k < threads; ++k) pool.emplace_back([&, k] {
            in[k] = (uint8_t*)aligned_alloc(64, per_thread + 32768);
            out[k] = (uint8_t*)aligned_alloc(64, per_thread / 4 + 64);
            for (size_t i = 0; i < per_thread + 32768; ++i) in[k][i] = (uint8_t)((i * 2654435761u) >> 30);
            memset(out[k], 0, per_thread / 4 + 64);
        });
        for (auto& t : pool) t.join();
        for (const Variant& v : vs) {
            if (!strncmp(v.name, "avx512", 6) && !(__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw"))) continue;
            double best = 0;
            for (int rep = 0; rep < 3; ++rep) {
                pool.clear();
                const auto t0 = std::chrono::steady_clock::now();
                for (int k = 0; k < threads; ++k) pool.emplace_back([&, k] {
                    // pieces of 2 MiB of input, as the PACK lanes see them
                    for (size_t c = 0; c < per_thread; c += (2u << 20)) v.fn(in[k] + c, out[k] + c / 4, (2u << 20));
                });
                for (auto& t : pool) t.join();
                const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                const double gbs = (double)per_thread * threads / dt / 1e9;
                if (gbs > best) best = gbs;
            }
            printf("{\"threads\": %d, \"variant\": \"%s\", \"input_GBps\": %.1f, \"per_thread_GBps\": %.2f}\n", threads, v.name, best, best / threads);
            fflush(stdout);
        }
        for (int k = 0; k < threads; ++k) { free(in[k]); free(out[k]); }
    }
    return 0;
}
