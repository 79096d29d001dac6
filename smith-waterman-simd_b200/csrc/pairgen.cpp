// pairgen.cpp -- synthetic inputs for the verification and benchmark harness (host code).
//
//  * swb200_gen_reference_stream: the input stream of the reference's differential test
//    (/root/reference/source.cpp:2944-2953): std::mt19937_64 seeded with 10000, one draw per
//    base, a[i] and b[i] drawn alternately.  The reference maps a draw to a base with
//    std::uniform_int_distribution<int>(0,3), whose output is implementation-defined; under
//    libstdc++ it is the top two bits of the draw, which is what is stated
//    here so that the stream -- and the known-answer checksums that go with it -- is the
//    same on every toolchain.
//  * swb200_gen_counter_pairs: a counter-based stream for the 100 M-pair and streaming
//    configurations: pair k is a pure function of (seed, k), so any index range can be
//    produced by any thread, rank or GPU shard.  One splitmix64
//    draw yields 32 bases; a pair takes 8 draws.
//  * swb200_gen_related_pairs: long related pairs in the manner of the reference's TestSemiGlobal
//    (source.cpp:2750-2771: seq2 = seq1 with 10 % mismatches, 10 % insertions, 10 % deletions), again
//    counter-based (pair k depends only on (seed, k)) and with the three rates as arguments.
//  * swb200_fnv1a64_i32: the score checksum SURVEY.md §8(c) defines.
#include "../../include/swb200.h"

#include <cstring>
#include <exception>
#if defined(__x86_64__)
#include <immintrin.h>
#endif
#include <random>
#include <thread>
#include <vector>

namespace {

inline uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// word w (0..7) of pair k: bases 32w .. 32w+31 of the 256 bases (seq1 then seq2) of the pair
inline uint64_t pair_word(uint64_t seed, uint64_t k, unsigned w)
{
    return splitmix64(k * 8ull + w + seed * 0x9E3779B97F4A7C15ull);
}

// 64 bits = 32 bases of 2 bits -> 32 byte codes, base i from bits 2i (the reference's `unpack`, source.cpp:1580-1583).
// The byte-coded stream was bound by this loop (profiles/r01/stream_100m_bytes_1gpu.json: 1.32 of 1.34 s in host
// generation), so on x86-64 it is done in registers: the four 2-bit planes of the eight source bytes (x, x>>2, x>>4,
// x>>6) interleaved byte-wise and then word-wise put plane k of byte i at output 4i+k; one mask at the end.
template <bool STREAM = false>
inline void expand32(uint64_t x, uint8_t* d)
{
#if defined(__SSE2__)
    const __m128i v0 = _mm_cvtsi64_si128((long long)x);
    const __m128i p01 = _mm_unpacklo_epi8(v0, _mm_srli_epi64(v0, 2));                         // b0>>0, b0>>2, b1>>0, b1>>2, ...
    const __m128i p23 = _mm_unpacklo_epi8(_mm_srli_epi64(v0, 4), _mm_srli_epi64(v0, 6));      // b0>>4, b0>>6, b1>>4, ...
    const __m128i m3 = _mm_set1_epi8(3);
    const __m128i lo = _mm_and_si128(_mm_unpacklo_epi16(p01, p23), m3);           // source bytes 0..3 -> codes 0..15
    const __m128i hi = _mm_and_si128(_mm_unpackhi_epi16(p01, p23), m3);           // source bytes 4..7 -> codes 16..31
    if (STREAM) {      // the consumer of a stream buffer is the DMA engine or a packing core, not this core: no read-for-ownership, no cache line
        _mm_stream_si128((__m128i*)d, lo);
        _mm_stream_si128((__m128i*)(d + 16), hi);
    } else {
        _mm_storeu_si128((__m128i*)d, lo);
        _mm_storeu_si128((__m128i*)(d + 16), hi);
    }
#else
    for (int i = 0; i < 32; ++i) d[i] = (uint8_t)((x >> (2 * i)) & 3);
#endif
}

#if defined(__x86_64__)
// The packed stream is 64 raw bytes per pair -- its eight splitmix64 draws -- so with AVX-512 (64-bit lane multiply,
// vpmullq) one vector computes a whole pair: lanes 0..3 are seq1's 32 packed bytes, lanes 4..7 seq2's.  Eight GPUs
// consume about 3.4 G packed pairs/s; the scalar loop gives 46 M pairs/s per thread.
template <bool STREAM>
__attribute__((target("avx512f,avx512dq")))
void gen_range_packed_avx512(uint64_t seed, uint64_t first, uint64_t lo, uint64_t hi, uint8_t* seq1, uint8_t* seq2)
{
    const __m512i lane = _mm512_setr_epi64(0, 1, 2, 3, 4, 5, 6, 7);
    const __m512i g = _mm512_set1_epi64((long long)0x9E3779B97F4A7C15ull);
    const __m512i m1 = _mm512_set1_epi64((long long)0xBF58476D1CE4E5B9ull);
    const __m512i m2 = _mm512_set1_epi64((long long)0x94D049BB133111EBull);
    const __m512i base = _mm512_add_epi64(lane, _mm512_set1_epi64((long long)(seed * 0x9E3779B97F4A7C15ull)));
    for (uint64_t p = lo; p < hi; ++p) {
        __m512i x = _mm512_add_epi64(base, _mm512_set1_epi64((long long)((first + p) * 8ull)));   // pair_word's argument, 8 lanes
        x = _mm512_add_epi64(x, g);                                                               // splitmix64
        x = _mm512_mullo_epi64(_mm512_xor_si512(x, _mm512_srli_epi64(x, 30)), m1);
        x = _mm512_mullo_epi64(_mm512_xor_si512(x, _mm512_srli_epi64(x, 27)), m2);
        x = _mm512_xor_si512(x, _mm512_srli_epi64(x, 31));
        if (STREAM) {      // large ranges into aligned (pinned) buffers: nobody reads them from this core's cache, see gen_range
            _mm256_stream_si256((__m256i*)(seq1 + p * 32), _mm512_castsi512_si256(x));
            _mm256_stream_si256((__m256i*)(seq2 + p * 32), _mm512_extracti64x4_epi64(x, 1));
        } else {
            _mm256_storeu_si256((__m256i*)(seq1 + p * 32), _mm512_castsi512_si256(x));
            _mm256_storeu_si256((__m256i*)(seq2 + p * 32), _mm512_extracti64x4_epi64(x, 1));
        }
    }
    if (STREAM) _mm_sfence();
}
#endif

void gen_range(uint64_t seed, uint64_t first, uint64_t lo, uint64_t hi, uint8_t* seq1, uint8_t* seq2, bool packed)
{
#if defined(__x86_64__)
    static const bool have_avx512 = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512dq");
    if (packed && have_avx512) {
        if (hi - lo >= 4096 && ((((uintptr_t)seq1) | ((uintptr_t)seq2)) & 31u) == 0) gen_range_packed_avx512<true>(seed, first, lo, hi, seq1, seq2);
        else gen_range_packed_avx512<false>(seed, first, lo, hi, seq1, seq2);
        return;
    }
#endif
#if defined(__SSE2__)
    // Large byte-coded ranges into 16-byte aligned arrays (the pinned ring buffers of the streaming mode) are written with
    // non-temporal stores: round 2's 100 M-pair byte stream was bound by this loop's memory traffic (38 GB/s of stores plus
    // as much read-for-ownership on 15 threads), and nobody reads the lines from this core's cache afterwards.
    if (!packed && hi - lo >= 4096 && ((((uintptr_t)seq1) | ((uintptr_t)seq2)) & 15u) == 0) {
        for (uint64_t p = lo; p < hi; ++p)
            for (unsigned w = 0; w < 8; ++w)
                expand32<true>(pair_word(seed, first + p, w), ((w < 4) ? seq1 : seq2) + p * 128 + (w & 3) * 32);
        _mm_sfence();
        return;
    }
#endif
    for (uint64_t p = lo; p < hi; ++p) {
        for (unsigned w = 0; w < 8; ++w) {
            const uint64_t x = pair_word(seed, first + p, w);
            uint8_t* dst = (w < 4) ? seq1 : seq2;
            if (packed) {
                // 32 bases = 8 packed bytes; byte i holds bases 4i..4i+3, base j at bits 2j (source.cpp:1580-1583)
                std::memcpy(dst + p * 32 + (w & 3) * 8, &x, 8);
            } else {
                expand32(x, dst + p * 128 + (w & 3) * 32);
            }
        }
    }
}

// [0, n) in `threads` contiguous parts, one host thread each.  If a thread cannot be started, the parts that have no
// thread are done by the caller's: a generator has no reason to fail, and the C ABI lets no exception out.
template <class Body>
int run_split(int threads, uint64_t n, Body body)
{
    std::vector<std::thread> pool;
    int started = 0;
    try {
        pool.reserve((size_t)threads);
        for (; started < threads; ++started) {
            const uint64_t lo = n * (uint64_t)started / threads, hi = n * (uint64_t)(started + 1) / threads;
            pool.emplace_back([&body, lo, hi] { body(lo, hi); });
        }
    } catch (const std::exception&) {
    }
    if (started < threads) body(n * (uint64_t)started / threads, n);
    for (auto& th : pool) th.join();
    return SWB200_OK;
}

int gen_counter(uint64_t seed, uint64_t first, uint64_t n, uint8_t* seq1, uint8_t* seq2, int threads, bool packed)
{
    if (n && (!seq1 || !seq2)) return SWB200_ERR_ARG;
    if (threads < 1) threads = 1;
    if ((uint64_t)threads > n) threads = n ? (int)n : 1;
    if (threads == 1) { gen_range(seed, first, 0, n, seq1, seq2, packed); return SWB200_OK; }
    return run_split(threads, n, [&](uint64_t lo, uint64_t hi) { gen_range(seed, first, lo, hi, seq1, seq2, packed); });
}

// TestSemiGlobal's construction (source.cpp:2750-2771) with a per-pair splitmix64 stream instead of
// the shared mt19937_64: a = iid bases; b walks a, each step drawing p in [0,100).
void gen_related_range(uint64_t seed, uint64_t first, uint64_t lo, uint64_t hi, int len, int sub_pct, int ins_pct, int del_pct,
                       uint8_t* seq1, uint8_t* seq2)
{
    for (uint64_t p = lo; p < hi; ++p) {
        uint64_t state = splitmix64((first + p) * 0x9E3779B97F4A7C15ull + seed);
        uint64_t bits = 0; int have = 0;
        auto draw = [&](int nbits) -> uint32_t {
            if (have < nbits) { state = splitmix64(state); bits = state; have = 64; }
            const uint32_t v = (uint32_t)(bits & ((1ull << nbits) - 1));
            bits >>= nbits; have -= nbits;
            return v;
        };
        auto pct = [&]() -> int { uint32_t v; do { v = draw(7); } while (v >= 100); return (int)v; };   // uniform on 0..99
        uint8_t* a = seq1 + p * (uint64_t)len;
        uint8_t* b = seq2 + p * (uint64_t)len;
        for (int i = 0; i < len; ++i) a[i] = (uint8_t)draw(2);
        for (int i = 0, j = 0; i < len;) {
            if (j == len) { b[i++] = (uint8_t)draw(2); continue; }
            const int q = pct();
            if (q < sub_pct) { b[i++] = (uint8_t)draw(2); ++j; }
            else if (q < sub_pct + ins_pct) { b[i++] = (uint8_t)draw(2); }
            else if (q < sub_pct + ins_pct + del_pct) { ++j; }
            else { b[i++] = a[j++]; }
        }
    }
}

} // namespace

extern "C" {

int swb200_gen_related_pairs(uint64_t seed, uint64_t first, uint64_t n, int32_t seq_len, int sub_pct, int ins_pct, int del_pct,
                             uint8_t* seq1, uint8_t* seq2, int threads)
{
    if (n && (!seq1 || !seq2)) return SWB200_ERR_ARG;
    if (seq_len < 1 || sub_pct < 0 || ins_pct < 0 || del_pct < 0 || sub_pct + ins_pct + del_pct > 100) return SWB200_ERR_ARG;
    if (threads < 1) threads = 1;
    if ((uint64_t)threads > n) threads = n ? (int)n : 1;
    return run_split(threads, n, [&](uint64_t lo, uint64_t hi) {
        gen_related_range(seed, first, lo, hi, (int)seq_len, sub_pct, ins_pct, del_pct, seq1, seq2);
    });
}


int swb200_gen_reference_stream(uint64_t seed, uint64_t n, uint8_t* seq1, uint8_t* seq2)
{
    if (n && (!seq1 || !seq2)) return SWB200_ERR_ARG;
    std::mt19937_64 rnd(seed);
    for (uint64_t p = 0; p < n; ++p)
        for (int i = 0; i < SWB200_SEQ_LEN; ++i) {
            seq1[p * SWB200_SEQ_LEN + i] = (uint8_t)(rnd() >> 62);
            seq2[p * SWB200_SEQ_LEN + i] = (uint8_t)(rnd() >> 62);
        }
    return SWB200_OK;
}

int swb200_gen_counter_pairs(uint64_t seed, uint64_t first, uint64_t n, uint8_t* seq1, uint8_t* seq2, int threads)
{
    return gen_counter(seed, first, n, seq1, seq2, threads, false);
}

int swb200_gen_counter_pairs_packed(uint64_t seed, uint64_t first, uint64_t n, uint8_t* seq1_packed, uint8_t* seq2_packed, int threads)
{
    return gen_counter(seed, first, n, seq1_packed, seq2_packed, threads, true);
}

uint64_t swb200_fnv1a64_i32(const int32_t* scores, uint64_t n)
{
    uint64_t h = 1469598103934665603ull;
    for (uint64_t i = 0; i < n; ++i) h = (h ^ (uint64_t)(uint32_t)scores[i]) * 1099511628211ull;
    return h;
}

} // extern "C"
