// hostpack.cpp -- byte codes -> the reference's 2-bit packing, on the HOST.
//
// The inverse of the reference's `unpack` (/root/reference/source.cpp:1580-1583:
// dest[i*4+j] = (src[i] >> 2j) & 3), so packed[i] = c[4i] | c[4i+1]<<2 | c[4i+2]<<4 | c[4i+3]<<6.
// This is wire-format compression for the PCIe link (256 B -> 64 B per pair), nothing more:
// no scoring happens on the host, and the device expands the bytes again with the unpack kernel.
// Codes are masked to two bits (a code above 3 is the caller's error, include/swb200.h).
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace swb {

static inline uint16_t pack8_swar(uint64_t x)
{
    x &= 0x0303030303030303ull;
    x = (x | (x >> 6)) & 0x000f000f000f000full;
    x = (x | (x >> 12)) & 0x000000ff000000ffull;
    return (uint16_t)(x | (x >> 24));
}

static void pack2bit_swar(const uint8_t* codes, uint8_t* packed, size_t n_codes)
{
    for (size_t i = 0; i + 8 <= n_codes; i += 8) {
        uint64_t x;
        memcpy(&x, codes + i, 8);
        const uint16_t p = pack8_swar(x);
        memcpy(packed + i / 4, &p, 2);
    }
}

#if defined(__x86_64__)
// 128 codes -> 32 packed bytes per iteration: two multiply-adds turn 4 codes into one byte value per 32-bit lane
// (c0 + 4 c1, then n0 + 16 n1), two saturating packs bring the 32 values of four registers together, one cross-lane
// permute restores their order.  One 32-byte store per iteration; when the destination is 32-byte aligned (the
// library's pinned staging always is) the store is non-temporal: the packed bytes are read next by the DMA engine,
// not by this core, so they need neither a read-for-ownership nor a place in the cache.
// A core streams from DRAM at what its fill buffers allow (~7-10 GB/s on the hosts measured, 8 threads saturate at
// 56 GB/s, profiles/r01/hostpack_rate_16core_box.jsonl), so the input is also prefetched 2 KiB ahead of the loads.
// In the authoring container (one thread, 128 MB): 6.4 GB/s for the previous 64-byte loop, 8.0 with this loop shape,
// 9.1 with the prefetch, 9.8 with the non-temporal store as well.  On the B200 box (tools/packbench, profiles/r02/
// packbench_b200_box.jsonl): prefetching into L2 (T1) 8-16 KiB ahead instead of into L1 2 KiB ahead lifts one thread from
// 16.3 to 27-28 GB/s and eight from 101 to 117 (a core's L1 has ~a dozen fill buffers, its L2 several times as many
// requests in flight); fifteen threads sit at the box's DRAM bandwidth either way (140 -> 144 GB/s).
template <bool STREAM>
__attribute__((target("avx2")))
static void pack2bit_avx2_body(const uint8_t* codes, uint8_t* packed, size_t n_codes)
{
    const __m256i m3 = _mm256_set1_epi8(3);
    const __m256i w14 = _mm256_set1_epi16(0x0401);          // bytes (1,4): c0 + 4*c1 per 16-bit lane
    const __m256i w116 = _mm256_set1_epi32(0x00100001);     // words (1,16): n0 + 16*n1 per 32-bit lane
    const __m256i order = _mm256_setr_epi32(0, 4, 1, 5, 2, 6, 3, 7);   // undoes the in-lane interleave of the two packs
    size_t i = 0;
    for (; i + 128 <= n_codes; i += 128) {
        _mm_prefetch((const char*)(codes + i + 8192), _MM_HINT_T1);      // into L2, 8 KiB ahead: see the note above
        _mm_prefetch((const char*)(codes + i + 8256), _MM_HINT_T1);
        __m256i a = _mm256_and_si256(_mm256_loadu_si256((const __m256i*)(codes + i)), m3);
        __m256i b = _mm256_and_si256(_mm256_loadu_si256((const __m256i*)(codes + i + 32)), m3);
        __m256i c = _mm256_and_si256(_mm256_loadu_si256((const __m256i*)(codes + i + 64)), m3);
        __m256i d = _mm256_and_si256(_mm256_loadu_si256((const __m256i*)(codes + i + 96)), m3);
        a = _mm256_madd_epi16(_mm256_maddubs_epi16(a, w14), w116);      // 8 x 32-bit, each one packed byte (0..255)
        b = _mm256_madd_epi16(_mm256_maddubs_epi16(b, w14), w116);
        c = _mm256_madd_epi16(_mm256_maddubs_epi16(c, w14), w116);
        d = _mm256_madd_epi16(_mm256_maddubs_epi16(d, w14), w116);
        const __m256i ab = _mm256_packus_epi32(a, b);                   // lanes: a0-3 b0-3 | a4-7 b4-7   (16-bit)
        const __m256i cd = _mm256_packus_epi32(c, d);
        __m256i r = _mm256_packus_epi16(ab, cd);                        // a0-3 b0-3 c0-3 d0-3 | a4-7 b4-7 c4-7 d4-7  (bytes)
        r = _mm256_permutevar8x32_epi32(r, order);                      // a0-7 b0-7 c0-7 d0-7
        if (STREAM) _mm256_stream_si256((__m256i*)(packed + i / 4), r);
        else        _mm256_storeu_si256((__m256i*)(packed + i / 4), r);
    }
    if (STREAM) _mm_sfence();                                            // before the caller hands the buffer to the DMA engine
    if (i < n_codes) pack2bit_swar(codes + i, packed + i / 4, n_codes - i);
}

static void pack2bit_avx2(const uint8_t* codes, uint8_t* packed, size_t n_codes)
{
    // non-temporal stores only for buffers big enough that the cache could not hold them anyway.
    // SWB200_PACK_STREAM=0 keeps plain stores (tuning knob: a small pinned staging ring then stays in the core's L2, where
    // the DMA engine's reads can be served without a trip to DRAM).
    static const bool stream = [] { const char* e = getenv("SWB200_PACK_STREAM"); return !(e && e[0] == '0'); }();
    if (stream && (((uintptr_t)packed) & 31u) == 0 && n_codes >= (1u << 16)) pack2bit_avx2_body<true>(codes, packed, n_codes);
    else pack2bit_avx2_body<false>(codes, packed, n_codes);
}
#endif

// ---------------------------------------------------------------------------------------------------------------
// The other direction, the reference's `unpack` itself (source.cpp:1580-1583) on the host: the semi-global aligner's
// move strings cross the link four to a byte and are expanded into the caller's byte rows here (sg_pipe.inc).
static void unpack2bit_plain(const uint8_t* packed, uint8_t* codes, size_t n_codes)
{
    for (size_t i = 0; i < n_codes / 4; ++i) {
        const uint32_t b = packed[i];
        const uint32_t w = (b & 3u) | ((b & 0x0cu) << 6) | ((b & 0x30u) << 12) | ((b & 0xc0u) << 18);
        memcpy(codes + 4 * i, &w, 4);
    }
    for (size_t j = n_codes & ~(size_t)3; j < n_codes; ++j) codes[j] = (packed[j / 4] >> (2 * (j & 3))) & 3u;
}

#if defined(__x86_64__)
// 32 packed bytes -> 128 codes per iteration: the four 2-bit fields of every byte as four vectors (shift, mask), then two
// rounds of interleaves put field j of byte i at 4i + j; the interleaves work inside 128-bit halves, four half-swaps
// restore the order.  19 vector operations per 128 bytes written.
template <bool STREAM>
__attribute__((target("avx2")))
static void unpack2bit_avx2_body(const uint8_t* packed, uint8_t* codes, size_t n_codes)
{
    const __m256i m3 = _mm256_set1_epi8(3);
    size_t i = 0;
    for (; i + 128 <= n_codes; i += 128) {
        const __m256i v = _mm256_loadu_si256((const __m256i*)(packed + i / 4));
        const __m256i v0 = _mm256_and_si256(v, m3);
        const __m256i v1 = _mm256_and_si256(_mm256_srli_epi16(v, 2), m3);
        const __m256i v2 = _mm256_and_si256(_mm256_srli_epi16(v, 4), m3);
        const __m256i v3 = _mm256_and_si256(_mm256_srli_epi16(v, 6), m3);
        const __m256i a_lo = _mm256_unpacklo_epi8(v0, v1), a_hi = _mm256_unpackhi_epi8(v0, v1);     // (field 0, field 1) of bytes 0-7 | 16-23, 8-15 | 24-31
        const __m256i b_lo = _mm256_unpacklo_epi8(v2, v3), b_hi = _mm256_unpackhi_epi8(v2, v3);
        const __m256i c0 = _mm256_unpacklo_epi16(a_lo, b_lo), c1 = _mm256_unpackhi_epi16(a_lo, b_lo);   // bytes 0-3 | 16-19, 4-7 | 20-23
        const __m256i c2 = _mm256_unpacklo_epi16(a_hi, b_hi), c3 = _mm256_unpackhi_epi16(a_hi, b_hi);   // bytes 8-11 | 24-27, 12-15 | 28-31
        const __m256i o0 = _mm256_permute2x128_si256(c0, c1, 0x20), o1 = _mm256_permute2x128_si256(c2, c3, 0x20);
        const __m256i o2 = _mm256_permute2x128_si256(c0, c1, 0x31), o3 = _mm256_permute2x128_si256(c2, c3, 0x31);
        if (STREAM) {
            _mm256_stream_si256((__m256i*)(codes + i), o0);      _mm256_stream_si256((__m256i*)(codes + i + 32), o1);
            _mm256_stream_si256((__m256i*)(codes + i + 64), o2); _mm256_stream_si256((__m256i*)(codes + i + 96), o3);
        } else {
            _mm256_storeu_si256((__m256i*)(codes + i), o0);      _mm256_storeu_si256((__m256i*)(codes + i + 32), o1);
            _mm256_storeu_si256((__m256i*)(codes + i + 64), o2); _mm256_storeu_si256((__m256i*)(codes + i + 96), o3);
        }
    }
    if (STREAM) _mm_sfence();
    if (i < n_codes) unpack2bit_plain(packed + i / 4, codes + i, n_codes - i);
}
#endif

// codes[4i + j] = (packed[i] >> 2j) & 3 for n_codes codes (any count; the last packed byte may be partly used).
void unpack2bit_host(const uint8_t* packed, uint8_t* codes, size_t n_codes)
{
#if defined(__x86_64__)
    static const bool have_avx2 = __builtin_cpu_supports("avx2");
    if (have_avx2) {
        // non-temporal stores when the destination allows: the expanded rows are the caller's, this core will not read them
        if ((((uintptr_t)codes) & 31u) == 0 && n_codes >= 4096) unpack2bit_avx2_body<true>(packed, codes, n_codes);
        else unpack2bit_avx2_body<false>(packed, codes, n_codes);
        return;
    }
#endif
    unpack2bit_plain(packed, codes, n_codes);
}

// n_codes must be a multiple of 8 (sequences are multiples of 128).
void pack2bit_host(const uint8_t* codes, uint8_t* packed, size_t n_codes)
{
#if defined(__x86_64__)
    static const bool have_avx2 = __builtin_cpu_supports("avx2");
    if (have_avx2) { pack2bit_avx2(codes, packed, n_codes); return; }
#endif
    pack2bit_swar(codes, packed, n_codes);
}

} // namespace swb
