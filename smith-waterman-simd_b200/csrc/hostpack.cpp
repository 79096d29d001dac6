// hostpack.cpp -- byte codes -> the reference's 2-bit packing, on the HOST.
//
// The inverse of the reference's `unpack` (/root/reference/source.cpp:1580-1583:
// dest[i*4+j] = (src[i] >> 2j) & 3), so packed[i] = c[4i] | c[4i+1]<<2 | c[4i+2]<<4 | c[4i+3]<<6.
// This is wire-format compression for the PCIe link (256 B -> 64 B per pair), nothing more:
// no scoring happens on the host, and the device expands the bytes again with the unpack kernel.
// Codes are masked to two bits (a code above 3 is the caller's error, include/swb200.h).
#include <stddef.h>
#include <stdint.h>
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
__attribute__((target("avx2")))
static void pack2bit_avx2(const uint8_t* codes, uint8_t* packed, size_t n_codes)
{
    const __m256i m3 = _mm256_set1_epi8(3);
    const __m256i w14 = _mm256_set1_epi16(0x0401);          // bytes (1,4): c0 + 4*c1 per 16-bit lane
    const __m256i w116 = _mm256_set1_epi32(0x00100001);     // words (1,16): n0 + 16*n1 per 32-bit lane
    const __m256i gather = _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
                                            0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
    size_t i = 0;
    for (; i + 64 <= n_codes; i += 64) {
        __m256i a = _mm256_and_si256(_mm256_loadu_si256((const __m256i*)(codes + i)), m3);
        __m256i b = _mm256_and_si256(_mm256_loadu_si256((const __m256i*)(codes + i + 32)), m3);
        a = _mm256_madd_epi16(_mm256_maddubs_epi16(a, w14), w116);
        b = _mm256_madd_epi16(_mm256_maddubs_epi16(b, w14), w116);
        a = _mm256_shuffle_epi8(a, gather);
        b = _mm256_shuffle_epi8(b, gather);
        uint32_t o[4];
        o[0] = (uint32_t)_mm256_extract_epi32(a, 0); o[1] = (uint32_t)_mm256_extract_epi32(a, 4);
        o[2] = (uint32_t)_mm256_extract_epi32(b, 0); o[3] = (uint32_t)_mm256_extract_epi32(b, 4);
        memcpy(packed + i / 4, o, 16);
    }
    if (i < n_codes) pack2bit_swar(codes + i, packed + i / 4, n_codes - i);
}
#endif

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
