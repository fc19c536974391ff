// crc32_fast.hpp -- CRC-32 (the gzip / BGZF polynomial, reflected 0xEDB88320) of a whole buffer, by carry-less
// multiplication where the CPU has it (PCLMULQDQ + SSE4.1: four 128-bit lanes folded per 64 input bytes, then folded
// down and Barrett-reduced; the fold constants are x^(n) mod P for the distances involved), zlib's table-driven crc32
// otherwise. Every BGZF block the ingest accepts is checked against its trailer, including the blocks the GPU inflate
// engines produce, so at 15+ GB/s of inflated output the check itself is a host cost worth an order of magnitude.
// The vector path is used only after it has reproduced zlib's result on a probe buffer in this process.
#pragma once

#include <cstddef>
#include <cstdint>
#include <cstring>
#include <zlib.h>

#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>
#define INQ_CRC_X86 1
#endif

namespace inqhost {

#ifdef INQ_CRC_X86
// buf 16-byte alignment is not required; len >= 64 and a multiple of 16. Returns the updated (non-inverted) register.
__attribute__((target("pclmul,sse4.1"))) inline uint32_t crc32_clmul_body(uint32_t crc, const uint8_t *buf, size_t len)
{
    // x^(4*128+32), x^(4*128-32) | x^(128+32), x^(128-32) | x^64, - | P', mu   (all mod P, bit-reflected)
    alignas(16) static const uint64_t k1k2[2] = {0x0154442bd4ull, 0x01c6e41596ull};
    alignas(16) static const uint64_t k3k4[2] = {0x01751997d0ull, 0x00ccaa009eull};
    alignas(16) static const uint64_t k5k0[2] = {0x0163cd6124ull, 0x0000000000ull};
    alignas(16) static const uint64_t poly[2] = {0x01db710641ull, 0x01f7011641ull};
    __m128i x0, x1, x2, x3, x4, x5, x6, x7, x8, y5, y6, y7, y8;
    x1 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(buf + 0x00));
    x2 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(buf + 0x10));
    x3 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(buf + 0x20));
    x4 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(buf + 0x30));
    x1 = _mm_xor_si128(x1, _mm_cvtsi32_si128((int)crc));
    x0 = _mm_load_si128(reinterpret_cast<const __m128i *>(k1k2));
    buf += 64;
    len -= 64;
    while (len >= 64) {                                        // four lanes in flight
        x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
        x6 = _mm_clmulepi64_si128(x2, x0, 0x00);
        x7 = _mm_clmulepi64_si128(x3, x0, 0x00);
        x8 = _mm_clmulepi64_si128(x4, x0, 0x00);
        x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
        x2 = _mm_clmulepi64_si128(x2, x0, 0x11);
        x3 = _mm_clmulepi64_si128(x3, x0, 0x11);
        x4 = _mm_clmulepi64_si128(x4, x0, 0x11);
        y5 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(buf + 0x00));
        y6 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(buf + 0x10));
        y7 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(buf + 0x20));
        y8 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(buf + 0x30));
        x1 = _mm_xor_si128(_mm_xor_si128(x1, x5), y5);
        x2 = _mm_xor_si128(_mm_xor_si128(x2, x6), y6);
        x3 = _mm_xor_si128(_mm_xor_si128(x3, x7), y7);
        x4 = _mm_xor_si128(_mm_xor_si128(x4, x8), y8);
        buf += 64;
        len -= 64;
    }
    // four lanes -> one
    x0 = _mm_load_si128(reinterpret_cast<const __m128i *>(k3k4));
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
    x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
    x1 = _mm_xor_si128(_mm_xor_si128(x1, x3), x5);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
    x1 = _mm_xor_si128(_mm_xor_si128(x1, x4), x5);
    while (len >= 16) {                                        // the remaining whole 16-byte blocks
        x2 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(buf));
        x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
        x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
        x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
        buf += 16;
        len -= 16;
    }
    // 128 -> 64 bits
    x2 = _mm_clmulepi64_si128(x1, x0, 0x10);
    x3 = _mm_setr_epi32(~0, 0, ~0, 0);
    x1 = _mm_srli_si128(x1, 8);
    x1 = _mm_xor_si128(x1, x2);
    x0 = _mm_loadl_epi64(reinterpret_cast<const __m128i *>(k5k0));
    x2 = _mm_srli_si128(x1, 4);
    x1 = _mm_and_si128(x1, x3);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x1 = _mm_xor_si128(x1, x2);
    // Barrett reduction 64 -> 32 bits
    x0 = _mm_load_si128(reinterpret_cast<const __m128i *>(poly));
    x2 = _mm_and_si128(x1, x3);
    x2 = _mm_clmulepi64_si128(x2, x0, 0x10);
    x2 = _mm_and_si128(x2, x3);
    x2 = _mm_clmulepi64_si128(x2, x0, 0x00);
    x1 = _mm_xor_si128(x1, x2);
    return (uint32_t)_mm_extract_epi32(x1, 1);
}
#endif

// crc32 of buf[0, len) from scratch (what a BGZF trailer holds)
inline uint32_t crc32_buffer(const uint8_t *buf, size_t len)
{
#ifdef INQ_CRC_X86
    // the vector path must have the instructions and must agree with zlib on a probe (decided once per process)
    static const bool use_clmul = [] {
        if (!__builtin_cpu_supports("pclmul") || !__builtin_cpu_supports("sse4.1")) return false;
        uint8_t probe[64 * 5 + 16 * 3 + 7];
        uint32_t s = 0x9E3779B9u;
        for (size_t i = 0; i < sizeof(probe); ++i) { s = s * 1664525u + 1013904223u; probe[i] = (uint8_t)(s >> 24); }
        for (size_t n : {(size_t)64, (size_t)80, (size_t)128, sizeof(probe) / 16 * 16}) {
            const uint32_t a = ~crc32_clmul_body(0xFFFFFFFFu, probe, n);
            const uint32_t b = (uint32_t)crc32(crc32(0L, Z_NULL, 0), probe, (uInt)n);
            if (a != b) return false;
        }
        return true;
    }();
    if (use_clmul && len >= 64) {
        const size_t body = len & ~(size_t)15;
        uint32_t c = ~crc32_clmul_body(0xFFFFFFFFu, buf, body);
        if (body < len) c = (uint32_t)crc32(c, buf + body, (uInt)(len - body));      // the last < 16 bytes
        return c;
    }
#endif
    return (uint32_t)crc32(crc32(0L, Z_NULL, 0), buf, (uInt)len);
}

}  // namespace inqhost
