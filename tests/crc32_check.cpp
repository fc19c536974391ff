// crc32_buffer (carry-less multiplication where available) against zlib's crc32 on every length 0..600 and on random
// lengths / alignments up to a BGZF block; prints "bad 0" when they agree everywhere.
#include <cstdio>
#include <random>
#include <vector>
#include "crc32_fast.hpp"
int main()
{
    std::mt19937_64 rng(1);
    std::vector<uint8_t> b((1 << 17) + 64);
    for (auto &x : b) x = (uint8_t)rng();
    size_t bad = 0;
    for (int it = 0; it < 30000; ++it) {
        const size_t off = rng() % 64;
        size_t len = rng() % 66000;
        if (it <= 600) len = (size_t)it;
        const uint8_t *p = b.data() + off;
        if (inqhost::crc32_buffer(p, len) != (uint32_t)crc32(crc32(0L, Z_NULL, 0), p, (uInt)len)) ++bad;
    }
    printf("bad %zu\n", bad);
    return bad ? 1 : 0;
}
