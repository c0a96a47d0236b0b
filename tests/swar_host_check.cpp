// Host-side exhaustive / randomised check of phamers_b200/csrc/kmer_swar.h (compiled with g++ by
// tests/test_swar_host.py).  Exit code 0 = all good.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include "../phamers_b200/csrc/kmer_swar.h"

static int ref_code(uint8_t c) {
    switch (c) { case 'A': return 0; case 'T': return 1; case 'G': return 2; case 'C': return 3; default: return -1; }
}

#define CHECK(cond) do { if (!(cond)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #cond); return 1; } } while (0)

int main() {
    // invalid_bytes: exhaustive over one byte in each lane position, others valid
    for (int lane = 0; lane < 4; ++lane)
        for (int c = 0; c < 256; ++c) {
            uint32_t w = 0x41434754u;                                   // "TGCA"
            w = (w & ~(0xFFu << (8 * lane))) | ((uint32_t)c << (8 * lane));
            bool bad = phm::invalid_bytes(w) != 0;
            CHECK(bad == (ref_code((uint8_t)c) < 0));
        }
    std::mt19937_64 rng(12345);
    const char alphabet[] = "ATGCATGCATGCATGCNatgcRYKMSWBDHVU-* \n\t0123@EPQ\x7f\xff";
    const int na = (int)sizeof(alphabet) - 1;
    for (int trial = 0; trial < 200000; ++trial) {
        uint8_t bytes[32];
        bool dirty = (trial % 3) == 0;
        for (int i = 0; i < 32; ++i)
            bytes[i] = dirty ? (uint8_t)alphabet[rng() % na] : (uint8_t)"ATGC"[rng() % 4];
        if (trial % 7 == 0) bytes[rng() % 32] = (uint8_t)(rng() & 0xFF);
        uint32_t w[4], w2[4];
        std::memcpy(w, bytes, 16);
        std::memcpy(w2, bytes + 16, 16);
        uint32_t cur = phm::codes16_be(w), nxt = phm::codes16_be(w2);
        uint32_t blank = phm::blank_mask16_be(w), blank2 = phm::blank_mask16_be(w2);
        bool any = false;
        for (int b = 0; b < 16; ++b) {
            int rc = ref_code(bytes[b]);
            uint32_t pair = (blank >> (30 - 2 * b)) & 3u;
            CHECK(pair == (rc < 0 ? 3u : 0u));
            if (rc >= 0) CHECK(((cur >> (30 - 2 * b)) & 3u) == (uint32_t)rc);
            any |= rc < 0;
        }
        CHECK((phm::any_invalid16(w) != 0) == any);
        // windows of every width at every start, crossing into the next 16 bases
        for (int p = 0; p < 16; ++p) {
#define WIN(W) { \
            uint32_t want = 0; bool ok = true; \
            for (int j = 0; j < W; ++j) { int rc = ref_code(bytes[p + j]); if (rc < 0) ok = false; want = want * 4 + (uint32_t)(rc & 3); } \
            if (ok) { \
                CHECK(phm::window_bits<W>(cur, nxt, p) == want); \
                CHECK(phm::window_offset<W>(cur, nxt, p) == want * 4); \
            } \
            bool blank_hit = phm::window_bits<W>(blank, blank2, p) != 0; \
            CHECK(blank_hit == !ok); }
            WIN(1) WIN(2) WIN(3) WIN(4) WIN(5) WIN(6)
#undef WIN
        }
    }
    // outside_mask16_be
    for (int lo = -3; lo <= 18; ++lo)
        for (int hi = -3; hi <= 18; ++hi) {
            uint32_t m = phm::outside_mask16_be(lo, hi);
            for (int b = 0; b < 16; ++b) {
                bool inside = b >= lo && b < hi;
                CHECK(((m >> (30 - 2 * b)) & 3u) == (inside ? 0u : 3u));
            }
        }
    // revcomp_bin is an involution and matches the string definition
    for (int k = 1; k <= 6; ++k)
        for (uint32_t j = 0; j < (1u << (2 * k)); ++j) {
            CHECK(phm::revcomp_bin(phm::revcomp_bin(j, k), k) == j);
            const char sym[] = "ATGC", comp[] = "TACG";
            char s[8], r[8];
            for (int i = 0; i < k; ++i) s[i] = sym[(j >> (2 * (k - 1 - i))) & 3];
            for (int i = 0; i < k; ++i) r[i] = comp[std::strchr(sym, s[k - 1 - i]) - sym];
            uint32_t want = 0;
            for (int i = 0; i < k; ++i) want = want * 4 + (uint32_t)(std::strchr(sym, r[i]) - sym);
            CHECK(phm::revcomp_bin(j, k) == want);
        }
    std::printf("swar ok\n");
    return 0;
}
