// SWAR (SIMD-within-a-register) decode of 16 ASCII bases per lane, shared by the pack and the histogram
// kernels.  Plain C++ with host fallbacks for the three device intrinsics, so that tests/test_swar_host.py
// can compile this header with g++ and check every function exhaustively on the CPU.
//
// Semantics restated from the reference (scripts/kmer.py:183-196): only the bytes 'A' 'T' 'G' 'C' are symbols
// (A=0 T=1 G=2 C=3); every other byte value (lower case, N, IUPAC, digits, '-', blank ...) is a blank that
// voids each window it touches (scripts/kmer.py:49).
#pragma once
#include <stdint.h>

#if defined(__CUDA_ARCH__)
#define PHM_HD __host__ __device__ __forceinline__
#elif defined(__CUDACC__)
#define PHM_HD __host__ __device__ inline
#else
#define PHM_HD inline
#endif

namespace phm {

PHM_HD uint32_t byte_perm(uint32_t a, uint32_t b, uint32_t sel) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(a, b, sel);
#else
    uint64_t both = ((uint64_t)b << 32) | a;
    uint32_t out = 0;
    for (int i = 0; i < 4; ++i) {
        uint32_t n = (sel >> (4 * i)) & 0x7;               // selectors used here never set the sign bit
        out |= (uint32_t)((both >> (8 * n)) & 0xFF) << (8 * i);
    }
    return out;
#endif
}

PHM_HD uint32_t bit_reverse(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __brev(x);
#else
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
    x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
    return (x >> 16) | (x << 16);
#endif
}

// low 32 bits of ((hi:lo) >> shift), 0 <= shift < 32
PHM_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t shift) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, shift);
#else
    return (uint32_t)(((((uint64_t)hi) << 32) | lo) >> (shift & 31));
#endif
}

// ---------------------------------------------------------------------------------------------------
// 16 ASCII bytes (w[0] = bytes 0..3, little endian) -> one 32-bit "stream" word holding the 16 two-bit
// reference codes, FIRST base in the TOP two bits (base b in bits 31-2b : 30-2b), so that the bin index
// of the window starting at base p is the plain bit field  (stream >> (32 - 2p - 2k)) & (4^k - 1).
//
// Per word: bits 2:1 of an ASCII base are a 2-bit code already (A=0 C=1 T=2 G=3).  Masking them and
// multiplying by 2^23+2^17+2^11+2^5 gathers the four fields into the top byte (the cross terms land on
// disjoint lower bits, so there are no carries).  Three byte permutes assemble the 16 codes, a bit reversal
// puts the first base on top (swapping the two bits of each code), and one xor turns the swapped native
// code (l,h) into the reference code (A=0 T=1 G=2 C=3): hi' = l, lo' = h ^ l.
// Bytes that are not ATGC produce garbage codes here; validity is a separate test.
// ---------------------------------------------------------------------------------------------------
PHM_HD uint32_t codes16_be(const uint32_t w[4]) {
    const uint32_t M = 0x00820820u;
    uint32_t p0 = (w[0] & 0x06060606u) * M;
    uint32_t p1 = (w[1] & 0x06060606u) * M;
    uint32_t p2 = (w[2] & 0x06060606u) * M;
    uint32_t p3 = (w[3] & 0x06060606u) * M;
    uint32_t q01 = byte_perm(p0, p1, 0x0073u);
    uint32_t q23 = byte_perm(p2, p3, 0x0073u);
    uint32_t le = byte_perm(q01, q23, 0x5410u);
    uint32_t x = bit_reverse(le);
    return x ^ ((x >> 1) & 0x55555555u);
}

// Non-zero iff at least one of the 4 bytes is not one of 'A' 'C' 'G' 'T'.
// T is the only symbol with bit2 = 1 and bit1 = 0; the other six bits of a symbol are 0x50 for T and 0x41 for
// A, C, G.  (Carries cannot cross bytes: the per-byte T flag is 0 or 1 and 0x41 + 0x0F < 0x100.)
PHM_HD uint32_t invalid_bytes(uint32_t w) {
    uint32_t t = (w >> 2) & ~(w >> 1) & 0x01010101u;
    uint32_t expect = t * 0x0Fu + 0x41414141u;
    return (w ^ expect) & 0xF9F9F9F9u;
}

PHM_HD uint32_t any_invalid16(const uint32_t w[4]) {
    return invalid_bytes(w[0]) | invalid_bytes(w[1]) | invalid_bytes(w[2]) | invalid_bytes(w[3]);
}

// Precise blank mask of 16 bytes in the same big-endian two-bits-per-base layout as codes16_be: both bits of
// base b (bits 31-2b, 30-2b) are set iff byte b is not a symbol.  Slow path only.
PHM_HD uint32_t blank_mask16_be(const uint32_t w[4]) {
    uint32_t le = 0;
    for (int j = 0; j < 4; ++j) {
        uint32_t v = invalid_bytes(w[j]);
        // bit 7 of every non-zero byte
        uint32_t nz = (v | ((v & 0x7F7F7F7Fu) + 0x7F7F7F7Fu)) & 0x80808080u;
        // gather bits 7,15,23,31 -> even bits 0,2,4,6 of the top byte (same no-carry multiply as above)
        uint32_t g = ((nz >> 7) * 0x01041040u) >> 24;
        le |= (g | (g << 1)) << (8 * j);
    }
    return bit_reverse(le);
}

// Both bits of every base outside [lo, hi) (0 <= lo, hi <= 16, any order) in the big-endian layout.
PHM_HD uint32_t outside_mask16_be(int lo, int hi) {
    if (lo < 0) lo = 0;
    if (hi > 16) hi = 16;
    if (hi <= lo) return 0xFFFFFFFFu;
    uint32_t from_lo = (lo == 0) ? 0xFFFFFFFFu : (0xFFFFFFFFu >> (2 * lo));       // bases >= lo
    uint32_t from_hi = (hi == 16) ? 0u : (0xFFFFFFFFu >> (2 * hi));               // bases >= hi
    return ~(from_lo & ~from_hi);
}

// Window of W digits starting at base p (0..15) of `cur`, continuing into `nxt` (the following 16 bases).
// Returns the field pre-multiplied by 4 (a byte offset into a table of 32-bit counters).
template <int W>
PHM_HD uint32_t window_offset(uint32_t cur, uint32_t nxt, int p) {
    const uint32_t mask4 = ((1u << (2 * W)) - 1u) << 2;
    const int r = 64 - 2 * p - 2 * W;                      // position of the field's lowest bit in (cur:nxt)
    if (r >= 34) return (cur >> (r - 34)) & mask4;
    if (r >= 32) return (cur << (34 - r)) & mask4;
    return funnel_r(nxt, cur, (uint32_t)(r - 2)) & mask4;
}

template <int W>
PHM_HD uint32_t window_bits(uint32_t cur, uint32_t nxt, int p) {
    const uint32_t mask = (1u << (2 * W)) - 1u;
    const int r = 64 - 2 * p - 2 * W;
    if (r >= 32) return (cur >> (r - 32)) & mask;
    return funnel_r(nxt, cur, (uint32_t)r) & mask;
}

// Reverse complement of bin j for k digits in ATGC order (A<->T = 0<->1, G<->C = 2<->3): reverse the digits
// and flip the low bit of each.
PHM_HD uint32_t revcomp_bin(uint32_t j, int k) {
    uint32_t out = 0;
    for (int i = 0; i < k; ++i) {
        out = (out << 2) | ((j & 3u) ^ 1u);
        j >>= 2;
    }
    return out;
}

}  // namespace phm
