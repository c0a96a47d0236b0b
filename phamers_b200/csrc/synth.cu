// Synthetic metagenome generator (bench / tests only; not part of the timed path).
//
// Workload definition from SURVEY.md 8(d) config 2: contig lengths clip(round(exp(N(ln 10000, 1.0))), 1000, 100000),
// bases i.i.d. over ATGC with a per-contig GC fraction drawn from U(0.25, 0.75).  Everything is a pure function of
// (seed, global contig index, position), so a shard generated on any rank equals the same rows of the 1-GPU set.
#include "phm_common.cuh"

namespace phm {

__device__ __forceinline__ double u01(uint64_t h) {            // (0, 1)
    return ((double)(h >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

__global__ void synth_lengths_kernel(uint64_t seed, int64_t first, int64_t n, int64_t *lengths) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t key = mix64(seed ^ mix64((uint64_t)(first + i) * 0xD1342543DE82EF95ull + 1ull));
        const double u1 = u01(mix64(key ^ 0x1111ull));
        const double u2 = u01(mix64(key ^ 0x2222ull));
        const double z = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
        double len = rint(exp(log(10000.0) + z));
        len = fmin(fmax(len, 1000.0), 100000.0);
        lengths[i] = (int64_t)len;
    }
}

// 16 bases for (contig key, block index): byte b of the two hashes against the GC threshold picks the pair,
// bit b of a third hash picks the member of the pair.
__device__ __forceinline__ void synth_block16(uint64_t key, uint32_t thr, uint64_t block, uint8_t out[16]) {
    const uint64_t h0 = mix64(key ^ (block * 3ull + 0ull) * 0x9E3779B97F4A7C15ull);
    const uint64_t h1 = mix64(key ^ (block * 3ull + 1ull) * 0x9E3779B97F4A7C15ull);
    const uint64_t h2 = mix64(key ^ (block * 3ull + 2ull) * 0x9E3779B97F4A7C15ull);
#pragma unroll
    for (int b = 0; b < 16; ++b) {
        const uint32_t r = (uint32_t)(((b < 8 ? h0 : h1) >> (8 * (b & 7))) & 0xFFu);
        const bool gc = r < thr;
        const bool second = (h2 >> b) & 1ull;
        out[b] = gc ? (second ? 'C' : 'G') : (second ? 'T' : 'A');
    }
}

__global__ void synth_bases_kernel(uint64_t seed, int64_t first, int64_t n, const int64_t *__restrict__ off,
                                   uint8_t *__restrict__ seq) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t c = warp; c < n; c += n_warps) {
        const int64_t start = off[c], end = off[c + 1];
        const uint64_t key = mix64(seed ^ mix64((uint64_t)(first + c) * 0xD1342543DE82EF95ull + 1ull));
        const uint32_t thr = (uint32_t)(256.0 * (0.25 + 0.5 * u01(mix64(key ^ 0x3333ull))));
        // 16-byte blocks in CONTIG coordinates; stores are byte-wise at the ragged ends, 128-bit when aligned
        const int64_t n_blocks = (end - start + 15) >> 4;
        const bool aligned = ((start & 15) == 0);
        for (int64_t b = lane; b < n_blocks; b += 32) {
            uint8_t v[16];
            synth_block16(key, thr, (uint64_t)b, v);
            const int64_t p = start + (b << 4);
            if (aligned && p + 16 <= end) {
                uint4 q;
                memcpy(&q, v, 16);
                *reinterpret_cast<uint4 *>(seq + p) = q;
            } else {
                for (int i = 0; i < 16 && p + i < end; ++i) seq[p + i] = v[i];
            }
        }
    }
}

}  // namespace phm

using namespace phm;

extern "C" int phm_synth_lengths(uint64_t seed, int64_t first_contig, int64_t n_contigs, int64_t *d_lengths, void *stream) {
    PHM_REQUIRE(n_contigs >= 0, "n_contigs must be >= 0");
    if (n_contigs == 0) return PHM_OK;
    PHM_REQUIRE(d_lengths != nullptr, "null pointer");
    int64_t blocks = (n_contigs + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    synth_lengths_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(seed, first_contig, n_contigs, d_lengths);
    PHM_LAUNCH_CHECK();
    return PHM_OK;
}

extern "C" int phm_synth_bases(uint64_t seed, int64_t first_contig, int64_t n_contigs, const int64_t *d_offsets,
                               uint8_t *d_seq, void *stream) {
    PHM_REQUIRE(n_contigs >= 0, "n_contigs must be >= 0");
    if (n_contigs == 0) return PHM_OK;
    PHM_REQUIRE(d_offsets != nullptr && d_seq != nullptr, "null pointer");
    int64_t blocks = (n_contigs + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    synth_bases_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(seed, first_contig, n_contigs, d_offsets, d_seq);
    PHM_LAUNCH_CHECK();
    return PHM_OK;
}
