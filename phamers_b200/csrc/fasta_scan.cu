// FASTA ingest on the device: raw file bytes -> sequence bytes laid end to end + record offsets + header
// positions, i.e. the tokenisation kmer.count_file gets from Bio.SeqIO (reference scripts/kmer.py:131-139, scripts/fileIO.py:28-42)
// and that phamers_b200/fileIO.py::split_fasta_bytes restates on the host.
//
// Semantics (Bio.SeqIO FASTA parser as the reference uses it): a record starts at a line whose first byte is '>'; its sequence is
// every following line up to the next such line, with line feeds, carriage returns and blanks removed -- k-mers span line breaks but
// never records; text before the first '>' is ignored.  Tab / VT / FF are stripped by the reference only at line ends, which this
// scan does not model: it reports their presence (result[2]) and the caller takes the exact host path for such (rare) files.
//
// Every byte is in one of three line states: 0 = preamble (no header seen yet), 1 = inside a header line, 2 = inside a sequence line.
// A line start (first byte of the file, or the byte after '\n') moves the state: '>' -> 1; anything else: 1 -> 2, 0 and 2 stay.
// A byte is KEPT iff its state is 2 and it is none of '\n' '\r' ' '.  The state of a byte depends on everything before it, so the
// scan works on state MAPS (3 states -> 3 states, 6 bits), whose composition is associative:
//   pass 1  fasta_tile_kernel    per 4 KB tile: its map, the kept count for each incoming state, its number of header starts
//   pass 2  fasta_chain_kernel   one CTA: exclusive scan over the tiles -> incoming state, sequence offset and record index of each tile
//   pass 3  fasta_emit_kernel    per tile, state now known: compacts the kept bytes (staged in shared memory, coalesced stores),
//                                writes offsets[r] and header_pos[r] for the records that start in the tile
// Traffic: the file is read twice and the sequence written once (~3 bytes per input byte); pass 2 touches 32 bytes per tile.
#include "phm_common.cuh"

namespace phm {
namespace fasta {

constexpr int TILE_THREADS = 256;
constexpr int PER_THREAD = 16;
constexpr int TILE_BYTES = TILE_THREADS * PER_THREAD;      // 4096

struct TileSummary {               // 32 bytes
    uint32_t map;                  // incoming state s -> outgoing state, 2 bits each
    uint32_t n_headers;            // header starts in the tile
    uint32_t kept[3];              // kept bytes of the tile for incoming state 0 / 1 / 2
    uint32_t odd;                  // tab / VT / FF seen
    uint32_t pad[2];
};
struct TileStart {                 // 24 bytes, written by pass 2
    int64_t seq_base;              // kept bytes before the tile
    int64_t rec_base;              // header starts before the tile
    uint32_t state;                // incoming state
    uint32_t pad;
};

constexpr uint32_t MAP_ID = 0u | (1u << 2) | (2u << 4);
__device__ __forceinline__ uint32_t map_apply(uint32_t m, uint32_t s) { return (m >> (2 * s)) & 3u; }
__device__ __forceinline__ uint32_t map_then(uint32_t first, uint32_t second) {       // s -> second(first(s))
    return map_apply(second, map_apply(first, 0)) | (map_apply(second, map_apply(first, 1)) << 2) |
           (map_apply(second, map_apply(first, 2)) << 4);
}
__device__ __forceinline__ bool keepable(uint32_t c) { return c != 10u && c != 13u && c != 32u; }

// this thread's 16 bytes (bytes at or past n read as '\n': they are never kept and start no record) and the byte before them
__device__ __forceinline__ void load16(const uint8_t *raw, int64_t n, int64_t i0, uint32_t (&c)[PER_THREAD], uint32_t &prev) {
    uint4 v = make_uint4(0x0A0A0A0Au, 0x0A0A0A0Au, 0x0A0A0A0Au, 0x0A0A0A0Au);
    if (i0 < n) v = *reinterpret_cast<const uint4 *>(raw + i0);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < PER_THREAD; ++j) {
        c[j] = (w[j >> 2] >> (8 * (j & 3))) & 255u;
        if (i0 + j >= n) c[j] = 10u;
    }
    prev = (i0 > 0 && i0 <= n) ? raw[i0 - 1] : 10u;             // the first byte of the file starts a line
}

// ---------------- pass 1 ----------------
__global__ void __launch_bounds__(TILE_THREADS) fasta_tile_kernel(const uint8_t *__restrict__ raw, int64_t n, int64_t n_tiles,
                                                                  TileSummary *__restrict__ tiles) {
    __shared__ uint32_t s_map[TILE_THREADS / 32];
    __shared__ uint32_t s_sum[TILE_THREADS / 32][5];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t i0 = tile * TILE_BYTES + (int64_t)threadIdx.x * PER_THREAD;
        uint32_t c[PER_THREAD], prev;
        load16(raw, n, i0, c, prev);
        // thread-local: outgoing state and kept bytes for each incoming state
        uint32_t st[3] = {0u, 1u, 2u}, kept[3] = {0u, 0u, 0u}, heads = 0u, odd = 0u;
#pragma unroll
        for (int j = 0; j < PER_THREAD; ++j) {
            const bool ls = prev == 10u, hs = ls && c[j] == (uint32_t)'>';
            const bool keep = keepable(c[j]);
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                st[s] = hs ? 1u : ((ls && st[s] == 1u) ? 2u : st[s]);
                kept[s] += (st[s] == 2u && keep) ? 1u : 0u;
            }
            heads += hs ? 1u : 0u;
            odd |= (c[j] == 9u || c[j] == 11u || c[j] == 12u) ? 1u : 0u;
            prev = c[j];
        }
        const uint32_t mine = st[0] | (st[1] << 2) | (st[2] << 4);
        // exclusive scan of the maps over the tile: which state each thread starts in, for each state the tile may start in
        uint32_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl = map_then(up, incl);
        }
        uint32_t excl = __shfl_up_sync(FULL, incl, 1);
        if (lane == 0) excl = MAP_ID;
        __syncthreads();                                           // previous tile's shared memory is no longer read
        if (lane == 31) s_map[warp] = incl;
        __syncthreads();
        uint32_t before = MAP_ID;
        for (int w = 0; w < warp; ++w) before = map_then(before, s_map[w]);
        const uint32_t start = map_then(before, excl);             // tile's incoming state -> this thread's incoming state
        uint32_t k0 = kept[map_apply(start, 0)], k1 = kept[map_apply(start, 1)], k2 = kept[map_apply(start, 2)];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            k0 += __shfl_xor_sync(FULL, k0, o); k1 += __shfl_xor_sync(FULL, k1, o); k2 += __shfl_xor_sync(FULL, k2, o);
            heads += __shfl_xor_sync(FULL, heads, o); odd |= __shfl_xor_sync(FULL, odd, o);
        }
        if (lane == 0) { s_sum[warp][0] = k0; s_sum[warp][1] = k1; s_sum[warp][2] = k2; s_sum[warp][3] = heads; s_sum[warp][4] = odd; }
        __syncthreads();
        if (threadIdx.x == 0) {
            TileSummary t;
            uint32_t whole = MAP_ID;
            t.kept[0] = t.kept[1] = t.kept[2] = 0u; t.n_headers = 0u; t.odd = 0u; t.pad[0] = t.pad[1] = 0u;
            for (int w = 0; w < TILE_THREADS / 32; ++w) {
                whole = map_then(whole, s_map[w]);
                t.kept[0] += s_sum[w][0]; t.kept[1] += s_sum[w][1]; t.kept[2] += s_sum[w][2];
                t.n_headers += s_sum[w][3]; t.odd |= s_sum[w][4];
            }
            t.map = whole;
            tiles[tile] = t;
        }
    }
}

// ---------------- pass 2: one CTA chains the tiles ----------------
// result[0] = records, [1] = sequence bytes kept, [2] = 1 if a tab / VT / FF was seen, [3] = tiles
__global__ void __launch_bounds__(1024) fasta_chain_kernel(const TileSummary *__restrict__ tiles, int64_t n_tiles,
                                                           TileStart *__restrict__ starts, int64_t *__restrict__ result) {
    __shared__ uint32_t s_map[32];
    __shared__ unsigned long long s_a[32], s_b[32];
    __shared__ uint32_t s_odd;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t per = (n_tiles + 1023) / 1024;
    const int64_t lo = (int64_t)threadIdx.x * per, hi = (lo + per < n_tiles) ? lo + per : n_tiles;
    if (threadIdx.x == 0) s_odd = 0u;
    // sweep 1: this thread's chunk as one map
    uint32_t mine = MAP_ID, odd = 0u;
    for (int64_t t = lo; t < hi; ++t) { mine = map_then(mine, tiles[t].map); odd |= tiles[t].odd; }
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl = map_then(up, incl);
    }
    uint32_t excl = __shfl_up_sync(FULL, incl, 1);
    if (lane == 0) excl = MAP_ID;
    if (lane == 31) s_map[warp] = incl;
    __syncthreads();
    if (odd) atomicOr(&s_odd, 1u);
    uint32_t before = MAP_ID;
    for (int w = 0; w < warp; ++w) before = map_then(before, s_map[w]);
    const uint32_t state0 = map_apply(map_then(before, excl), 0u);       // the file starts in the preamble state
    // sweep 2: totals of the chunk given its real incoming state
    unsigned long long kept = 0, heads = 0;
    uint32_t st = state0;
    for (int64_t t = lo; t < hi; ++t) {
        const TileSummary s = tiles[t];
        kept += s.kept[st]; heads += s.n_headers;
        st = map_apply(s.map, st);
    }
    unsigned long long ki = kept, hi_ = heads;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long uk = __shfl_up_sync(FULL, ki, o), uh = __shfl_up_sync(FULL, hi_, o);
        if (lane >= o) { ki += uk; hi_ += uh; }
    }
    if (lane == 31) { s_a[warp] = ki; s_b[warp] = hi_; }
    __syncthreads();
    unsigned long long kb = ki - kept, hb = hi_ - heads;
    for (int w = 0; w < warp; ++w) { kb += s_a[w]; hb += s_b[w]; }
    // sweep 3: every tile's start
    st = state0;
    for (int64_t t = lo; t < hi; ++t) {
        const TileSummary s = tiles[t];
        TileStart out;
        out.seq_base = (int64_t)kb; out.rec_base = (int64_t)hb; out.state = st; out.pad = 0u;
        starts[t] = out;
        kb += s.kept[st]; hb += s.n_headers;
        st = map_apply(s.map, st);
    }
    if (threadIdx.x == 1023) {
        unsigned long long tk = 0, th = 0;
        for (int w = 0; w < 32; ++w) { tk += s_a[w]; th += s_b[w]; }
        result[0] = (int64_t)th; result[1] = (int64_t)tk; result[2] = (int64_t)s_odd; result[3] = n_tiles;
    }
}

// ---------------- pass 3 ----------------
__global__ void __launch_bounds__(TILE_THREADS) fasta_emit_kernel(const uint8_t *__restrict__ raw, int64_t n, int64_t n_tiles,
                                                                  const TileStart *__restrict__ starts, const int64_t *__restrict__ result,
                                                                  uint8_t *__restrict__ seq, int64_t *__restrict__ offsets,
                                                                  int64_t *__restrict__ header_pos, int64_t max_records) {
    __shared__ uint32_t s_map[TILE_THREADS / 32];
    __shared__ uint32_t s_k[TILE_THREADS / 32], s_h[TILE_THREADS / 32];
    __shared__ __align__(16) uint8_t s_bytes[TILE_BYTES];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (blockIdx.x == 0 && threadIdx.x == 0 && result[0] <= max_records) offsets[result[0]] = result[1];   // closing offset
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const TileStart ts = starts[tile];
        const int64_t i0 = tile * TILE_BYTES + (int64_t)threadIdx.x * PER_THREAD;
        uint32_t c[PER_THREAD], prev;
        load16(raw, n, i0, c, prev);
        // 1. this thread's map, to find the state it starts in
        uint32_t st[3] = {0u, 1u, 2u};
        uint32_t p = prev;
#pragma unroll
        for (int j = 0; j < PER_THREAD; ++j) {
            const bool ls = p == 10u, hs = ls && c[j] == (uint32_t)'>';
#pragma unroll
            for (int s = 0; s < 3; ++s) st[s] = hs ? 1u : ((ls && st[s] == 1u) ? 2u : st[s]);
            p = c[j];
        }
        uint32_t incl = st[0] | (st[1] << 2) | (st[2] << 4);
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl = map_then(up, incl);
        }
        uint32_t excl = __shfl_up_sync(FULL, incl, 1);
        if (lane == 0) excl = MAP_ID;
        __syncthreads();                                           // previous tile done with shared memory
        if (lane == 31) s_map[warp] = incl;
        __syncthreads();
        uint32_t before = MAP_ID;
        for (int w = 0; w < warp; ++w) before = map_then(before, s_map[w]);
        uint32_t state = map_apply(map_then(before, excl), ts.state);
        // 2. kept bytes and header starts of this thread, with the real state
        uint32_t keep_mask = 0u, head_mask = 0u;
        p = prev;
#pragma unroll
        for (int j = 0; j < PER_THREAD; ++j) {
            const bool ls = p == 10u, hs = ls && c[j] == (uint32_t)'>';
            state = hs ? 1u : ((ls && state == 1u) ? 2u : state);
            if (state == 2u && keepable(c[j])) keep_mask |= 1u << j;
            if (hs) head_mask |= 1u << j;
            p = c[j];
        }
        const uint32_t nk = __popc(keep_mask), nh = __popc(head_mask);
        uint32_t ki = nk, hi = nh;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t uk = __shfl_up_sync(FULL, ki, o), uh = __shfl_up_sync(FULL, hi, o);
            if (lane >= o) { ki += uk; hi += uh; }
        }
        if (lane == 31) { s_k[warp] = ki; s_h[warp] = hi; }
        __syncthreads();
        uint32_t kpos = ki - nk, hpos = hi - nh, tile_kept = 0u;
        for (int w = 0; w < TILE_THREADS / 32; ++w) {
            if (w < warp) { kpos += s_k[w]; hpos += s_h[w]; }
            tile_kept += s_k[w];
        }
        // 3. records that start here: offset = kept bytes before the header, position of its '>'
        uint32_t hm = head_mask;
        int64_t r = ts.rec_base + hpos;
        while (hm) {
            const int j = __ffs(hm) - 1;
            hm &= hm - 1u;
            if (r < max_records) {
                offsets[r] = ts.seq_base + kpos + __popc(keep_mask & ((1u << j) - 1u));
                header_pos[r] = i0 + j;
            }
            ++r;
        }
        // 4. compact into shared memory, then coalesced stores
        uint32_t km = keep_mask, at = kpos;
        while (km) {
            const int j = __ffs(km) - 1;
            km &= km - 1u;
            uint32_t v = c[0];
#pragma unroll
            for (int q = 1; q < PER_THREAD; ++q) v = (j == q) ? c[q] : v;
            s_bytes[at++] = (uint8_t)v;
        }
        __syncthreads();
        uint8_t *dst = seq + ts.seq_base;
        for (uint32_t b = threadIdx.x; b < tile_kept; b += TILE_THREADS) dst[b] = s_bytes[b];
    }
}

struct Workspace { TileSummary *tiles; TileStart *starts; size_t bytes; };
static Workspace carve(void *ws, int64_t n_bytes) {
    const int64_t n_tiles = (n_bytes + TILE_BYTES - 1) / TILE_BYTES;
    Workspace w;
    unsigned char *base = static_cast<unsigned char *>(ws);
    const size_t a = ((size_t)n_tiles * sizeof(TileSummary) + 255) & ~(size_t)255;
    const size_t b = ((size_t)n_tiles * sizeof(TileStart) + 255) & ~(size_t)255;
    w.tiles = reinterpret_cast<TileSummary *>(base);
    w.starts = reinterpret_cast<TileStart *>(base ? base + a : nullptr);
    w.bytes = a + b + 256;
    return w;
}

}  // namespace fasta
}  // namespace phm

using namespace phm;

extern "C" size_t phm_fasta_workspace_bytes(int64_t n_bytes) {
    return fasta::carve(nullptr, n_bytes < 0 ? 0 : n_bytes).bytes;
}

extern "C" int phm_fasta_index(const uint8_t *d_raw, int64_t n_bytes, int64_t *d_result, void *d_workspace, size_t workspace_bytes,
                               void *stream) {
    PHM_REQUIRE(n_bytes >= 0, "n_bytes must be >= 0");
    PHM_REQUIRE(d_result != nullptr && d_workspace != nullptr, "null pointer");
    PHM_REQUIRE(d_raw != nullptr || n_bytes == 0, "d_raw is null");
    PHM_REQUIRE((reinterpret_cast<uintptr_t>(d_raw) & 15u) == 0, "d_raw must be 16-byte aligned");
    fasta::Workspace w = fasta::carve(d_workspace, n_bytes);
    if (workspace_bytes < w.bytes) { set_error("workspace too small: %zu < %zu", workspace_bytes, w.bytes); return PHM_E_WORKSPACE; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t n_tiles = (n_bytes + fasta::TILE_BYTES - 1) / fasta::TILE_BYTES;
    if (n_tiles > 0) {
        int64_t grid = (int64_t)sm_count() * 8;
        if (grid > n_tiles) grid = n_tiles;
        fasta::fasta_tile_kernel<<<(unsigned)grid, fasta::TILE_THREADS, 0, st>>>(d_raw, n_bytes, n_tiles, w.tiles);
        PHM_CUDA_CHECK(cudaGetLastError());
    }
    fasta::fasta_chain_kernel<<<1, 1024, 0, st>>>(w.tiles, n_tiles, w.starts, d_result);
    PHM_CUDA_CHECK(cudaGetLastError());
    return PHM_OK;
}

extern "C" int phm_fasta_extract(const uint8_t *d_raw, int64_t n_bytes, const int64_t *d_result, uint8_t *d_seq, int64_t *d_offsets,
                                 int64_t *d_header_pos, int64_t max_records, const void *d_workspace, size_t workspace_bytes,
                                 void *stream) {
    PHM_REQUIRE(n_bytes >= 0 && max_records >= 0, "negative size");
    PHM_REQUIRE(d_result != nullptr && d_workspace != nullptr && d_offsets != nullptr, "null pointer");
    PHM_REQUIRE((d_raw != nullptr && d_seq != nullptr) || n_bytes == 0, "d_raw / d_seq is null");
    PHM_REQUIRE(d_header_pos != nullptr || max_records == 0, "d_header_pos is null");
    fasta::Workspace w = fasta::carve(const_cast<void *>(d_workspace), n_bytes);
    if (workspace_bytes < w.bytes) { set_error("workspace too small: %zu < %zu", workspace_bytes, w.bytes); return PHM_E_WORKSPACE; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t n_tiles = (n_bytes + fasta::TILE_BYTES - 1) / fasta::TILE_BYTES;
    int64_t grid = (int64_t)sm_count() * 8;
    if (grid > n_tiles) grid = n_tiles;
    if (grid < 1) grid = 1;
    fasta::fasta_emit_kernel<<<(unsigned)grid, fasta::TILE_THREADS, 0, st>>>(d_raw, n_bytes, n_tiles, w.starts, d_result, d_seq, d_offsets,
                                                                             d_header_pos, max_records);
    PHM_CUDA_CHECK(cudaGetLastError());
    return PHM_OK;
}
