// K4 + K5 on the 5th-generation tensor cores: contig x reference distance contraction with a fused shortlist epilogue,
// followed by an exact float64 re-rank.
//
// Replaces learning.knn (reference scripts/learning.py:118-128), the per-contig nearest-centroid loop
// (scripts/phamer.py:251-255 -> learning.closest_to :59-66) and proximity_metric (scripts/phamer.py:198-210).
//
//   d2(a, b) = |a|^2 + |b|^2 - 2 a.b          a = contig features [n, 256], b = reference rows / centroids
//
// The a.b term is a dense contraction and runs as tcgen05.mma (kind::f16, FP32 accumulators in tensor memory).  A single
// FP16 pass is not accurate enough to rank neighbours (nearest-neighbour d2 ~ 2e-5 against |x|^2 ~ 5e-3),
// so every operand is split into two FP16 terms of a 2^12-scaled value (x = hi + lo, 22 significant bits) and the product is
// accumulated as hi.hi + hi.lo + lo.hi in the SAME accumulator (the dropped lo.lo term is 2^-22 relative).  With FP32
// accumulation over 3 x 256 products the error of the ranking value is below 2^-15 (|a|^2 + |b|^2) (measured 2^-17.3, see
// rerank_kernel); it is used for SHORTLIST SELECTION ONLY.  The epilogue keeps, straight out of tensor memory, the 8 best
// references and the 4 best centroids of each class per contig; rerank_kernel then decides the vote from the error band
// around the k-th ranking value, re-measures exactly (float64, direct difference) whatever the band leaves open plus the
// nearest centroids, and evaluates tanh((e_n - e_p)/(e_p + e_n)).  Rows whose band reaches the end of the shortlist are
// appended to a list and re-scored by the exhaustive float64 kernel (score_exact.cu).
//
// Kernel layout (one CTA per SM, persistent over 128-contig tiles, 6 warps):
//   warp 0    TMA producer   A tile (128 contigs x 256 features, hi + lo, 128 KB, SWIZZLE_128B) once per contig tile;
//                            B stages (128 references x 64 features, hi + lo, 32 KB) through a 3-deep mbarrier ring
//   warp 1    MMA issuer     one elected lane: 12 tcgen05.mma (128x128x16) per stage, 48 per reference tile, into one of two
//                            128-column accumulators; tcgen05.commit frees the stage / publishes the accumulator
//   warps 2-5 epilogue       tcgen05.ld 32 columns at a time, v = |b|^2 * 2^23 - acc, branch-free sorted insertion into the
//                            per-thread shortlists (thread = contig row = tensor-memory lane)
#include <cuda.h>
#include <cuda_fp16.h>
#include <math.h>

#include "score_common.cuh"

namespace phm {

int score_collect_stats = 0;      // option "score_stats": re-measure every candidate and record the ranking error

namespace tc {

constexpr int KDIM = 256;                 // feature width handled by this kernel (k = 4)
constexpr int BM = 128, BN = 128, BK = 64;
constexpr int NKC = KDIM / BK;            // 4 K-chunks of 128 bytes
constexpr int NSTAGE = 4;                 // ring of single operand blocks: hi(kc), lo(kc), hi(kc+1), ...
constexpr int BLOCK_BYTES = BM * BK * 2;  // 16 KB: 128 rows x 128 B
constexpr int A_BYTES = 2 * NKC * BLOCK_BYTES;          // hi + lo
constexpr int NGROUP = 2;                 // epilogue warp groups (4 warps each); group g owns reference tiles with tile % 2 == g
constexpr int STACK_BYTES = NGROUP * 4 * 32 * 32 * 4;   // per epilogue warp: 32 ranking values of each of its 32 rows
constexpr int SMEM_EXTRA = 2048;
constexpr int SMEM_BYTES = A_BYTES + NSTAGE * BLOCK_BYTES + STACK_BYTES + SMEM_EXTRA + 1024;   // + alignment slack
constexpr int NTHREADS = 64 + NGROUP * 128;
constexpr int TMEM_COLS = 256;            // two 128-column FP32 accumulators
constexpr int LREF = 8;                   // shortlist of references per contig and epilogue group
constexpr int LCEN = 4;                   // shortlist of centroids per class
constexpr int NSLOT = LREF + 2 * LCEN;    // 16 candidate slots per group
constexpr int NCAND = NGROUP * NSLOT;     // 32 per contig
constexpr float SCALE = 4096.0f;          // operands are scaled by 2^12 -> accumulator = 2^24 a.b
constexpr float NORM_SCALE = 8388608.0f;  // 2^23: v = 2^23 (|b|^2 - 2 a.b)
constexpr float PAD_NORM = 3.0e38f;
constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);   // F16 x F16 -> F32, K-major

// ---------------- PTX helpers ----------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) __trap();        // a protocol bug must fail loudly, never hang the GPU
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int x, int y) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(IDESC), "r"(accumulate), "r"(0u) : "memory");
}
// K-major, SWIZZLE_128B operand tile: rows of 128 bytes, 8-row swizzle atoms 1024 bytes apart
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;                    // leading byte offset (unused for swizzled K-major), 16-byte units
    d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                    // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// branch-free sorted insertion (ascending, ties keep the earlier entry)
template <int L>
__device__ __forceinline__ void shortlist_insert(float (&w)[L], int (&id)[L], float v, int j) {
#pragma unroll
    for (int i = L - 1; i > 0; --i) {
        const bool shift = v < w[i - 1];
        const bool here = v < w[i];
        id[i] = shift ? id[i - 1] : (here ? j : id[i]);
        w[i] = shift ? w[i - 1] : (here ? v : w[i]);
    }
    const bool first = v < w[0];
    id[0] = first ? j : id[0];
    w[0] = first ? v : w[0];
}

// One reference tile (128 columns) for this thread's contig row, CW columns at a time: the ranking values are formed
// branch-free (tcgen05.ld, |b|^2 - acc), parked in shared memory, and the sign bits of (value - threshold) are funnelled
// into a per-thread column mask.  Only then does the thread walk its set bits and merge those values into the sorted
// shortlist, so the divergent ~35-instruction insertion runs max-over-lanes(popcount) times per chunk instead of once per
// candidate column, and everything before it is straight-line code with instruction-level parallelism.
template <int L>
__device__ __forceinline__ void scan_tile(uint32_t taddr, uint32_t nbs_addr, uint32_t vbuf_addr, int col0,
                                          float (&w)[L], int (&id)[L]) {
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
        float v[32];
        __syncwarp();                                  // tcgen05.ld is .sync.aligned: the warp must be converged
        tmem_ld32(taddr + (uint32_t)c, v);
        const float limit = w[L - 1];
        uint32_t h[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            const uint4 nb = lds_v4(nbs_addr + 4u * (c + 4 * g));
            float4 x;
            x.x = __uint_as_float(nb.x) - v[4 * g];
            x.y = __uint_as_float(nb.y) - v[4 * g + 1];
            x.z = __uint_as_float(nb.z) - v[4 * g + 2];
            x.w = __uint_as_float(nb.w) - v[4 * g + 3];
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(vbuf_addr + 512u * g), "f"(x.x), "f"(x.y), "f"(x.z), "f"(x.w) : "memory");
            // shift the sign of (value - limit) into the mask: first column of the chunk ends up in the top bit
            uint32_t &m = h[g >> 1];
            m = __funnelshift_l(__float_as_uint(x.x - limit), m, 1);
            m = __funnelshift_l(__float_as_uint(x.y - limit), m, 1);
            m = __funnelshift_l(__float_as_uint(x.z - limit), m, 1);
            m = __funnelshift_l(__float_as_uint(x.w - limit), m, 1);
        }
        uint32_t hits = (h[0] << 24) | (h[1] << 16) | (h[2] << 8) | h[3];
        const int rounds = __reduce_max_sync(FULL, (unsigned)__popc(hits));
        for (int r = 0; r < rounds; ++r) {
            if (hits) {
                const int i = __clz(hits);
                hits &= ~(0x80000000u >> i);
                const float x = __uint_as_float(lds_u32(vbuf_addr + 512u * (i >> 2) + 4u * (i & 3)));
                if (x < w[L - 1]) shortlist_insert<L>(w, id, x, col0 + c + i);
            }
        }
    }
}

struct TcParams {
    int64_t n_points;
    int n_mtiles;
    int nt_ref, nt_pos, nt_neg;          // reference tiles of each class (each class padded to a multiple of 128 rows)
    const float *nbs;                    // [ (nt_ref + nt_pos + nt_neg) * 128 ] scaled squared norms, PAD_NORM on padding rows
    int *cand;                           // [n_points, 16] candidate indices within their class, -1 = none
    float *approx;                       // [n_points, 16] ranking values v = 2^23 (|b|^2 - 2 a.b) of the candidates
};

__global__ void __launch_bounds__(NTHREADS, 1)
score_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo, TcParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sm_a = smem_base;                                   // [hi chunk 0..3][lo chunk 0..3]
    const uint32_t sm_b = smem_base + A_BYTES;                         // NSTAGE operand blocks
    const uint32_t sm_stack = sm_b + NSTAGE * BLOCK_BYTES;             // per epilogue warp: [8 column quads][32 lanes] x 16 bytes
    const uint32_t sm_x = sm_stack + STACK_BYTES;                      // barriers, tmem pointer, norm staging
    const uint32_t bar_a_full = sm_x + 0, bar_a_empty = sm_x + 8;
    const uint32_t bar_b_full = sm_x + 16, bar_b_empty = sm_x + 16 + 8 * NSTAGE;
    const uint32_t bar_t_full = sm_x + 16 + 16 * NSTAGE, bar_t_empty = bar_t_full + 16;
    const uint32_t tmem_slot = bar_t_empty + 16;
    const uint32_t sm_nbs = sm_x + 512;                                // [NGROUP][128] floats
    unsigned char *generic_x = smem_raw + (sm_x - smem_u32(smem_raw));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int nt_total = p.nt_ref + p.nt_pos + p.nt_neg;

    if (threadIdx.x == 0) {
        mbar_init(bar_a_full, 1);
        mbar_init(bar_a_empty, 1);
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(bar_b_full + 8 * s, 1); mbar_init(bar_b_empty + 8 * s, 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(bar_t_full + 8 * b, 1); mbar_init(bar_t_empty + 8 * b, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(generic_x + (tmem_slot - sm_x));

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            uint32_t bstage = 0, bphase = 0;
            int it = 0;
            for (int mt = blockIdx.x; mt < p.n_mtiles; mt += gridDim.x, ++it) {
                mbar_wait(bar_a_empty, (uint32_t)((it & 1) ^ 1));
                mbar_expect_tx(bar_a_full, A_BYTES);
                for (int kc = 0; kc < NKC; ++kc) {
                    tma_load_2d(sm_a + kc * BLOCK_BYTES, &map_a_hi, bar_a_full, kc * BK, mt * BM);
                    tma_load_2d(sm_a + (NKC + kc) * BLOCK_BYTES, &map_a_lo, bar_a_full, kc * BK, mt * BM);
                }
                for (int nt = 0; nt < nt_total; ++nt) {
                    for (int blk = 0; blk < 2 * NKC; ++blk) {           // hi(0), lo(0), hi(1), lo(1), ...
                        mbar_wait(bar_b_empty + 8 * bstage, bphase ^ 1u);
                        mbar_expect_tx(bar_b_full + 8 * bstage, BLOCK_BYTES);
                        tma_load_2d(sm_b + bstage * BLOCK_BYTES, (blk & 1) ? &map_b_lo : &map_b_hi, bar_b_full + 8 * bstage,
                                    (blk >> 1) * BK, nt * BN);
                        if (++bstage == NSTAGE) { bstage = 0; bphase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            uint32_t bstage = 0, bphase = 0, tile = 0;
            int it = 0;
            for (int mt = blockIdx.x; mt < p.n_mtiles; mt += gridDim.x, ++it) {
                mbar_wait(bar_a_full, (uint32_t)(it & 1));
                for (int nt = 0; nt < nt_total; ++nt, ++tile) {
                    const uint32_t buf = tile & 1u;
                    mbar_wait(bar_t_empty + 8 * buf, ((tile >> 1) & 1u) ^ 1u);
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + buf * BN;
                    for (int blk = 0; blk < 2 * NKC; ++blk) {
                        const int kc = blk >> 1;
                        mbar_wait(bar_b_full + 8 * bstage, bphase);
                        tc_fence_after();
                        const uint32_t a_hi = sm_a + kc * BLOCK_BYTES, a_lo = sm_a + (NKC + kc) * BLOCK_BYTES;
                        const uint32_t b = sm_b + bstage * BLOCK_BYTES;
#pragma unroll
                        for (int ks = 0; ks < BK / 16; ++ks) {
                            const uint32_t o = ks * 32;          // 16 halves = 32 bytes along K inside the swizzle row
                            if ((blk & 1) == 0) {                // b = hi block: hi.hi and lo.hi
                                umma_f16(tmem_d, smem_desc(a_hi + o), smem_desc(b + o), (blk | ks) ? 1u : 0u);
                                umma_f16(tmem_d, smem_desc(a_lo + o), smem_desc(b + o), 1u);
                            } else {                             // b = lo block: hi.lo
                                umma_f16(tmem_d, smem_desc(a_hi + o), smem_desc(b + o), 1u);
                            }
                        }
                        umma_commit(bar_b_empty + 8 * bstage);    // block reusable once these MMAs have read it
                        if (++bstage == NSTAGE) { bstage = 0; bphase ^= 1u; }
                    }
                    umma_commit(bar_t_full + 8 * buf);            // accumulator complete
                }
                umma_commit(bar_a_empty);                          // A tile no longer read
            }
        }
    } else {
        // ================= epilogue: 2 groups x 4 warps; thread = contig row = tensor-memory lane =================
        const int q = warp & 3;                                    // tensor-memory lane quarter this warp may read
        const int group = (warp - 2) >> 2;                         // owns reference tiles with (tile & 1) == group
        const int epi_tid = ((warp - 2) & 3) * 32 + lane;          // 0..127 inside the group
        const uint32_t stack_addr = sm_stack + (uint32_t)(warp - 2) * 4096u + 16u * lane;
        const uint32_t nbs_addr = sm_nbs + (uint32_t)group * (BN * 4);
        for (int mt = blockIdx.x, it = 0; mt < p.n_mtiles; mt += gridDim.x, ++it) {
            float wr[LREF], wp[LCEN], wn[LCEN];
            int ir[LREF], ip[LCEN], in_[LCEN];
#pragma unroll
            for (int i = 0; i < LREF; ++i) { wr[i] = INFINITY; ir[i] = -1; }
#pragma unroll
            for (int i = 0; i < LCEN; ++i) { wp[i] = INFINITY; ip[i] = -1; wn[i] = INFINITY; in_[i] = -1; }

            const uint32_t tile0 = (uint32_t)it * (uint32_t)nt_total;      // running tile index of this CTA
            for (int nt = 0; nt < nt_total; ++nt) {
                const uint32_t tile = tile0 + (uint32_t)nt;
                if ((int)(tile & 1u) != group) continue;
                const uint32_t buf = tile & 1u;
                asm volatile("bar.sync %0, 128;" ::"r"(1 + group) : "memory");       // previous tile's norm reads are done
                sts_u32(nbs_addr + 4u * epi_tid, __float_as_uint(p.nbs[(int64_t)nt * BN + epi_tid]));
                mbar_wait(bar_t_full + 8 * buf, (tile >> 1) & 1u);
                tc_fence_after();
                asm volatile("bar.sync %0, 128;" ::"r"(1 + group) : "memory");       // norm staging visible to the group
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN;
                if (nt < p.nt_ref) scan_tile<LREF>(taddr, nbs_addr, stack_addr, nt * BN, wr, ir);
                else if (nt < p.nt_ref + p.nt_pos) scan_tile<LCEN>(taddr, nbs_addr, stack_addr, (nt - p.nt_ref) * BN, wp, ip);
                else scan_tile<LCEN>(taddr, nbs_addr, stack_addr, (nt - p.nt_ref - p.nt_pos) * BN, wn, in_);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_t_empty + 8 * buf);
            }
            const int64_t row = (int64_t)mt * BM + q * 32 + lane;
            if (row < p.n_points) {
                int *c = p.cand + row * NCAND + group * NSLOT;
                float *a = p.approx + row * NCAND + group * NSLOT;
#pragma unroll
                for (int i = 0; i < LREF; ++i) { c[i] = ir[i]; a[i] = wr[i]; }
#pragma unroll
                for (int i = 0; i < LCEN; ++i) { c[LREF + i] = ip[i]; a[LREF + i] = wp[i]; c[LREF + LCEN + i] = in_[i]; a[LREF + LCEN + i] = wn[i]; }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}

// ---------------- operand preparation: float64 rows -> centred, 2^12-scaled FP16 hi / lo ----------------
// Distances are translation invariant, so every row is shifted by the uniform vector 1/256 before the split: frequency
// rows sum to 1, which makes |x - u|^2 = |x|^2 - 1/256 about 4.5x smaller than |x|^2 and shrinks the absolute error of the
// FP32-accumulated ranking value by the same factor.  (Exact distances are always formed from the unshifted float64 rows.)
__device__ __forceinline__ void split_store(double x, __half *hi, __half *lo, int64_t o) {
    const double s = x * (double)SCALE;
    const __half h = __float2half_rn((float)s);
    const __half l = __float2half_rn((float)(s - (double)__half2float(h)));
    hi[o] = h;
    lo[o] = l;
}

// rows [0, n_src) from src, rows [n_src, n_rows) zero padding; squared norms of the raw rows (float64, for the exact
// kernel), of the centred rows (float64, for the error bound) and of the centred rows scaled to ranking units (FP32)
//
// Reference rows are laid out in a scrambled order, destination row r <- source row (perm_a * r + perm_c) mod n_src with
// perm_a ~ 0.618 n_src coprime to n_src.  The shipped tables are sorted by taxonomy, so distances to a contig run in long
// monotone stretches along the file and the running-shortlist threshold of the epilogue would be beaten far more often
// than the 8/j of an exchangeable order; a golden-ratio stride makes every prefix an even sample of the whole file.
__global__ void tc_prep_rows_kernel(const double *__restrict__ src, int64_t n_src, int64_t n_rows, int64_t perm_a,
                                    int64_t perm_c, __half *__restrict__ hi, __half *__restrict__ lo,
                                    double *__restrict__ norm64, double *__restrict__ cnorm64, float *__restrict__ nbs) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const double shift = 1.0 / (double)KDIM;
    for (int64_t r = warp; r < n_rows; r += n_warps) {
        double s = 0.0, sc = 0.0;
        const int64_t sr = (r < n_src) ? (perm_a * r + perm_c) % n_src : 0;
        for (int d = lane; d < KDIM; d += 32) {
            const double x = (r < n_src) ? src[sr * KDIM + d] : shift;
            const double xc = x - shift;
            split_store(xc, hi, lo, r * KDIM + d);
            s = fma(x, x, s);
            sc = fma(xc, xc, sc);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s += __shfl_xor_sync(FULL, s, o);
            sc += __shfl_xor_sync(FULL, sc, o);
        }
        if (lane == 0) {
            if (norm64 && r < n_src) norm64[sr] = s;
            if (cnorm64 && r < n_src) cnorm64[sr] = sc;
            if (nbs) nbs[r] = (r < n_src) ? (float)(sc * (double)NORM_SCALE) : PAD_NORM;
        }
    }
}

// ---------------- decision + exact re-measurement of the shortlists ----------------
// The ranking value v_j = 2^23 (|b_j|^2 - 2 a.b_j) (centred rows) of reference j differs from the exact
// 2^23 (d2_j - |a|^2) by at most Ev = 2^23 * RANK_EPS * (|a|^2 + |b_j|^2): tcgen05 accumulates 48 MMAs per value in FP32 with
// truncation (<= 48 * 2^-23 = 2^-17.4 relative to the sum of the magnitudes), the split drops 2^-22; measured maximum on
// 1.3e6 (contig, reference) pairs: 2^-17.3.  RANK_EPS = 2^-15 leaves a factor 5.  For any reference that could still matter
// |b|^2 <= 2 |a|^2 + 2 d2, so one bound per row is used: E = RANK_EPS * (3 |a|^2 + 2 d2_last).
//
// k nearest neighbours: every reference outside the band {v <= v_k + 2 Ev} is provably farther than the k-th nearest.
// If the band reaches the end of the shortlist, unseen references may belong to it: the row goes to the exhaustive kernel.
// Otherwise the vote is read off the band when it has exactly k members or a single label; only a mixed band is
// re-measured exactly (float64, direct difference).  Centroids: the band around the best ranking value is re-measured
// exactly and the nearest taken; its exact distance enters the score.
constexpr double RANK_EPS = 1.0 / 32768.0;      // 2^-15

struct RerankParams {
    const double *points; int64_t n_points;
    const double *refs; int64_t n_refs; int64_t n_positive;
    int64_t perm_a, perm_c;            // reference candidate index -> original row: (perm_a * idx + perm_c) mod n_refs
    const double *cent_pos; int64_t n_cent_pos;
    const double *cent_neg; int64_t n_cent_neg;
    const double *cnorm_points;        // centred squared norms of the query rows
    const double *cnorm_refs;
    const int *cand; const float *approx;
    int k_neighbors;
    double *knn, *kmeans, *combo;
    int64_t *fallback_rows; unsigned long long *fallback_count;
    float *max_rank_error;             // when non-null: re-measure every reference candidate and record the ranking error
    unsigned long long *rows_remeasured;
};

__device__ __forceinline__ double warp_exact_d2(const double (&x)[KDIM / 32], const double *__restrict__ b, int lane) {
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < KDIM / 32; ++i) {
        const double t = x[i] - b[lane + 32 * i];
        acc = fma(t, t, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
    return acc;
}

__device__ __forceinline__ float warp_min_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}

// One warp per contig; lane l holds candidate l: epilogue group l / 16, slot l % 16 (0..7 references, 8..11 positive
// centroids, 12..15 negative centroids).  Each group lists the best candidates among ITS reference tiles, so the union of
// the two lists contains the global best and every unlisted reference of group g ranks behind slot 7 of group g.
__global__ void __launch_bounds__(256) rerank_kernel(RerankParams p) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const double to_d2 = 1.0 / (double)NORM_SCALE;
    const int kn = p.k_neighbors;
    const int slot = lane & (NSLOT - 1);
    const bool ref_lane = slot < LREF;
    for (int64_t row = warp; row < p.n_points; row += n_warps) {
        const double *pt = p.points + row * KDIM;
        double x[KDIM / 32];
#pragma unroll
        for (int i = 0; i < KDIM / 32; ++i) x[i] = pt[lane + 32 * i];
        const double na = p.cnorm_points[row];
        int my_cand = p.cand[row * NCAND + lane];
        const float my_v = (my_cand >= 0) ? p.approx[row * NCAND + lane] : INFINITY;
        if (ref_lane && my_cand >= 0) my_cand = (int)((p.perm_a * my_cand + p.perm_c) % p.n_refs);   // back to file order
        bool fallback = false;
        double knn = NAN, km = NAN;

        if (!isnan(na)) {
            // ---------------- k nearest references ----------------
            {
                const bool valid = ref_lane && my_cand >= 0;
                // every unlisted reference ranks at or behind the last slot of its own group's (full) list
                const float v_lim = fminf(__shfl_sync(FULL, my_v, LREF - 1), __shfl_sync(FULL, my_v, NSLOT + LREF - 1));
                const float v_max = warp_max_f(valid ? my_v : -INFINITY);
                const double d_bound = fmax(na + (double)(isinf(v_lim) ? v_max : v_lim) * to_d2, 0.0);
                const float ev2 = (float)(2.0 * RANK_EPS * (3.0 * na + 2.0 * d_bound) * (double)NORM_SCALE);
                int rank = 0;                                         // position of my candidate in the merged order
#pragma unroll
                for (int g = 0; g < NGROUP; ++g)
#pragma unroll
                    for (int sl = 0; sl < LREF; ++sl) {
                        const int src = g * NSLOT + sl;
                        const float ov = __shfl_sync(FULL, my_v, src);
                        const int oi = __shfl_sync(FULL, my_cand, src);
                        if (oi >= 0 && src != lane && (ov < my_v || (ov == my_v && oi < my_cand))) ++rank;
                    }
                const unsigned kth = __ballot_sync(FULL, valid && rank == kn - 1);     // exists: k_neighbors <= n_refs
                const float v_k = __shfl_sync(FULL, my_v, __ffs(kth) - 1);
                const bool in_band = valid && my_v <= v_k + ev2;
                const unsigned band = __ballot_sync(FULL, in_band);
                const unsigned pos_mask = __ballot_sync(FULL, in_band && my_cand < p.n_positive);
                const bool complete = isinf(v_lim) || (v_lim > v_k + ev2);
                const int n_band = __popc(band);
                if (!complete || kth == 0u) {
                    fallback = true;
                } else if (n_band == kn || pos_mask == 0u || pos_mask == band) {
                    // the k nearest are exactly the band, or every possible member votes the same way
                    const int pos = (pos_mask == band) ? kn : ((pos_mask == 0u) ? 0 : __popc(pos_mask));
                    knn = (2 * pos > kn) ? 1.0 : -1.0;               // 2 * (predict - 0.5), scripts/learning.py:128
                } else {
                    // mixed band: exact distances of its members, k smallest (ties: lower reference index first)
                    double my_d2 = INFINITY;
                    unsigned rest = band;
                    while (rest) {
                        const int s = __ffs(rest) - 1;
                        rest &= rest - 1;
                        const int idx = __shfl_sync(FULL, my_cand, s);
                        const double d = warp_exact_d2(x, p.refs + (int64_t)idx * KDIM, lane);
                        if (lane == s) my_d2 = d;
                    }
                    int erank = 0;
                    rest = band;
                    while (rest) {
                        const int s = __ffs(rest) - 1;
                        rest &= rest - 1;
                        const double od = __shfl_sync(FULL, my_d2, s);
                        const int oi = __shfl_sync(FULL, my_cand, s);
                        if (s != lane && (od < my_d2 || (od == my_d2 && oi < my_cand))) ++erank;
                    }
                    const unsigned top = __ballot_sync(FULL, in_band && erank < kn);
                    const int pos = __popc(top & pos_mask);
                    knn = (2 * pos > kn) ? 1.0 : -1.0;
                    if (p.rows_remeasured && lane == 0) atomicAdd(p.rows_remeasured, 1ull);
                }
                if (p.max_rank_error) {           // diagnostics: ranking error of every reference candidate
                    float worst = 0.f, worst_rel = 0.f;
                    for (int s = 0; s < NCAND; ++s) {
                        const int idx = __shfl_sync(FULL, my_cand, s);
                        const float vs = __shfl_sync(FULL, my_v, s);
                        if (idx < 0 || (s & (NSLOT - 1)) >= LREF) continue;
                        const double d = warp_exact_d2(x, p.refs + (int64_t)idx * KDIM, lane);
                        const double err = fabs(na + (double)vs * to_d2 - d);
                        worst = fmaxf(worst, (float)err);
                        worst_rel = fmaxf(worst_rel, (float)(err / (na + p.cnorm_refs[idx])));
                    }
                    if (lane == 0) {
                        atomicMax(reinterpret_cast<int *>(p.max_rank_error), __float_as_int(worst));
                        atomicMax(reinterpret_cast<int *>(p.max_rank_error + 1), __float_as_int(worst_rel));
                    }
                }
            }
            // ---------------- nearest centroid of each class ----------------
            if (p.n_cent_pos > 0 && p.n_cent_neg > 0) {
                double e2[2];
#pragma unroll
                for (int cls = 0; cls < 2; ++cls) {
                    const int base = LREF + cls * LCEN;
                    const bool mine = slot >= base && slot < base + LCEN && my_cand >= 0;
                    const float v_0 = warp_min_f(mine ? my_v : INFINITY);
                    const float v_lim = fminf(__shfl_sync(FULL, my_v, base + LCEN - 1), __shfl_sync(FULL, my_v, NSLOT + base + LCEN - 1));
                    const float v_max = warp_max_f(mine ? my_v : -INFINITY);
                    const double d_bound = fmax(na + (double)(isinf(v_lim) ? v_max : v_lim) * to_d2, 0.0);
                    const float ev2 = (float)(2.0 * RANK_EPS * (3.0 * na + 2.0 * d_bound) * (double)NORM_SCALE);
                    const double *cents = cls ? p.cent_neg : p.cent_pos;
                    unsigned rest = __ballot_sync(FULL, mine && my_v <= v_0 + ev2);
                    if (!(isinf(v_lim) || v_lim > v_0 + ev2)) fallback = true;
                    double best = INFINITY;
                    while (rest) {
                        const int s = __ffs(rest) - 1;
                        rest &= rest - 1;
                        const int idx = __shfl_sync(FULL, my_cand, s);
                        best = fmin(best, warp_exact_d2(x, cents + (int64_t)idx * KDIM, lane));
                    }
                    e2[cls] = best;
                }
                const double e_pos = sqrt(e2[0]), e_neg = sqrt(e2[1]);
                km = tanh((e_neg - e_pos) / (e_pos + e_neg));          // scripts/phamer.py:206-209
            }
        }
        if (lane == 0) {
            if (fallback) {
                const unsigned long long slot_out = atomicAdd(p.fallback_count, 1ull);
                p.fallback_rows[slot_out] = row;
            } else {
                if (p.knn) p.knn[row] = knn;
                if (p.kmeans) p.kmeans[row] = km;
                if (p.combo) p.combo[row] = knn + km;                   // scripts/phamer.py:313
            }
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// [rows, 256] FP16 row-major, box = 64 features x 128 rows, 128-byte swizzle (the UMMA K-major canonical layout)
static int make_map(CUtensorMap *map, const void *base, int64_t rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return PHM_E_UNSUPPORTED; }
    cuuint64_t dims[2] = {(cuuint64_t)KDIM, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)KDIM * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
    cuuint32_t elem[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(base), dims, strides, box, elem,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with %d", (int)r); return PHM_E_CUDA; }
    return PHM_OK;
}

static inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct TcWorkspace {
    __half *a_hi, *a_lo, *b_hi, *b_lo;
    float *nbs;
    int *cand; float *approx;
    double *norm_points, *norm_refs, *norm_cpos, *norm_cneg;
    double *cnorm_points, *cnorm_refs;
    int64_t *fallback_rows; unsigned long long *fallback_count; float *max_rank_error;
    unsigned long long *rows_remeasured;
    size_t bytes;
};

static TcWorkspace carve_tc(void *ws, int64_t n, int64_t r_pad, int64_t n_refs, int64_t n_cp, int64_t n_cn) {
    TcWorkspace w;
    size_t off = 0;
    unsigned char *base = static_cast<unsigned char *>(ws);
    auto take = [&](size_t bytes) { unsigned char *p = base ? base + off : nullptr; off += align256(bytes); return p; };
    w.fallback_count = reinterpret_cast<unsigned long long *>(take(256));
    w.max_rank_error = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(w.fallback_count) + 64);
    w.rows_remeasured = reinterpret_cast<unsigned long long *>(reinterpret_cast<unsigned char *>(w.fallback_count) + 128);
    w.a_hi = reinterpret_cast<__half *>(take((size_t)n * KDIM * 2));
    w.a_lo = reinterpret_cast<__half *>(take((size_t)n * KDIM * 2));
    w.b_hi = reinterpret_cast<__half *>(take((size_t)r_pad * KDIM * 2));
    w.b_lo = reinterpret_cast<__half *>(take((size_t)r_pad * KDIM * 2));
    w.nbs = reinterpret_cast<float *>(take((size_t)r_pad * 4));
    w.cand = reinterpret_cast<int *>(take((size_t)n * NCAND * 4));
    w.approx = reinterpret_cast<float *>(take((size_t)n * NCAND * 4));
    w.norm_points = reinterpret_cast<double *>(take((size_t)n * 8));
    w.norm_refs = reinterpret_cast<double *>(take((size_t)n_refs * 8));
    w.norm_cpos = reinterpret_cast<double *>(take((size_t)n_cp * 8));
    w.norm_cneg = reinterpret_cast<double *>(take((size_t)n_cn * 8));
    w.cnorm_points = reinterpret_cast<double *>(take((size_t)n * 8));
    w.cnorm_refs = reinterpret_cast<double *>(take((size_t)n_refs * 8));
    w.fallback_rows = reinterpret_cast<int64_t *>(take((size_t)n * 8));
    w.bytes = off;
    return w;
}

size_t score_tc_workspace_bytes(int64_t n, int64_t n_refs, int64_t n_cp, int64_t n_cn) {
    const int64_t r_pad = round_up(n_refs, BN) + round_up(n_cp, BN) + round_up(n_cn, BN);
    return carve_tc(nullptr, n, r_pad, n_refs, n_cp, n_cn).bytes;
}

bool score_tc_supported(int dim, int k_neighbors, int64_t n_cp, int64_t n_cn) {
    return dim == KDIM && k_neighbors <= 5 && n_cp > 0 && n_cn > 0;
}

static int launch_prep(const double *src, int64_t n_src, int64_t n_rows, int64_t perm_a, int64_t perm_c, __half *hi, __half *lo,
                       double *norm64, double *cnorm64, float *nbs, cudaStream_t st) {
    if (n_rows == 0) return PHM_OK;
    int64_t blocks = (n_rows + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    tc_prep_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>(src, n_src, n_rows, perm_a, perm_c, hi, lo, norm64, cnorm64, nbs);
    PHM_CUDA_CHECK(cudaGetLastError());
    return PHM_OK;
}

int score_tc(const ScoreArgs &a, void *ws, size_t ws_bytes, cudaStream_t st, int *kernels_launched) {
    const int64_t n = a.n_points;
    const int64_t ref_pad = round_up(a.n_refs, BN), cp_pad = round_up(a.n_cent_pos, BN), cn_pad = round_up(a.n_cent_neg, BN);
    const int64_t r_pad = ref_pad + cp_pad + cn_pad;
    TcWorkspace w = carve_tc(ws, n, r_pad, a.n_refs, a.n_cent_pos, a.n_cent_neg);
    if (ws_bytes < w.bytes) { set_error("workspace too small: %zu < %zu", ws_bytes, w.bytes); return PHM_E_WORKSPACE; }
    PHM_REQUIRE(n < ((int64_t)1 << 31) * BM / 2 && r_pad < ((int64_t)1 << 31), "problem too large for 32-bit TMA coordinates");

    PHM_CUDA_CHECK(cudaMemsetAsync(w.fallback_count, 0, 256, st));
    int rc;
    // operands
    // golden-ratio stride coprime to n_refs (see tc_prep_rows_kernel)
    int64_t perm_a = (int64_t)(0.6180339887498949 * (double)a.n_refs);
    if (perm_a < 1) perm_a = 1;
    auto gcd = [](int64_t x, int64_t y) { while (y) { const int64_t t = x % y; x = y; y = t; } return x; };
    while (gcd(perm_a, a.n_refs) != 1) ++perm_a;
    const int64_t perm_c = a.n_refs / 3;
    if ((rc = launch_prep(a.points, n, n, 1, 0, w.a_hi, w.a_lo, w.norm_points, w.cnorm_points, nullptr, st)) != PHM_OK) return rc;
    if ((rc = launch_prep(a.refs, a.n_refs, ref_pad, perm_a, perm_c, w.b_hi, w.b_lo, w.norm_refs, w.cnorm_refs, w.nbs, st)) != PHM_OK) return rc;
    if ((rc = launch_prep(a.cent_pos, a.n_cent_pos, cp_pad, 1, 0, w.b_hi + ref_pad * KDIM, w.b_lo + ref_pad * KDIM, w.norm_cpos,
                          nullptr, w.nbs + ref_pad, st)) != PHM_OK) return rc;
    if ((rc = launch_prep(a.cent_neg, a.n_cent_neg, cn_pad, 1, 0, w.b_hi + (ref_pad + cp_pad) * KDIM, w.b_lo + (ref_pad + cp_pad) * KDIM,
                          w.norm_cneg, nullptr, w.nbs + ref_pad + cp_pad, st)) != PHM_OK) return rc;

    CUtensorMap map_a_hi, map_a_lo, map_b_hi, map_b_lo;
    if ((rc = make_map(&map_a_hi, w.a_hi, n)) != PHM_OK) return rc;
    if ((rc = make_map(&map_a_lo, w.a_lo, n)) != PHM_OK) return rc;
    if ((rc = make_map(&map_b_hi, w.b_hi, r_pad)) != PHM_OK) return rc;
    if ((rc = make_map(&map_b_lo, w.b_lo, r_pad)) != PHM_OK) return rc;

    TcParams p;
    p.n_points = n;
    p.n_mtiles = (int)((n + BM - 1) / BM);
    p.nt_ref = (int)(ref_pad / BN); p.nt_pos = (int)(cp_pad / BN); p.nt_neg = (int)(cn_pad / BN);
    p.nbs = w.nbs; p.cand = w.cand; p.approx = w.approx;
    PHM_CUDA_CHECK(cudaFuncSetAttribute(score_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    int grid = sm_count();
    if (grid > p.n_mtiles) grid = p.n_mtiles;
    score_tc_kernel<<<grid, NTHREADS, SMEM_BYTES, st>>>(map_a_hi, map_a_lo, map_b_hi, map_b_lo, p);
    PHM_CUDA_CHECK(cudaGetLastError());

    RerankParams r;
    r.points = a.points; r.n_points = n;
    r.refs = a.refs; r.n_refs = a.n_refs; r.n_positive = a.n_positive;
    r.perm_a = perm_a; r.perm_c = perm_c;
    r.cent_pos = a.cent_pos; r.n_cent_pos = a.n_cent_pos; r.cent_neg = a.cent_neg; r.n_cent_neg = a.n_cent_neg;
    r.cnorm_points = w.cnorm_points; r.cnorm_refs = w.cnorm_refs; r.cand = w.cand; r.approx = w.approx;
    r.k_neighbors = a.k_neighbors;
    r.knn = a.knn; r.kmeans = a.kmeans; r.combo = a.combo;
    r.fallback_rows = w.fallback_rows; r.fallback_count = w.fallback_count;
    r.max_rank_error = score_collect_stats ? w.max_rank_error : nullptr;
    r.rows_remeasured = w.rows_remeasured;
    int64_t blocks = (n + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    rerank_kernel<<<(unsigned)blocks, 256, 0, st>>>(r);
    PHM_CUDA_CHECK(cudaGetLastError());

    // rows whose shortlist could not be proven complete: exhaustive float64 kernel, count read on the device
    ScoreArgs f = a;
    f.norm_points = w.norm_points; f.norm_refs = w.norm_refs; f.norm_cpos = w.norm_cpos; f.norm_cneg = w.norm_cneg;
    f.row_list = w.fallback_rows; f.n_rows_dev = w.fallback_count; f.n_rows = n;
    if ((rc = launch_score_exact(f, st)) != PHM_OK) return rc;
    if (kernels_launched) *kernels_launched = 7;
    return PHM_OK;
}

// statistics of the last score_tc call on this workspace: [0] = fallback rows, [1] = max |approx - exact| d2 (as float bits)
int score_tc_stats(const void *ws, unsigned long long *fallback_rows, float *max_rank_error, cudaStream_t st) {
    unsigned char host[256];
    PHM_CUDA_CHECK(cudaMemcpyAsync(host, ws, 256, cudaMemcpyDeviceToHost, st));
    PHM_CUDA_CHECK(cudaStreamSynchronize(st));
    memcpy(fallback_rows, host, 8);
    memcpy(max_rank_error, host + 64, 8);      // [0] absolute, [1] relative to |a|^2 + |b|^2
    unsigned long long remeasured = 0;
    memcpy(&remeasured, host + 128, 8);
    max_rank_error[2] = (float)remeasured;     // rows whose mixed neighbour band was re-measured exactly
    return PHM_OK;
}

}  // namespace tc
}  // namespace phm
