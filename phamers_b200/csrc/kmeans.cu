// Lloyd iterations of k-means on the device: the clustering of a reference set that the reference runs as
// scikit-learn KMeans(n_clusters = 86, random_state = 10).fit(data).labels_ (scripts/learning.py:131-146, twice per scoring call,
// scripts/phamer.py:245-248).  It only matters when the reference set changes from call to call (cross validation: 40 fits).
//
// What is reproduced is scikit-learn's single Lloyd run (sklearn/cluster/_kmeans.py::_kmeans_single_lloyd) on the CENTRED data and
// from the k-means++ centres that scikit-learn's own seeding gives (the caller obtains both with scikit-learn's functions, so the
// random stream is the reference's):
//     repeat:  E step   label_i = argmin_c (|c|^2 - 2 x_i.c), first minimum on ties          (lloyd_iter_chunked_dense)
//              M step   centre_c = mean of its points; an empty cluster is NOT relocated here -- the call reports it and the
//                       caller falls back to scikit-learn for that fit
//              stop     labels unchanged (strict convergence), or sum_c |centre_c - old centre_c|^2 <= tol, or max_iter
//     if not strictly converged: one more E step with the final centres
// All arithmetic is float64.  Dot products and centre sums are formed in a different order than scikit-learn's BLAS / per-thread
// partial sums form them, so centres agree to rounding, and labels agree unless two centres are equidistant from a point to within
// ~1e-16 -- tests/test_gpu_kmeans.py checks label-for-label equality on the shipped reference sets and on cross-validation folds.
#include "phm_common.cuh"

namespace phm {
namespace km {

constexpr int CHUNK = 32;          // centres staged in shared memory at a time

// E step.  One warp per point: the point's features live in registers (dim <= 1024: at most 32 per lane); the centres stream
// through shared memory CHUNK at a time.  changed += 1 for every point whose label moved.
template <int PER_LANE>
__global__ void __launch_bounds__(256) assign_kernel(const double *__restrict__ x, int64_t n, int dim, const double *__restrict__ centres,
                                                     int k, int32_t *__restrict__ labels, unsigned long long *__restrict__ changed) {
    extern __shared__ double s_c[];                      // [CHUNK][dim] centres, then [CHUNK] squared norms
    double *s_norm = s_c + (size_t)CHUNK * dim;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t rows_per_block = blockDim.x >> 5;
    for (int64_t base = (int64_t)blockIdx.x * rows_per_block; base < n; base += (int64_t)gridDim.x * rows_per_block) {
        const int64_t i = base + warp;
        double xv[PER_LANE];
#pragma unroll
        for (int j = 0; j < PER_LANE; ++j) {
            const int d = lane + 32 * j;
            xv[j] = (i < n && d < dim) ? x[i * dim + d] : 0.0;
        }
        double best = INFINITY;
        int best_c = 0;
        for (int c0 = 0; c0 < k; c0 += CHUNK) {
            const int nc = k - c0 < CHUNK ? k - c0 : CHUNK;
            __syncthreads();
            for (int idx = threadIdx.x; idx < nc * dim; idx += blockDim.x) s_c[idx] = centres[(int64_t)c0 * dim + idx];
            __syncthreads();
            for (int c = warp; c < nc; c += (int)rows_per_block) {        // squared norms of the staged centres
                double a = 0.0;
                for (int d = lane; d < dim; d += 32) a = fma(s_c[c * dim + d], s_c[c * dim + d], a);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(FULL, a, o);
                if (lane == 0) s_norm[c] = a;
            }
            __syncthreads();
            for (int c = 0; c < nc; ++c) {
                double dot = 0.0;
#pragma unroll
                for (int j = 0; j < PER_LANE; ++j) {
                    const int d = lane + 32 * j;
                    if (d < dim) dot = fma(xv[j], s_c[c * dim + d], dot);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(FULL, dot, o);
                const double dist = s_norm[c] - 2.0 * dot;                 // |x|^2 is the same for every centre
                if (dist < best) { best = dist; best_c = c0 + c; }         // strict: the first minimum wins
            }
        }
        if (i < n && lane == 0) {
            if (labels[i] != best_c) atomicAdd(changed, 1ull);
            labels[i] = best_c;
        }
    }
}

// M step.  One CTA per centre, one thread per feature: the members are added in ascending point order (deterministic); the squared
// shift of the centre is added to shift2[0], an empty cluster sets empty[0] and keeps its centre.
__global__ void __launch_bounds__(256) update_kernel(const double *__restrict__ x, int64_t n, int dim, const int32_t *__restrict__ labels,
                                                     double *__restrict__ centres, int k, double *__restrict__ shift2,
                                                     unsigned long long *__restrict__ empty) {
    __shared__ double s_red[256];
    const int c = blockIdx.x;
    long long count = 0;
    double part = 0.0;
    for (int d0 = 0; d0 < dim; d0 += blockDim.x) {
        const int d = d0 + threadIdx.x;
        double acc = 0.0;
        count = 0;
        for (int64_t i = 0; i < n; ++i) {
            if (labels[i] == c) {                          // the same address for the whole CTA: a broadcast load
                if (d < dim) acc += x[i * dim + d];
                ++count;
            }
        }
        if (d < dim && count > 0) {
            const double nv = acc / (double)count, ov = centres[(int64_t)c * dim + d];
            centres[(int64_t)c * dim + d] = nv;
            part += (nv - ov) * (nv - ov);
        }
    }
    s_red[threadIdx.x] = part;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) s_red[threadIdx.x] += s_red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        atomicAdd(shift2, s_red[0]);
        if (count == 0) atomicAdd(empty, 1ull);
    }
}

struct Control { unsigned long long changed, empty; double shift2; double pad; };

}  // namespace km
}  // namespace phm

using namespace phm;

extern "C" size_t phm_kmeans_workspace_bytes(int64_t, int, int) { return 256; }

// Synchronising entry point (one small device->host read per iteration decides convergence on the host, as scikit-learn does).
// d_x float64[n, dim] centred data, d_centres float64[k, dim] in: initial centres, out: final centres, d_labels int32[n] out.
// h_info int64[3] out (host): [0] iterations run, [1] 1 if converged strictly (labels stopped changing), [2] 1 if some cluster
// became empty (the result is then NOT scikit-learn's, which relocates empty clusters: the caller must fall back).
extern "C" int phm_kmeans_lloyd(const double *d_x, int64_t n, int dim, double *d_centres, int k, int max_iter, double tol,
                                int32_t *d_labels, int64_t *h_info, void *d_workspace, size_t workspace_bytes, void *stream) {
    PHM_REQUIRE(n >= 1 && dim >= 1 && dim <= 1024 && k >= 1 && max_iter >= 1, "bad shape");
    PHM_REQUIRE(d_x && d_centres && d_labels && h_info && d_workspace, "null pointer");
    if (workspace_bytes < 256) { set_error("workspace too small"); return PHM_E_WORKSPACE; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    km::Control *ctl = static_cast<km::Control *>(d_workspace);
    PHM_CUDA_CHECK(cudaMemsetAsync(d_labels, 0xFF, (size_t)n * sizeof(int32_t), st));            // -1: every label "changes" in the first pass
    const size_t smem = ((size_t)km::CHUNK * dim + km::CHUNK) * sizeof(double);
    const int per_lane = (dim + 31) / 32;
    auto assign = [&](void) -> int {
        int64_t blocks = (n + 7) / 8;
        if (blocks > (int64_t)sm_count() * 8) blocks = (int64_t)sm_count() * 8;
#define PHM_KM_ASSIGN(P)                                                                                                              \
        do {                                                                                                                          \
            PHM_CUDA_CHECK(cudaFuncSetAttribute(km::assign_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
            km::assign_kernel<P><<<(unsigned)blocks, 256, smem, st>>>(d_x, n, dim, d_centres, k, d_labels, &ctl->changed);            \
        } while (0)
        if (per_lane <= 8) PHM_KM_ASSIGN(8);
        else if (per_lane <= 16) PHM_KM_ASSIGN(16);
        else PHM_KM_ASSIGN(32);
#undef PHM_KM_ASSIGN
        PHM_LAUNCH_CHECK();
        return PHM_OK;
    };
    h_info[0] = 0; h_info[1] = 0; h_info[2] = 0;
    int rc;
    for (int it = 0; it < max_iter; ++it) {
        PHM_CUDA_CHECK(cudaMemsetAsync(ctl, 0, sizeof(km::Control), st));
        if ((rc = assign()) != PHM_OK) return rc;                                                   // E step with the current centres
        km::update_kernel<<<(unsigned)k, 256, 0, st>>>(d_x, n, dim, d_labels, d_centres, k, &ctl->shift2, &ctl->empty);   // M step
        PHM_LAUNCH_CHECK();
        km::Control h;
        PHM_CUDA_CHECK(cudaMemcpyAsync(&h, ctl, sizeof(h), cudaMemcpyDeviceToHost, st));
        PHM_CUDA_CHECK(cudaStreamSynchronize(st));
        h_info[0] = it + 1;
        if (h.empty) { h_info[2] = 1; return PHM_OK; }
        if (h.changed == 0) { h_info[1] = 1; break; }                                               // labels equal to the previous iteration's
        if (h.shift2 <= tol) break;
    }
    if (!h_info[1]) {                                                                               // labels of the FINAL centres
        PHM_CUDA_CHECK(cudaMemsetAsync(ctl, 0, sizeof(km::Control), st));
        if ((rc = assign()) != PHM_OK) return rc;
        PHM_CUDA_CHECK(cudaStreamSynchronize(st));
    }
    return PHM_OK;
}
