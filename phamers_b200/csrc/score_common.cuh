// Declarations shared by the two scoring paths (score_exact.cu, score_tc.cu).
#pragma once
#include "phm_common.cuh"

namespace phm {

struct ScoreArgs {
    const double *points; int64_t n_points; int dim;
    const uint32_t *point_counts;             // optional: raw count rows instead of `points` (features = row / row sum, formed on the fly)
    const double *refs; int64_t n_refs; int64_t n_positive;
    const double *cent_pos; int64_t n_cent_pos;
    const double *cent_neg; int64_t n_cent_neg;
    const double *norm_points, *norm_refs, *norm_cpos, *norm_cneg;    // squared row norms (float64)
    const int64_t *row_list;                  // optional: score only these rows
    const unsigned long long *n_rows_dev;     // optional: number of rows in row_list, read on the device
    int64_t n_rows;                           // rows to score (upper bound when n_rows_dev is set)
    int k_neighbors;
    double *knn, *kmeans, *combo;
};

int launch_score_exact(const ScoreArgs &a, cudaStream_t st);
int launch_row_norms(const double *x, int64_t n_rows, int dim, double *out, cudaStream_t st);

namespace tc {
bool score_tc_supported(int dim, int k_neighbors, int64_t n_refs, int64_t n_cent_pos, int64_t n_cent_neg);
size_t score_tc_workspace_bytes(int64_t n_points, int64_t n_refs, int64_t n_cent_pos, int64_t n_cent_neg);
int score_tc(const ScoreArgs &a, void *ws, size_t ws_bytes, cudaStream_t st, int *kernels_launched);
int score_tc_last_ms(float *ms);
int score_tc_stats(const void *ws, unsigned long long *fallback_rows, float *max_rank_error, cudaStream_t st);
}  // namespace tc

extern int score_collect_stats;
extern int score_path;        // 0 = auto (tensor cores when supported), 1 = exact float64 only, 2 = tensor cores required

}  // namespace phm
