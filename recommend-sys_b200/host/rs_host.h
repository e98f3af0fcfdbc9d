/*
 * rs_host.h — host-side (CPU) pieces of the drop-in that stay on the host in the reference
 * too and feed the device path through include/rs_knn.h.  Not part of the accelerated path.
 *
 *   rs_host_inner_ids     NewTrainSet's id maps (core/data.go:137-151): inner id = order of
 *                         first appearance; returns the number of distinct ids.
 *   rs_host_baseline_sgd  BaseLine.Fit (core/base.go:135-163): strictly sequential SGD over
 *                         the ratings in dataset order (every step depends on the previous
 *                         one through globalBias, so it cannot be parallelised bit-exactly).
 *                         KNN-baseline only needs its bias vectors (core/knn.go:179-187).
 * Built with -ffp-contract=off (Go on amd64 never fuses x*y+z).
 */
#ifndef RS_HOST_H
#define RS_HOST_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

int64_t rs_host_inner_ids(const int64_t *raw, int64_t n, int32_t *inner_out);

void rs_host_baseline_sgd(const int32_t *inner_user, const int32_t *inner_item, const double *rating,
                          int64_t n, int32_t n_users, int32_t n_items, double reg, double lr,
                          int32_t n_epochs, double *user_bias, double *item_bias, double *global_bias);

/* Batch ConvertUserID / ConvertItemID (core/data.go:157-183) through a dense raw -> inner table
 * (table[raw] = inner id or -1; raw ids outside [0, n_table) are new ids = -1). */
/* Routing of test pairs to the cyclic row shards of a multi-GPU Fit: order[] = pair indices grouped by owner
 * shard (stable), counts[world]; owner = (left / block) % world, unknown left ids round-robin. */
void rs_host_route_pairs(const int32_t *left_inner, int64_t n, int32_t world, int32_t block, int64_t *order,
                         int64_t *counts);
/* SURVEY.md §8 f-4.  Neighbour lists (int32 idx, float64 sim)[n_rows][k] on disk, checksummed; replaces
 * gob Save/Load (core/dump.go:11-36) for the one artefact of a top-k-only Fit.  load with idx == NULL
 * returns the header only.  0 ok, -1 io error, -2 not a neighbour-list file / checksum mismatch. */
int32_t rs_host_save_neighbors(const char *path, int64_t n_rows, int32_t k, const int32_t *idx, const double *sim);
int32_t rs_host_load_neighbors(const char *path, int64_t *n_rows, int32_t *k, int32_t *idx, double *sim);
/* LoadDataFromFile (core/data.go:287-310).  float_ratings = 0 reproduces the reference (strconv.Atoi on the
 * rating: half-stars and headers become 0); != 0 keeps fractional ratings.  users == NULL counts rows. */
int64_t rs_host_load_ratings(const char *path, const char *sep, int32_t float_ratings, int32_t skip_header,
                             int64_t *users, int64_t *items, double *ratings, int64_t cap);
/* TrainSet.GlobalMean (core/data.go:134): sequential sum / n. */
double rs_host_mean_seq(const double *x, int64_t n);
void rs_host_convert_dense(const int32_t *table, int64_t n_table, const int64_t *raw, int64_t n, int32_t *inner_out);
/* the same on `threads` host threads (large test sets; the conversion runs beside the similarity kernel) */
void rs_host_convert_dense_mt(const int32_t *table, int64_t n_table, const int64_t *raw, int64_t n, int32_t *inner_out,
                              int32_t threads);

/* Deterministic synthetic rating matrices of the BASELINE.json shapes:
 * unique (user,item) pairs, heavy-tailed degrees, integer ratings 1..5 with a MovieLens-like
 * marginal, rows emitted in a seeded shuffled order.  Returns the number of ratings written
 * (<= nnz_target; exactly nnz_target unless the matrix is too small). */
int64_t rs_host_synth_ratings(int32_t n_users, int32_t n_items, int64_t nnz_target, uint64_t seed,
                              int64_t *users, int64_t *items, double *ratings);

#ifdef __cplusplus
}
#endif
#endif
