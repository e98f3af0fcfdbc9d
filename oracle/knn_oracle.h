/*
 * knn_oracle.h — CPU restatement of the reference's KNN hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (recommend-sys_b200/) never links, imports or calls anything in oracle/.
 *
 * What it restates (all paths relative to /root/reference):
 *   core/data.go:131-154   NewTrainSet (inner ids by first appearance, GlobalMean)
 *   core/data.go:185-216   UserRatings / ItemRatings (adjacency in dataset order)
 *   core/data.go:222-243   means, sorts (in-place sort of every left row by id)
 *   core/sim.go:10-81      Cosine, MSD, Pearson (verbatim operation order)
 *   core/knn.go:143-217    KNN.Fit (NaN matrix, diagonal unset, NaN results not stored)
 *   core/knn.go:75-141     KNN.Predict (<= mink, no sim>0 filter, sums in sorted order)
 *   core/base.go:122-163   BaseLine.Fit / Predict (sequential SGD biases)
 *   core/data.go:98-105    DataSet.Predict (serial batch loop)
 *   core/utils.go:110-120  newNanMatrix
 *
 * Third-party arithmetic on the path that is NOT under /root/reference:
 *   - Go stdlib sort.Sort (pdqsort, Go >= 1.19; go.mod:3 says go 1.24) — restated in
 *     pdqsort_go() from the published algorithm.  No reference
 *     test pins its tie permutation: the `OR_TIE_GO` policy is PARITY UNPINNED.
 *   - gonum v0.9.1 stat.Mean (go.mod:7), call site core/data.go:134.  Restated as a
 *     sequential sum / n; for integer ratings the sum is exact in any order, so the
 *     result is order independent.  For non-integer ratings gonum's SIMD summation
 *     order is not reproduced: PARITY UNPINNED for that sub-case.
 *
 * Pinning: tests/test_oracle_pin.py checks this file against every known-answer the
 * reference's own tests hold for the path (core/sim_test.go:10-59 at exact float64
 * value) and against the statistical bounds of core/base_test.go:50-64 on the real
 * ml-100k folds.  Everything else (item-based mode, Cosine/Pearson inside KNN,
 * individual predictions, tie order, NaN handling) is pinned only by this restatement.
 *
 * Build: gcc -O2 -ffp-contract=off (Go on amd64/GOAMD64=v1 never fuses x*y+z).
 */
#ifndef KNN_ORACLE_H
#define KNN_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { OR_SIM_COSINE = 0, OR_SIM_MSD = 1, OR_SIM_PEARSON = 2, OR_SIM_PEARSON_BASELINE = 3 };
enum { OR_KNN_BASIC = 0, OR_KNN_CENTERED = 1, OR_KNN_ZSCORE = 2, OR_KNN_BASELINE = 3 };
enum { OR_TIE_GO = 0, OR_TIE_CANONICAL = 1 };

/* core/data.go:124-127 — Go's `int` is 64-bit, so the struct is 16 bytes as in Go. */
typedef struct { int64_t id; double rating; } or_idrating;

typedef struct or_trainset or_trainset;
typedef struct or_knn or_knn;

typedef struct {
    int sim;            /* OR_SIM_*   (Parameters["sim"],       default MSD,  core/knn.go:145) */
    int knn_type;       /* OR_KNN_*   (fixed by constructor,    core/knn.go:50-73)             */
    int user_based;     /*            (Parameters["userBased"], default true, core/knn.go:146) */
    int k;              /*            (Parameters["k"],         default 40,   core/knn.go:80)  */
    int min_k;          /*            (Parameters["mink"],      default 1,    core/knn.go:81)  */
    int n_jobs;         /*            (Parameters["nJobs"],     default NumCPU, core/knn.go:148) */
    int tie_policy;     /* OR_TIE_*                                                            */
    double reg;         /* baseline   (Parameters["reg"],       default 0.02, core/base.go:137) */
    double lr;          /* baseline   (Parameters["lr"],        default 0.005, core/base.go:138) */
    int n_epochs;       /* baseline   (Parameters["nEpochs"],   default 20,   core/base.go:139) */
    double shrinkage;   /* PearsonBaseline extension only (not in the reference)               */
    int baseline_als;   /* EXTENSION: 1 = ALS baselines (or_baseline_als) instead of the SGD           */
    int als_epochs;     /*            default 10                                                       */
    double reg_u;       /*            default 15                                                       */
    double reg_i;       /*            default 10                                                       */
} or_params;

void or_params_default(or_params *p);

/* ---- similarity functions on explicit lists (core/sim.go; sim_test.go builds them
 *      with NewSortedIdRatings which sorts by id first, core/data.go:249-253) ---- */
double or_sim_lists(int sim, const or_idrating *a, int64_t na, const or_idrating *b, int64_t nb);
/* Go's sort.Sort applied to an IDRating list by id (core/data.go:245-265). */
void or_sort_by_id(or_idrating *a, int64_t n);

/* ---- TrainSet (core/data.go:109-216) ---- */
or_trainset *or_trainset_new(const int64_t *users, const int64_t *items, const double *ratings, int64_t n);
void or_trainset_free(or_trainset *t);
int64_t or_trainset_user_count(const or_trainset *t);
int64_t or_trainset_item_count(const or_trainset *t);
double or_trainset_global_mean(const or_trainset *t);
int64_t or_trainset_convert_user(const or_trainset *t, int64_t raw);  /* -1 = newID */
int64_t or_trainset_convert_item(const or_trainset *t, int64_t raw);
/* inner ids of every rating row, in dataset order (what the cgo bridge marshals) */
const int32_t *or_trainset_inner_users(const or_trainset *t);
const int32_t *or_trainset_inner_items(const or_trainset *t);

/* ---- KNN (core/knn.go) ---- */
or_knn *or_knn_new(const or_params *p);
void or_knn_free(or_knn *k);
/* Fit keeps a pointer to `t` (the reference keeps a copy of the TrainSet, core/knn.go:150). */
void or_knn_fit(or_knn *k, or_trainset *t);
/* Only rows [row0,row1) of the left matrix against all N (timing slab for big shapes).
 * Entries outside the slab rows stay NaN unless written by symmetry. */
void or_knn_fit_rows(or_knn *k, or_trainset *t, int64_t row0, int64_t row1);
double or_knn_predict(const or_knn *k, int64_t raw_user, int64_t raw_item);
/* core/data.go:98-105: serial loop.  n_threads > 1 = the "parallel over pairs" variant. */
void or_knn_predict_batch(const or_knn *k, const int64_t *users, const int64_t *items, int64_t n,
                          double *out, int n_threads);
/* Neighbours actually used by Predict, in accumulation order (ids of the left side).
 * Returns the count (0 when the GlobalMean branch was taken). */
int or_knn_predict_neighbors(const or_knn *k, int64_t raw_user, int64_t raw_item,
                             int64_t *ids, double *sims, int cap);

int64_t or_knn_n(const or_knn *k);                   /* rows of Sims            */
const double *or_knn_sims_row(const or_knn *k, int64_t row);
const double *or_knn_means(const or_knn *k);         /* NULL unless centered/zscore */
const double *or_knn_stddevs(const or_knn *k);       /* NULL unless zscore      */
const double *or_knn_bias(const or_knn *k);          /* NULL unless baseline    */
double or_knn_global_mean(const or_knn *k);

/* Row-wise top-k of the Sims matrix under the canonical order (similarity desc, id asc),
 * NaN skipped; unused slots get idx -1 / sim NaN.  This artefact is not in the reference
 * (SURVEY.md §7.3 item 3); it is config 4's own output. */
void or_knn_topk(const or_knn *k, int kk, int64_t row0, int64_t row1, int32_t *idx, double *sim);

/* Exact integer co-rating sums for one pair of left rows (after Fit, rows are id-sorted):
 * out = {count, sum_x, sum_y, sum_xx, sum_yy, sum_xy}.  Ratings must be integers. */
void or_knn_pair_sums(const or_knn *k, int64_t a, int64_t b, int64_t out[6]);

/* ---- BaseLine (core/base.go:108-163) ---- */
/* EXTENSION, parity unpinned: ALS baselines (see knn_oracle.c) — the checker of rs_baseline_als. */
void or_baseline_als(or_trainset *t, double reg_u, double reg_i, int n_epochs, double *user_bias,
                     double *item_bias);
void or_baseline_fit(const or_trainset *t, double reg, double lr, int n_epochs,
                     double *user_bias, double *item_bias, double *global_bias);

/* ---- metrics (intended signatures; core/utils.go:162-180) ---- */
double or_rmse(const double *pred, const double *truth, int64_t n);
double or_mae(const double *pred, const double *truth, int64_t n);

/* Go sort.Sort exposed for tests of the port itself. */
typedef struct {
    void *ctx;
    int (*less)(void *ctx, int64_t i, int64_t j);
    void (*swap)(void *ctx, int64_t i, int64_t j);
} or_sort_iface;
void or_go_sort(or_sort_iface *s, int64_t n);

/* ---- Slope One (core/slope_one.go) ---- */
typedef struct or_slope or_slope;
or_slope *or_slope_fit(or_trainset *t, int n_jobs);
or_slope *or_slope_fit_rows(or_trainset *t, int n_jobs, int64_t row0, int64_t row1);   /* bench slabs only */     /* nJobs = runtime.NumCPU() in the reference */
void or_slope_free(or_slope *s);
int64_t or_slope_n(const or_slope *s);
const double *or_slope_dev(const or_slope *s);          /* n x n, row-major */
const double *or_slope_user_means(const or_slope *s);
double or_slope_predict(const or_slope *s, int64_t raw_user, int64_t raw_item);
void or_slope_predict_batch(const or_slope *s, const int64_t *users, const int64_t *items, int64_t n, double *out);

/* Similarities of a few left rows against all N (test helper for shapes whose N x N matrix does not
 * fit in host memory); out is n_rows x N, NaN on the diagonal and where there is no co-rating. */
void or_rows_sims(or_trainset *t, int sim, int user_based, const int64_t *rows, int64_t n_rows, double *out);

#ifdef __cplusplus
}
#endif
#endif
