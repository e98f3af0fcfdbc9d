/*
 * knn_oracle.c — CPU restatement of the reference's KNN hot path (see knn_oracle.h).
 * TEST INFRASTRUCTURE ONLY.  Every function cites the reference lines it follows
 * (paths relative to /root/reference).  Build with -O2 -ffp-contract=off.
 */
#define _GNU_SOURCE
#include "knn_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* =====================================================================================
 * Go stdlib sort.Sort — pdqsort (sort/zsortinterface.go, Go >= 1.19).  Not under
 * /root/reference; restated from the published algorithm.
 * Call sites in the reference: core/knn.go:108 (tie order matters) and
 * core/data.go:240,251 (unique keys).
 * ===================================================================================== */
#define LESS(s, i, j) ((s)->less((s)->ctx, (i), (j)))
#define SWAP(s, i, j) ((s)->swap((s)->ctx, (i), (j)))

enum { HINT_UNKNOWN = 0, HINT_INCREASING = 1, HINT_DECREASING = 2 };

static int bits_len(uint64_t x) { int n = 0; while (x) { n++; x >>= 1; } return n; }

static void go_insertion_sort(or_sort_iface *s, int64_t a, int64_t b) {
    for (int64_t i = a + 1; i < b; i++)
        for (int64_t j = i; j > a && LESS(s, j, j - 1); j--) SWAP(s, j, j - 1);
}

static void go_sift_down(or_sort_iface *s, int64_t lo, int64_t hi, int64_t first) {
    int64_t root = lo;
    for (;;) {
        int64_t child = 2 * root + 1;
        if (child >= hi) return;
        if (child + 1 < hi && LESS(s, first + child, first + child + 1)) child++;
        if (!LESS(s, first + root, first + child)) return;
        SWAP(s, first + root, first + child);
        root = child;
    }
}

static void go_heap_sort(or_sort_iface *s, int64_t a, int64_t b) {
    int64_t first = a, lo = 0, hi = b - a;
    for (int64_t i = (hi - 1) / 2; i >= 0; i--) go_sift_down(s, i, hi, first);
    for (int64_t i = hi - 1; i >= 0; i--) {
        SWAP(s, first, first + i);
        go_sift_down(s, lo, i, first);
    }
}

static void go_order2(or_sort_iface *s, int64_t *a, int64_t *b, int *swaps) {
    if (LESS(s, *b, *a)) { int64_t t = *a; *a = *b; *b = t; (*swaps)++; }
}

static int64_t go_median(or_sort_iface *s, int64_t a, int64_t b, int64_t c, int *swaps) {
    go_order2(s, &a, &b, swaps);
    go_order2(s, &b, &c, swaps);
    go_order2(s, &a, &b, swaps);
    return b;
}

static int64_t go_median_adjacent(or_sort_iface *s, int64_t a, int *swaps) {
    return go_median(s, a - 1, a, a + 1, swaps);
}

static int64_t go_choose_pivot(or_sort_iface *s, int64_t a, int64_t b, int *hint) {
    const int64_t shortest_ninther = 50;
    const int max_swaps = 4 * 3;
    int64_t l = b - a;
    int swaps = 0;
    int64_t i = a + l / 4 * 1, j = a + l / 4 * 2, k = a + l / 4 * 3;
    if (l >= 8) {
        if (l >= shortest_ninther) {
            i = go_median_adjacent(s, i, &swaps);
            j = go_median_adjacent(s, j, &swaps);
            k = go_median_adjacent(s, k, &swaps);
        }
        j = go_median(s, i, j, k, &swaps);
    }
    if (swaps == 0) *hint = HINT_INCREASING;
    else if (swaps == max_swaps) *hint = HINT_DECREASING;
    else *hint = HINT_UNKNOWN;
    return j;
}

static void go_reverse_range(or_sort_iface *s, int64_t a, int64_t b) {
    int64_t i = a, j = b - 1;
    while (i < j) { SWAP(s, i, j); i++; j--; }
}

static int go_partial_insertion_sort(or_sort_iface *s, int64_t a, int64_t b) {
    const int max_steps = 5;
    const int64_t shortest_shifting = 50;
    int64_t i = a + 1;
    for (int step = 0; step < max_steps; step++) {
        while (i < b && !LESS(s, i, i - 1)) i++;
        if (i == b) return 1;
        if (b - a < shortest_shifting) return 0;
        SWAP(s, i, i - 1);
        if (i - a >= 2) {
            for (int64_t j = i - 1; j >= 1; j--) {
                if (!LESS(s, j, j - 1)) break;
                SWAP(s, j, j - 1);
            }
        }
        if (b - i >= 2) {
            for (int64_t j = i + 1; j < b; j++) {
                if (!LESS(s, j, j - 1)) break;
                SWAP(s, j, j - 1);
            }
        }
    }
    return 0;
}

static void go_break_patterns(or_sort_iface *s, int64_t a, int64_t b) {
    int64_t length = b - a;
    if (length >= 8) {
        uint64_t r = (uint64_t)length;
        uint64_t modulus = (uint64_t)1 << bits_len((uint64_t)length);
        int64_t idx = a + (length / 4) * 2 - 1;
        for (int i = 0; i < 3; i++) {
            r ^= r << 13; r ^= r >> 7; r ^= r << 17;
            int64_t other = (int64_t)(r & (modulus - 1));
            if (other >= length) other -= length;
            SWAP(s, idx - 1 + i, a + other);
        }
    }
}

static int64_t go_partition_equal(or_sort_iface *s, int64_t a, int64_t b, int64_t pivot) {
    SWAP(s, a, pivot);
    int64_t i = a + 1, j = b - 1;
    for (;;) {
        while (i <= j && !LESS(s, a, i)) i++;
        while (i <= j && LESS(s, a, j)) j--;
        if (i > j) break;
        SWAP(s, i, j);
        i++; j--;
    }
    return i;
}

static int64_t go_partition(or_sort_iface *s, int64_t a, int64_t b, int64_t pivot, int *already) {
    SWAP(s, a, pivot);
    int64_t i = a + 1, j = b - 1;
    while (i <= j && LESS(s, i, a)) i++;
    while (i <= j && !LESS(s, j, a)) j--;
    if (i > j) { SWAP(s, j, a); *already = 1; return j; }
    SWAP(s, i, j);
    i++; j--;
    for (;;) {
        while (i <= j && LESS(s, i, a)) i++;
        while (i <= j && !LESS(s, j, a)) j--;
        if (i > j) break;
        SWAP(s, i, j);
        i++; j--;
    }
    SWAP(s, j, a);
    *already = 0;
    return j;
}

static void go_pdqsort(or_sort_iface *s, int64_t a, int64_t b, int limit) {
    const int64_t max_insertion = 12;
    int was_balanced = 1, was_partitioned = 1;
    for (;;) {
        int64_t length = b - a;
        if (length <= max_insertion) { go_insertion_sort(s, a, b); return; }
        if (limit == 0) { go_heap_sort(s, a, b); return; }
        if (!was_balanced) { go_break_patterns(s, a, b); limit--; }
        int hint;
        int64_t pivot = go_choose_pivot(s, a, b, &hint);
        if (hint == HINT_DECREASING) {
            go_reverse_range(s, a, b);
            pivot = (b - 1) - (pivot - a);
            hint = HINT_INCREASING;
        }
        if (was_balanced && was_partitioned && hint == HINT_INCREASING) {
            if (go_partial_insertion_sort(s, a, b)) return;
        }
        if (a > 0 && !LESS(s, a - 1, pivot)) {
            a = go_partition_equal(s, a, b, pivot);
            continue;
        }
        int already;
        int64_t mid = go_partition(s, a, b, pivot, &already);
        was_partitioned = already;
        int64_t left_len = mid - a, right_len = b - mid;
        int64_t balance_threshold = length / 8;
        if (left_len < right_len) {
            was_balanced = left_len >= balance_threshold;
            go_pdqsort(s, a, mid, limit);
            a = mid + 1;
        } else {
            was_balanced = right_len >= balance_threshold;
            go_pdqsort(s, mid + 1, b, limit);
            b = mid;
        }
    }
}

void or_go_sort(or_sort_iface *s, int64_t n) {
    if (n <= 1) return;
    go_pdqsort(s, 0, n, bits_len((uint64_t)n));
}

/* ---- SortedIdRatings as a sort.Interface (core/data.go:255-265) ---- */
static int idr_less(void *ctx, int64_t i, int64_t j) {
    or_idrating *d = (or_idrating *)ctx;
    return d[i].id < d[j].id;
}
static void idr_swap(void *ctx, int64_t i, int64_t j) {
    or_idrating *d = (or_idrating *)ctx;
    or_idrating t = d[i]; d[i] = d[j]; d[j] = t;
}
void or_sort_by_id(or_idrating *a, int64_t n) {
    or_sort_iface s = { a, idr_less, idr_swap };
    or_go_sort(&s, n);
}

/* =====================================================================================
 * Similarities — core/sim.go, verbatim operation order.
 * ===================================================================================== */

/* core/sim.go:10-25 */
static double sim_cosine(const or_idrating *a, int64_t na, const or_idrating *b, int64_t nb) {
    double m = .0, n = .0, l = .0;
    int64_t ptr = 0;
    for (int64_t x = 0; x < na; x++) {
        while (ptr < nb && b[ptr].id < a[x].id) ptr++;
        if (ptr < nb && b[ptr].id == a[x].id) {
            double ir = a[x].rating, jr = b[ptr].rating;
            m += ir * ir;
            n += jr * jr;
            l += ir * jr;
        }
    }
    return l / (sqrt(m) * sqrt(n));
}

/* core/sim.go:28-44 */
static double sim_msd(const or_idrating *a, int64_t na, const or_idrating *b, int64_t nb) {
    double count = 0.0, sum = 0.0;
    int64_t ptr = 0;
    for (int64_t x = 0; x < na; x++) {
        while (ptr < nb && b[ptr].id < a[x].id) ptr++;
        if (ptr < nb && b[ptr].id == a[x].id) {
            double ir = a[x].rating, jr = b[ptr].rating;
            sum += (ir - jr) * (ir - jr);
            count++;
        }
    }
    return 1.0 / (sum / count + 1.0);
}

/* core/sim.go:47-81 */
static double sim_pearson(const or_idrating *a, int64_t na, const or_idrating *b, int64_t nb) {
    double count = .0, sum = .0;
    for (int64_t x = 0; x < na; x++) { sum += a[x].rating; count += 1; }
    double mean_a = sum / count;
    count = .0; sum = .0;
    for (int64_t x = 0; x < nb; x++) { sum += b[x].rating; count += 1; }
    double mean_b = sum / count;
    double m = .0, n = .0, l = .0;
    int64_t ptr = 0;
    for (int64_t x = 0; x < na; x++) {
        while (ptr < nb && b[ptr].id < a[x].id) ptr++;
        if (ptr < nb && b[ptr].id == a[x].id) {
            double rating_a = a[x].rating - mean_a;
            double rating_b = b[ptr].rating - mean_b;
            m += rating_a * rating_a;
            n += rating_b * rating_b;
            l += rating_a * rating_b;
        }
    }
    return l / (sqrt(m) * sqrt(n));
}

double or_sim_lists(int sim, const or_idrating *a, int64_t na, const or_idrating *b, int64_t nb) {
    switch (sim) {
    case OR_SIM_COSINE: return sim_cosine(a, na, b, nb);
    case OR_SIM_MSD: return sim_msd(a, na, b, nb);
    case OR_SIM_PEARSON: return sim_pearson(a, na, b, nb);
    default: return NAN;
    }
}

/* =====================================================================================
 * TrainSet — core/data.go:109-216.
 * ===================================================================================== */
typedef struct { int64_t *keys; int32_t *vals; uint64_t cap; } idmap;  /* Go map[int]int */

static void idmap_init(idmap *m, uint64_t hint) {
    uint64_t cap = 16;
    while (cap < hint * 2 + 2) cap <<= 1;
    m->cap = cap;
    m->keys = (int64_t *)malloc(cap * sizeof(int64_t));
    m->vals = (int32_t *)malloc(cap * sizeof(int32_t));
    for (uint64_t i = 0; i < cap; i++) m->vals[i] = -1;
}
static uint64_t idmap_hash(int64_t k) {
    uint64_t x = (uint64_t)k;
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}
static int32_t idmap_get(const idmap *m, int64_t k) {
    uint64_t h = idmap_hash(k) & (m->cap - 1);
    while (m->vals[h] != -1) {
        if (m->keys[h] == k) return m->vals[h];
        h = (h + 1) & (m->cap - 1);
    }
    return -1;
}
static void idmap_put(idmap *m, int64_t k, int32_t v) {
    uint64_t h = idmap_hash(k) & (m->cap - 1);
    while (m->vals[h] != -1) h = (h + 1) & (m->cap - 1);
    m->keys[h] = k; m->vals[h] = v;
}
static void idmap_free(idmap *m) { free(m->keys); free(m->vals); }

struct or_trainset {
    int64_t n;
    int64_t *users, *items;
    double *ratings;
    double global_mean;
    int64_t user_count, item_count;
    idmap inner_users, inner_items;
    int32_t *iu, *ii;                 /* inner ids per row, dataset order */
    or_idrating **user_ratings;       /* lazily built, core/data.go:185-199 */
    int64_t *user_len;
    or_idrating *user_store;
    or_idrating **item_ratings;       /* core/data.go:202-216 */
    int64_t *item_len;
    or_idrating *item_store;
};

/* core/data.go:131-154 */
or_trainset *or_trainset_new(const int64_t *users, const int64_t *items, const double *ratings, int64_t n) {
    or_trainset *t = (or_trainset *)calloc(1, sizeof(*t));
    t->n = n;
    t->users = (int64_t *)malloc((size_t)(n + 1) * sizeof(int64_t));
    t->items = (int64_t *)malloc((size_t)(n + 1) * sizeof(int64_t));
    t->ratings = (double *)malloc((size_t)(n + 1) * sizeof(double));
    memcpy(t->users, users, (size_t)n * sizeof(int64_t));
    memcpy(t->items, items, (size_t)n * sizeof(int64_t));
    memcpy(t->ratings, ratings, (size_t)n * sizeof(double));
    /* core/data.go:134 — stat.Mean(Ratings, nil): sum / len (see header for the order caveat) */
    double sum = 0.0;
    for (int64_t i = 0; i < n; i++) sum += ratings[i];
    t->global_mean = sum / (double)n;
    t->iu = (int32_t *)malloc((size_t)(n + 1) * sizeof(int32_t));
    t->ii = (int32_t *)malloc((size_t)(n + 1) * sizeof(int32_t));
    /* core/data.go:137-143 */
    idmap_init(&t->inner_users, (uint64_t)n);
    for (int64_t i = 0; i < n; i++) {
        int32_t v = idmap_get(&t->inner_users, users[i]);
        if (v < 0) { v = (int32_t)t->user_count; idmap_put(&t->inner_users, users[i], v); t->user_count++; }
        t->iu[i] = v;
    }
    /* core/data.go:145-151 */
    idmap_init(&t->inner_items, (uint64_t)n);
    for (int64_t i = 0; i < n; i++) {
        int32_t v = idmap_get(&t->inner_items, items[i]);
        if (v < 0) { v = (int32_t)t->item_count; idmap_put(&t->inner_items, items[i], v); t->item_count++; }
        t->ii[i] = v;
    }
    return t;
}

void or_trainset_free(or_trainset *t) {
    if (!t) return;
    free(t->users); free(t->items); free(t->ratings); free(t->iu); free(t->ii);
    idmap_free(&t->inner_users); idmap_free(&t->inner_items);
    free(t->user_ratings); free(t->user_len); free(t->user_store);
    free(t->item_ratings); free(t->item_len); free(t->item_store);
    free(t);
}

int64_t or_trainset_user_count(const or_trainset *t) { return t->user_count; }
int64_t or_trainset_item_count(const or_trainset *t) { return t->item_count; }
double or_trainset_global_mean(const or_trainset *t) { return t->global_mean; }
/* core/data.go:171-182 */
int64_t or_trainset_convert_user(const or_trainset *t, int64_t raw) { return idmap_get(&t->inner_users, raw); }
int64_t or_trainset_convert_item(const or_trainset *t, int64_t raw) { return idmap_get(&t->inner_items, raw); }
const int32_t *or_trainset_inner_users(const or_trainset *t) { return t->iu; }
const int32_t *or_trainset_inner_items(const or_trainset *t) { return t->ii; }

/* Adjacency lists appended in dataset order — core/data.go:185-199 / 202-216. */
static void build_adjacency(int64_t n, int64_t rows, const int32_t *row_of, const int32_t *col_of,
                            const double *ratings, or_idrating ***out_rows, int64_t **out_len,
                            or_idrating **out_store) {
    int64_t *len = (int64_t *)calloc((size_t)rows + 1, sizeof(int64_t));
    for (int64_t i = 0; i < n; i++) len[row_of[i]]++;
    or_idrating *store = (or_idrating *)malloc((size_t)(n + 1) * sizeof(or_idrating));
    or_idrating **r = (or_idrating **)malloc((size_t)(rows + 1) * sizeof(*r));
    int64_t off = 0;
    for (int64_t u = 0; u < rows; u++) { r[u] = store + off; off += len[u]; len[u] = 0; }
    for (int64_t i = 0; i < n; i++) {
        int32_t u = row_of[i];
        r[u][len[u]].id = col_of[i];
        r[u][len[u]].rating = ratings[i];
        len[u]++;
    }
    *out_rows = r; *out_len = len; *out_store = store;
}

static void trainset_user_ratings(or_trainset *t) {
    if (!t->user_ratings)
        build_adjacency(t->n, t->user_count, t->iu, t->ii, t->ratings, &t->user_ratings, &t->user_len, &t->user_store);
}
static void trainset_item_ratings(or_trainset *t) {
    if (!t->item_ratings)
        build_adjacency(t->n, t->item_count, t->ii, t->iu, t->ratings, &t->item_ratings, &t->item_len, &t->item_store);
}

/* =====================================================================================
 * BaseLine — core/base.go:122-163.  Strictly sequential SGD in dataset order.
 * ===================================================================================== */
void or_baseline_fit(const or_trainset *t, double reg, double lr, int n_epochs,
                     double *user_bias, double *item_bias, double *global_bias_out) {
    double global_bias = 0.0;
    for (int64_t u = 0; u < t->user_count; u++) user_bias[u] = 0.0;   /* core/base.go:142 */
    for (int64_t i = 0; i < t->item_count; i++) item_bias[i] = 0.0;   /* core/base.go:143 */
    for (int epoch = 0; epoch < n_epochs; epoch++) {                  /* core/base.go:145 */
        for (int64_t i = 0; i < t->n; i++) {
            double rating = t->ratings[i];
            int32_t iu = t->iu[i], ii = t->ii[i];
            double ub = user_bias[iu];
            double ib = item_bias[ii];
            /* core/base.go:122-134 — Predict: ret := globalBias; ret += userBias; ret += itemBias */
            double ret = global_bias;
            ret += user_bias[iu];
            ret += item_bias[ii];
            double diff = ret - rating;                               /* core/base.go:153 */
            double grad_global = diff;
            double grad_user = diff + reg * ub;
            double grad_item = diff + reg * ib;
            global_bias -= lr * grad_global;                          /* core/base.go:158-160 */
            user_bias[iu] -= lr * grad_user;
            item_bias[ii] -= lr * grad_item;
        }
    }
    if (global_bias_out) *global_bias_out = global_bias;
}

/* =====================================================================================
 * ALS baselines — EXTENSION (BASELINE.json config 3 names "ALS baselines"; the reference only
 * has the sequential SGD above, so this has no reference counterpart: PARITY UNPINNED, this
 * function is the definition the device kernel (csrc/baseline.cu) is checked against).
 * Koren's alternating least squares as popularised by Surprise's `baseline_only`:
 *     repeat n_epochs:  b_i = sum_{u in R(i)} (r_ui - mu - b_u) / (reg_i + |R(i)|)   for every item
 *                       b_u = sum_{i in R(u)} (r_ui - mu - b_i) / (reg_u + |R(u)|)   for every user
 * with mu = the global mean.  Summation order (fixed so the result is reproducible bit for bit):
 * a row's terms, in dataset order, are dealt round-robin to 32 partial sums (term t goes to
 * partial t % 32, each partial accumulated sequentially), then combined by the butterfly
 * p[l] += p[l ^ 16], ^8, ^4, ^2, ^1 — exactly what one warp does with __shfl_xor.
 * ===================================================================================== */
static double als_row_sum(const or_idrating *row, int64_t len, double mu, const double *other) {
    double part[32];
    for (int l = 0; l < 32; l++) part[l] = 0.0;
    for (int64_t t = 0; t < len; t++) {
        double term = (row[t].rating - mu) - other[row[t].id];
        part[t & 31] += term;
    }
    for (int o = 16; o > 0; o >>= 1) {
        double nxt[32];
        for (int l = 0; l < 32; l++) nxt[l] = part[l] + part[l ^ o];
        for (int l = 0; l < 32; l++) part[l] = nxt[l];
    }
    return part[0];
}

void or_baseline_als(or_trainset *t, double reg_u, double reg_i, int n_epochs, double *user_bias,
                     double *item_bias) {
    /* private adjacency in dataset order (KNN.Fit sorts the shared lists in place, core/data.go:236-243) */
    or_idrating **ur, **ir, *ustore, *istore;
    int64_t *ulen, *ilen;
    build_adjacency(t->n, t->user_count, t->iu, t->ii, t->ratings, &ur, &ulen, &ustore);
    build_adjacency(t->n, t->item_count, t->ii, t->iu, t->ratings, &ir, &ilen, &istore);
    const double mu = t->global_mean;
    for (int64_t u = 0; u < t->user_count; u++) user_bias[u] = 0.0;
    for (int64_t i = 0; i < t->item_count; i++) item_bias[i] = 0.0;
    for (int epoch = 0; epoch < n_epochs; epoch++) {
        for (int64_t i = 0; i < t->item_count; i++)
            item_bias[i] = als_row_sum(ir[i], ilen[i], mu, user_bias) / (reg_i + (double)ilen[i]);
        for (int64_t u = 0; u < t->user_count; u++)
            user_bias[u] = als_row_sum(ur[u], ulen[u], mu, item_bias) / (reg_u + (double)ulen[u]);
    }
    free(ur); free(ulen); free(ustore); free(ir); free(ilen); free(istore);
}

/* =====================================================================================
 * KNN — core/knn.go.
 * ===================================================================================== */
struct or_knn {
    or_params p;
    or_trainset *data;
    double global_mean;
    int64_t n;                       /* rows of Sims */
    double *sims;                    /* n*n, NaN = unset (core/utils.go:110-120) */
    or_idrating **left, **right;
    int64_t *left_len, *right_len;
    int64_t n_right;
    double *means, *stddevs, *bias;
    double *user_bias, *item_bias;   /* both kept for the PearsonBaseline extension */
    double global_bias;
};

void or_params_default(or_params *p) {
    p->sim = OR_SIM_MSD; p->knn_type = OR_KNN_BASIC; p->user_based = 1; p->k = 40; p->min_k = 1;
    p->n_jobs = 1; p->tie_policy = OR_TIE_CANONICAL; p->reg = 0.02; p->lr = 0.005; p->n_epochs = 20;
    p->shrinkage = 0.0;
    p->baseline_als = 0; p->als_epochs = 10; p->reg_u = 15.0; p->reg_i = 10.0;
}

or_knn *or_knn_new(const or_params *p) {
    or_knn *k = (or_knn *)calloc(1, sizeof(*k));
    k->p = *p;
    return k;
}

void or_knn_free(or_knn *k) {
    if (!k) return;
    free(k->sims); free(k->means); free(k->stddevs); free(k->user_bias); free(k->item_bias);
    free(k);
}

/* PearsonBaseline — EXTENSION, not in the reference.  Residuals
 * e = r - (globalBias + b_left + b_right) with the SGD baseline of core/base.go, cosine of
 * the residuals over the co-rated entries in ascending id, optional shrinkage
 * (n-1)/(n-1+shrinkage).  PARITY UNPINNED: this oracle is its only definition. */
static double sim_pearson_baseline(const or_knn *k, int64_t ia, int64_t ib) {
    const or_idrating *a = k->left[ia], *b = k->left[ib];
    int64_t na = k->left_len[ia], nb = k->left_len[ib];
    const double *lb = k->p.user_based ? k->user_bias : k->item_bias;
    const double *rb = k->p.user_based ? k->item_bias : k->user_bias;
    double m = .0, n = .0, l = .0, cnt = .0;
    int64_t ptr = 0;
    for (int64_t x = 0; x < na; x++) {
        while (ptr < nb && b[ptr].id < a[x].id) ptr++;
        if (ptr < nb && b[ptr].id == a[x].id) {
            double base_a = k->global_bias + lb[ia];
            base_a += rb[a[x].id];
            double base_b = k->global_bias + lb[ib];
            base_b += rb[a[x].id];
            double ea = a[x].rating - base_a;
            double eb = b[ptr].rating - base_b;
            m += ea * ea;
            n += eb * eb;
            l += ea * eb;
            cnt += 1;
        }
    }
    double rho = l / (sqrt(m) * sqrt(n));
    if (k->p.shrinkage > 0.0) rho = (cnt - 1.0) / (cnt - 1.0 + k->p.shrinkage) * rho;
    return rho;
}

static double knn_sim(const or_knn *k, int64_t i, int64_t j) {
    if (k->p.sim == OR_SIM_PEARSON_BASELINE) return sim_pearson_baseline(k, i, j);
    return or_sim_lists(k->p.sim, k->left[i], k->left_len[i], k->left[j], k->left_len[j]);
}

typedef struct { or_knn *k; int64_t begin, end; } fit_job;

static inline int is_nan_bits(const double *p) {
    uint64_t u = __atomic_load_n((const uint64_t *)p, __ATOMIC_RELAXED);
    double d; memcpy(&d, &u, 8);
    return isnan(d);
}
static inline void store_bits(double *p, double v) {
    uint64_t u; memcpy(&u, &v, 8);
    __atomic_store_n((uint64_t *)p, u, __ATOMIC_RELAXED);
}

/* core/knn.go:199-213 — the goroutine body. */
static void *fit_worker(void *arg) {
    fit_job *job = (fit_job *)arg;
    or_knn *k = job->k;
    int64_t n = k->n;
    for (int64_t i = job->begin; i < job->end; i++) {
        for (int64_t j = 0; j < n; j++) {
            if (i != j) {
                if (is_nan_bits(&k->sims[i * n + j])) {
                    double ret = knn_sim(k, i, j);
                    if (!isnan(ret)) {
                        store_bits(&k->sims[i * n + j], ret);
                        store_bits(&k->sims[j * n + i], ret);
                    }
                }
            }
        }
    }
    return NULL;
}

static void knn_fit_impl(or_knn *k, or_trainset *t, int64_t row0, int64_t row1, int slab) {
    int user_based = k->p.user_based;
    k->data = t;
    k->global_mean = t->global_mean;                                  /* core/knn.go:152 */
    trainset_user_ratings(t);
    trainset_item_ratings(t);
    if (user_based) {                                                 /* core/knn.go:154-162 */
        k->left = t->user_ratings; k->left_len = t->user_len;
        k->right = t->item_ratings; k->right_len = t->item_len;
        k->n = t->user_count; k->n_right = t->item_count;
    } else {
        k->left = t->item_ratings; k->left_len = t->item_len;
        k->right = t->user_ratings; k->right_len = t->user_len;
        k->n = t->item_count; k->n_right = t->user_count;
    }
    int64_t n = k->n;
    free(k->sims);
    k->sims = (double *)malloc((size_t)n * (size_t)n * sizeof(double) + 8);
    for (int64_t i = 0; i < n * n; i++) k->sims[i] = NAN;             /* core/utils.go:110-120 */
    free(k->means); free(k->stddevs); free(k->user_bias); free(k->item_bias);
    k->means = k->stddevs = k->bias = k->user_bias = k->item_bias = NULL;
    if (k->p.knn_type == OR_KNN_CENTERED || k->p.knn_type == OR_KNN_ZSCORE) {
        /* core/data.go:222-235 — means over the left rows in dataset order (before sorts) */
        k->means = (double *)malloc((size_t)(n + 1) * sizeof(double));
        for (int64_t i = 0; i < n; i++) {
            double sum = 0.0, count = 0.0;
            for (int64_t x = 0; x < k->left_len[i]; x++) { sum += k->left[i][x].rating; count++; }
            k->means[i] = sum / count;
        }
    }
    if (k->p.knn_type == OR_KNN_ZSCORE) {                             /* core/knn.go:167-177 */
        k->stddevs = (double *)malloc((size_t)(n + 1) * sizeof(double));
        for (int64_t i = 0; i < n; i++) {
            double sum = 0.0, count = 0.0;
            for (int64_t x = 0; x < k->left_len[i]; x++) {
                double r = k->left[i][x].rating;
                sum += (r - k->means[i]) * (r - k->means[i]);
                count++;
            }
            k->stddevs[i] = sqrt(sum / count) + 1e-5;
        }
    }
    if (k->p.knn_type == OR_KNN_BASELINE || k->p.sim == OR_SIM_PEARSON_BASELINE) {  /* core/knn.go:179-187 */
        k->user_bias = (double *)malloc((size_t)(t->user_count + 1) * sizeof(double));
        k->item_bias = (double *)malloc((size_t)(t->item_count + 1) * sizeof(double));
        if (k->p.baseline_als) {
            or_baseline_als(t, k->p.reg_u, k->p.reg_i, k->p.als_epochs, k->user_bias, k->item_bias);
            k->global_bias = t->global_mean;
        } else {
            or_baseline_fit(t, k->p.reg, k->p.lr, k->p.n_epochs, k->user_bias, k->item_bias, &k->global_bias);
        }
        if (k->p.knn_type == OR_KNN_BASELINE) k->bias = user_based ? k->user_bias : k->item_bias;
    }
    /* core/knn.go:190 → core/data.go:236-243: in-place sort.Sort of every left row by id */
    for (int64_t i = 0; i < n; i++) or_sort_by_id(k->left[i], k->left_len[i]);
    /* core/knn.go:192-216: nJobs goroutines over a static row split */
    int n_jobs = k->p.n_jobs < 1 ? 1 : k->p.n_jobs;
    int64_t length = slab ? (row1 - row0) : n;
    int64_t base = slab ? row0 : 0;
    pthread_t *th = (pthread_t *)malloc((size_t)n_jobs * sizeof(pthread_t));
    fit_job *jobs = (fit_job *)malloc((size_t)n_jobs * sizeof(fit_job));
    for (int j = 0; j < n_jobs; j++) {
        jobs[j].k = k;
        jobs[j].begin = base + length * j / n_jobs;
        jobs[j].end = base + length * (j + 1) / n_jobs;
        if (n_jobs == 1) fit_worker(&jobs[j]);
        else pthread_create(&th[j], NULL, fit_worker, &jobs[j]);
    }
    if (n_jobs > 1) for (int j = 0; j < n_jobs; j++) pthread_join(th[j], NULL);
    free(th); free(jobs);
}

void or_knn_fit(or_knn *k, or_trainset *t) { knn_fit_impl(k, t, 0, 0, 0); }
void or_knn_fit_rows(or_knn *k, or_trainset *t, int64_t row0, int64_t row1) { knn_fit_impl(k, t, row0, row1, 1); }

/* ---- CandidateSet as a sort.Interface — core/knn.go:28-48 ---- */
typedef struct { const double *sims; or_idrating *cand; } candset;
static int cand_less(void *ctx, int64_t i, int64_t j) {
    candset *c = (candset *)ctx;
    return c->sims[c->cand[i].id] > c->sims[c->cand[j].id];
}
static void cand_swap(void *ctx, int64_t i, int64_t j) {
    candset *c = (candset *)ctx;
    or_idrating t = c->cand[i]; c->cand[i] = c->cand[j]; c->cand[j] = t;
}
/* canonical policy: similarity desc, inner id asc (total order; ids in a right row are unique) */
static const double *g_canon_sims;  /* only used through the thread-local below */
static __thread const double *tl_canon_sims;
static int canon_cmp(const void *pa, const void *pb) {
    const or_idrating *a = (const or_idrating *)pa, *b = (const or_idrating *)pb;
    double sa = tl_canon_sims[a->id], sb = tl_canon_sims[b->id];
    if (sa > sb) return -1;
    if (sa < sb) return 1;
    return (a->id > b->id) - (a->id < b->id);
}

/* core/knn.go:75-141.  If nb_ids != NULL also reports the neighbours used. */
static double knn_predict_inner(const or_knn *k, int64_t inner_user, int64_t inner_item,
                                int64_t *nb_ids, double *nb_sims, int nb_cap, int *nb_count) {
    (void)g_canon_sims;
    if (nb_count) *nb_count = 0;
    int64_t left_id, right_id;
    if (k->p.user_based) { left_id = inner_user; right_id = inner_item; }
    else { left_id = inner_item; right_id = inner_user; }
    if (left_id == -1 || right_id == -1) return k->global_mean;      /* core/knn.go:89-91 */
    const double *row = k->sims + left_id * k->n;
    int64_t rn = k->right_len[right_id];
    or_idrating *cands = (or_idrating *)malloc((size_t)(rn + 1) * sizeof(or_idrating));
    int64_t nc = 0;
    for (int64_t x = 0; x < rn; x++) {                               /* core/knn.go:95-99 */
        or_idrating ir = k->right[right_id][x];
        if (!isnan(row[ir.id])) cands[nc++] = ir;
    }
    if (nc <= k->p.min_k) { free(cands); return k->global_mean; }    /* core/knn.go:102-104 */
    if (k->p.tie_policy == OR_TIE_GO) {                              /* core/knn.go:107-108 */
        candset cs = { row, cands };
        or_sort_iface s = { &cs, cand_less, cand_swap };
        or_go_sort(&s, nc);
    } else {
        tl_canon_sims = row;
        qsort(cands, (size_t)nc, sizeof(or_idrating), canon_cmp);
    }
    int64_t num = k->p.k;                                            /* core/knn.go:111-114 */
    if (num > nc) num = nc;
    double weight_sum = 0.0, weight_rating = 0.0;
    for (int64_t x = 0; x < num; x++) {                              /* core/knn.go:116-130 */
        or_idrating o = cands[x];
        weight_sum += row[o.id];
        double rating = o.rating;
        if (k->p.knn_type == OR_KNN_CENTERED) rating -= k->means[o.id];
        else if (k->p.knn_type == OR_KNN_ZSCORE) rating = (rating - k->means[o.id]) / k->stddevs[o.id];
        else if (k->p.knn_type == OR_KNN_BASELINE) rating -= k->bias[o.id];
        weight_rating += row[o.id] * rating;
        if (nb_ids && x < nb_cap) { nb_ids[x] = o.id; nb_sims[x] = row[o.id]; }
    }
    if (nb_count) *nb_count = (int)(num < nb_cap ? num : nb_cap);
    double prediction = weight_rating / weight_sum;                  /* core/knn.go:131-140 */
    if (k->p.knn_type == OR_KNN_CENTERED) prediction += k->means[left_id];
    else if (k->p.knn_type == OR_KNN_BASELINE) prediction += k->bias[left_id];
    else if (k->p.knn_type == OR_KNN_ZSCORE) { prediction *= k->stddevs[left_id]; prediction += k->means[left_id]; }
    free(cands);
    return prediction;
}

double or_knn_predict(const or_knn *k, int64_t raw_user, int64_t raw_item) {
    int64_t iu = or_trainset_convert_user(k->data, raw_user);        /* core/knn.go:76-77 */
    int64_t ii = or_trainset_convert_item(k->data, raw_item);
    return knn_predict_inner(k, iu, ii, NULL, NULL, 0, NULL);
}

int or_knn_predict_neighbors(const or_knn *k, int64_t raw_user, int64_t raw_item,
                             int64_t *ids, double *sims, int cap) {
    int64_t iu = or_trainset_convert_user(k->data, raw_user);
    int64_t ii = or_trainset_convert_item(k->data, raw_item);
    int cnt = 0;
    knn_predict_inner(k, iu, ii, ids, sims, cap, &cnt);
    return cnt;
}

typedef struct { const or_knn *k; const int64_t *u, *i; double *out; int64_t begin, end; } pred_job;
static void *pred_worker(void *arg) {
    pred_job *j = (pred_job *)arg;
    for (int64_t x = j->begin; x < j->end; x++) j->out[x] = or_knn_predict(j->k, j->u[x], j->i[x]);
    return NULL;
}

/* core/data.go:98-105 (n_threads == 1 is the reference's serial loop). */
void or_knn_predict_batch(const or_knn *k, const int64_t *users, const int64_t *items, int64_t n,
                          double *out, int n_threads) {
    if (n_threads < 1) n_threads = 1;
    pthread_t *th = (pthread_t *)malloc((size_t)n_threads * sizeof(pthread_t));
    pred_job *jobs = (pred_job *)malloc((size_t)n_threads * sizeof(pred_job));
    for (int j = 0; j < n_threads; j++) {
        jobs[j] = (pred_job){ k, users, items, out, n * j / n_threads, n * (j + 1) / n_threads };
        if (n_threads == 1) pred_worker(&jobs[j]);
        else pthread_create(&th[j], NULL, pred_worker, &jobs[j]);
    }
    if (n_threads > 1) for (int j = 0; j < n_threads; j++) pthread_join(th[j], NULL);
    free(th); free(jobs);
}

int64_t or_knn_n(const or_knn *k) { return k->n; }
const double *or_knn_sims_row(const or_knn *k, int64_t row) { return k->sims + row * k->n; }
const double *or_knn_means(const or_knn *k) { return k->means; }
const double *or_knn_stddevs(const or_knn *k) { return k->stddevs; }
const double *or_knn_bias(const or_knn *k) { return k->bias; }
double or_knn_global_mean(const or_knn *k) { return k->global_mean; }

typedef struct { double s; int32_t id; } topk_ent;
static int topk_cmp(const void *pa, const void *pb) {
    const topk_ent *a = (const topk_ent *)pa, *b = (const topk_ent *)pb;
    if (a->s > b->s) return -1;
    if (a->s < b->s) return 1;
    return (a->id > b->id) - (a->id < b->id);
}

void or_knn_topk(const or_knn *k, int kk, int64_t row0, int64_t row1, int32_t *idx, double *sim) {
    int64_t n = k->n;
    topk_ent *buf = (topk_ent *)malloc((size_t)(n + 1) * sizeof(topk_ent));
    for (int64_t r = row0; r < row1; r++) {
        const double *row = k->sims + r * n;
        int64_t c = 0;
        for (int64_t j = 0; j < n; j++) if (!isnan(row[j])) { buf[c].s = row[j]; buf[c].id = (int32_t)j; c++; }
        qsort(buf, (size_t)c, sizeof(topk_ent), topk_cmp);
        for (int x = 0; x < kk; x++) {
            int64_t o = (r - row0) * kk + x;
            if (x < c) { idx[o] = buf[x].id; sim[o] = buf[x].s; }
            else { idx[o] = -1; sim[o] = NAN; }
        }
    }
    free(buf);
}

void or_knn_pair_sums(const or_knn *k, int64_t ia, int64_t ib, int64_t out[6]) {
    const or_idrating *a = k->left[ia], *b = k->left[ib];
    int64_t na = k->left_len[ia], nb = k->left_len[ib];
    for (int i = 0; i < 6; i++) out[i] = 0;
    int64_t ptr = 0;
    for (int64_t x = 0; x < na; x++) {
        while (ptr < nb && b[ptr].id < a[x].id) ptr++;
        if (ptr < nb && b[ptr].id == a[x].id) {
            int64_t xa = (int64_t)a[x].rating, yb = (int64_t)b[ptr].rating;
            out[0] += 1; out[1] += xa; out[2] += yb; out[3] += xa * xa; out[4] += yb * yb; out[5] += xa * yb;
        }
    }
}

/* core/utils.go:162-180 with the intended ([]float64, []float64) signature */
double or_rmse(const double *pred, const double *truth, int64_t n) {
    double sum = 0.0;
    for (int64_t j = 0; j < n; j++) sum += (pred[j] - truth[j]) * (pred[j] - truth[j]);
    return sqrt(sum / (double)n);
}
double or_mae(const double *pred, const double *truth, int64_t n) {
    double sum = 0.0;
    for (int64_t j = 0; j < n; j++) sum += fabs(pred[j] - truth[j]);
    return sum / (double)n;
}

/* =====================================================================================
 * Slope One — core/slope_one.go (the step next to the KNN path; same co-rated
 * contraction, same Predict-style gather).  Restated op for op.
 * ===================================================================================== */
struct or_slope {
    or_trainset *data;
    double global_mean;
    int64_t n_items, n_users;
    or_idrating **user_ratings;      /* dataset order (core/slope_one.go:51) */
    int64_t *user_len;
    or_idrating *user_store;
    double *user_means;
    double *dev;                     /* n_items x n_items, zero-initialised (core/slope_one.go:53) */
};

typedef struct { double *dev; or_idrating **ir; int64_t *ilen; int64_t n, begin, end; } slope_job;
static void *slope_worker(void *arg) {
    slope_job *jb = (slope_job *)arg;
    double *dev = jb->dev;
    or_idrating **ir = jb->ir;
    int64_t *ilen = jb->ilen;
    const int64_t n = jb->n;
    for (int64_t i = jb->begin; i < jb->end; i++) {                    /* core/slope_one.go:71-90 */
        for (int64_t j = 0; j < i; j++) {
            double count = 0.0, sum = 0.0;
            int64_t ptr = 0;
            for (int64_t k = 0; k < ilen[i] && ptr < ilen[j]; k++) {
                or_idrating ur = ir[i][k];
                while (ptr < ilen[j] && ir[j][ptr].id < ur.id) ptr++;
                if (ptr < ilen[j] && ir[j][ptr].id == ur.id) {
                    count++;
                    sum += ur.rating - ir[j][ptr].rating;
                }
            }
            if (count > 0) {
                dev[i * n + j] = sum / count;
                dev[j * n + i] = -dev[i * n + j];
            }
        }
    }
    return NULL;
}

or_slope *or_slope_fit_rows(or_trainset *t, int n_jobs, int64_t row0, int64_t row1) {
    or_slope *s = (or_slope *)calloc(1, sizeof(*s));
    s->data = t;
    s->global_mean = t->global_mean;                                   /* core/slope_one.go:50 */
    s->n_items = t->item_count;
    s->n_users = t->user_count;
    /* private copies: the caller may already have fitted a KNN on this trainset, which sorts the
     * shared lists in place; Slope One reads userRatings in dataset order and sorts its own item lists */
    build_adjacency(t->n, t->user_count, t->iu, t->ii, t->ratings, &s->user_ratings, &s->user_len, &s->user_store);
    or_idrating **ir, *istore;
    int64_t *ilen;
    build_adjacency(t->n, t->item_count, t->ii, t->iu, t->ratings, &ir, &ilen, &istore);
    s->user_means = (double *)malloc((size_t)(s->n_users + 1) * sizeof(double));
    for (int64_t u = 0; u < s->n_users; u++) {                         /* core/data.go:222-235 means() */
        double sum = 0.0, count = 0.0;
        for (int64_t x = 0; x < s->user_len[u]; x++) { sum += s->user_ratings[u][x].rating; count++; }
        s->user_means[u] = sum / count;
    }
    const int64_t n = s->n_items;
    s->dev = (double *)calloc((size_t)n * (size_t)n + 1, sizeof(double));
    for (int64_t i = 0; i < n; i++) or_sort_by_id(ir[i], ilen[i]);     /* core/slope_one.go:55 sorts(itemRatings) */
    /* core/slope_one.go:59-92: nJobs goroutines over a static split of the rows */
    if (n_jobs < 1) n_jobs = 1;
    pthread_t *th = (pthread_t *)malloc((size_t)n_jobs * sizeof(pthread_t));
    slope_job *jobs = (slope_job *)malloc((size_t)n_jobs * sizeof(slope_job));
    for (int jb = 0; jb < n_jobs; jb++) {
        /* row0 < 0: all rows (the reference); otherwise only rows [row0,row1) x columns j < i (bench slabs) */
        const int64_t lo = row0 < 0 ? 0 : row0, len = (row0 < 0 ? n : row1) - lo;
        jobs[jb] = (slope_job){ s->dev, ir, ilen, n, lo + len * jb / n_jobs, lo + len * (jb + 1) / n_jobs };
        if (n_jobs == 1) slope_worker(&jobs[jb]);
        else pthread_create(&th[jb], NULL, slope_worker, &jobs[jb]);
    }
    if (n_jobs > 1) for (int jb = 0; jb < n_jobs; jb++) pthread_join(th[jb], NULL);
    free(th); free(jobs);
    free(ir); free(ilen); free(istore);
    return s;
}

or_slope *or_slope_fit(or_trainset *t, int n_jobs) { return or_slope_fit_rows(t, n_jobs, -1, -1); }

void or_slope_free(or_slope *s) {
    if (!s) return;
    free(s->user_ratings); free(s->user_len); free(s->user_store); free(s->user_means); free(s->dev);
    free(s);
}

int64_t or_slope_n(const or_slope *s) { return s->n_items; }
const double *or_slope_dev(const or_slope *s) { return s->dev; }
const double *or_slope_user_means(const or_slope *s) { return s->user_means; }

double or_slope_predict(const or_slope *s, int64_t raw_user, int64_t raw_item) {   /* core/slope_one.go:22-45 */
    int64_t iu = or_trainset_convert_user(s->data, raw_user);
    int64_t ii = or_trainset_convert_item(s->data, raw_item);
    double prediction = 0.0;
    if (iu >= 0) prediction = s->user_means[iu];                       /* newID = -1, core/data.go:129 */
    else prediction = s->global_mean;
    if (ii >= 0 && iu >= 0) {
        double sum = 0.0, count = 0.0;
        for (int64_t x = 0; x < s->user_len[iu]; x++) {
            sum += s->dev[ii * s->n_items + s->user_ratings[iu][x].id];
            count++;
        }
        if (count > 0) prediction += sum / count;
    }
    return prediction;
}

void or_slope_predict_batch(const or_slope *s, const int64_t *users, const int64_t *items, int64_t n, double *out) {
    for (int64_t x = 0; x < n; x++) out[x] = or_slope_predict(s, users[x], items[x]);
}

/* =====================================================================================
 * Test helper for shapes whose N x N matrix does not fit in host memory (BASELINE config 4:
 * 138,493 users): the similarities of a handful of left rows against all N, straight from
 * core/sim.go on lists sorted by id (core/knn.go:190), with the diagonal and "no co-rating"
 * cells NaN as KNN.Fit leaves them.  Cosine / MSD / Pearson only.
 * ===================================================================================== */
void or_rows_sims(or_trainset *t, int sim, int user_based, const int64_t *rows, int64_t n_rows, double *out) {
    or_idrating **lr, *store;
    int64_t *len;
    const int64_t n = user_based ? t->user_count : t->item_count;
    if (user_based) build_adjacency(t->n, t->user_count, t->iu, t->ii, t->ratings, &lr, &len, &store);
    else build_adjacency(t->n, t->item_count, t->ii, t->iu, t->ratings, &lr, &len, &store);
    for (int64_t i = 0; i < n; i++) or_sort_by_id(lr[i], len[i]);
    for (int64_t r = 0; r < n_rows; r++) {
        const int64_t i = rows[r];
        for (int64_t j = 0; j < n; j++)
            out[r * n + j] = (i == j) ? NAN : or_sim_lists(sim, lr[i], len[i], lr[j], len[j]);
    }
    free(lr); free(len); free(store);
}

