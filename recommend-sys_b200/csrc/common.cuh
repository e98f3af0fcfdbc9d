// common.cuh — shared declarations of librs_knn_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <mutex>
#include <vector>

#include "rs_knn.h"

void rs_set_error(const char *fmt, ...);

#define RS_CUDA(expr)                                                                       \
    do {                                                                                    \
        cudaError_t e_ = (expr);                                                            \
        if (e_ != cudaSuccess) {                                                            \
            rs_set_error("%s: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return (e_ == cudaErrorMemoryAllocation) ? RS_ERR_OOM : RS_ERR_CUDA;            \
        }                                                                                   \
    } while (0)

#define RS_TRY(expr)              \
    do {                          \
        int32_t rc_ = (expr);     \
        if (rc_ != RS_OK) return rc_; \
    } while (0)

// Rating classes (how a rating is represented as one byte; 0 = missing).
//   RS_CLASS_INT8 : every rating is an integer in [-11,11]; code = rating + 12.
//   RS_CLASS_TABLE: any other float64 ratings; no byte code (stream path only, reads the values).
enum { RS_CLASS_INT8 = 0, RS_CLASS_TABLE = 1 };
constexpr int RS_INT8_BIAS = 12;

// Column-chunk width of the streaming similarity kernel (threads * bytes per thread).
// Streaming similarity kernel: warps per CTA, each an independent (row, column-chunk) work item.
// The chunk width (rs_knn::stream_jc, 128 or 256 columns = 3 x JC doubles of accumulators per
// warp) is chosen per Fit: 128 gives small problems enough work items to balance the SMs
// (MovieLens-1M shape: 1.6 ms vs 2.2 ms), 256 halves the per-chunk walk of the row on large sparse
// ones (MovieLens-20M item shape: 85 ms vs 128 ms); profiles/r01_stream_notes.md.
constexpr int RS_STREAM_WARPS = 8;

// Tensor-core similarity kernel tile: 128 left rows (MMA M) x 64 left rows (MMA N),
// K blocked by 128 bytes (one SWIZZLE_128B atom) per pipeline stage.
constexpr int RS_TC_BM = 128;
constexpr int RS_TC_BN = 64;
constexpr int RS_TC_BK = 128;
constexpr int RS_TC_KBLK = 256;   // bytes of K per block of the K-blocked plane layout [plane][K/256][row][256]

// Cyclic row shards (RS_STORE_MATRIX with shard_count >= 2): rows are dealt to the shards in blocks of
// RS_CYC_B (block b belongs to shard b % count), so every shard gets the same mix of long and short
// rows; a shard stores its rows densely in dealing order.
constexpr int RS_CYC_B = 32;
constexpr int RS_MAX_PEERS = 16;
__host__ __device__ inline int64_t rs_cyc_local(int64_t i, int count) {
    return (i / ((int64_t)RS_CYC_B * count)) * RS_CYC_B + i % RS_CYC_B;
}
__host__ __device__ inline bool rs_cyc_owns(int64_t i, int count, int index) {
    return (i / RS_CYC_B) % count == index;
}
inline int64_t rs_cyc_rows(int64_t n, int count, int index) {
    const int64_t nblk = (n + RS_CYC_B - 1) / RS_CYC_B;
    int64_t rows = 0;
    for (int64_t b = index; b < nblk; b += count) rows += (b + 1) * RS_CYC_B <= n ? RS_CYC_B : n - b * RS_CYC_B;
    return rows;
}

struct rs_knn {
    std::recursive_mutex mu;             // serialises the ABI calls on this handle (api.cu Guard)
    rs_knn_params p{};
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaStream_t aux_stream = nullptr;   // read-backs of Fit statistics that must not wait for the similarity kernel
    cudaEvent_t ev_in = nullptr;         // the host inputs of rs_knn_fit have been consumed
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;   // heavy-row kernel on aux_stream beside the column walk
    cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_c = nullptr, ev_d = nullptr, ev_e = nullptr;
    rs_knn_profile prof{};
    // pending event pairs whose elapsed time has not been folded into prof yet
    bool sim_pending = false, pred_pending = false, prep_pending = false;

    bool fitted = false;
    int32_t n_left = 0, n_right = 0;
    int64_t nnz = 0;
    int64_t row_begin = 0, row_end = 0;  // resolved shard
    int64_t col_begin = 0;               // symmetric slabs: similarity tiles left of this column are not needed
    bool force_sym = false;              // symmetric slabs: compute only j > i although the rows are a slab
    int64_t topk_rows = 0;               // rows covered by topk_idx / topk_sim
    // cyclic row shards: cyc_R >= 2 shards, this handle is shard cyc_r and stores rows_local rows
    int32_t cyc_R = 0, cyc_r = 0;
    int64_t rows_local = 0;              // rows of `sims` (== row_end - row_begin unless cyclic)
    int64_t n_work_rows = 0;             // entries of row_order the stream kernel walks
    const double *peer_sims[RS_MAX_PEERS] = {nullptr};   // the shards' matrices (peer memory over NVLink), [cyc_r] = own
    bool peers_ready = false;
    double global_mean = 0.0, global_bias = 0.0;
    int rating_class = RS_CLASS_INT8;

    // Grow-only device arena: Fit bump-allocates from it and the next Fit reuses the same
    // chunks, so a refit of the same shape performs no cudaMalloc / cudaFree at all.
    struct Chunk { char *p; size_t bytes; };
    std::vector<Chunk> chunks;
    size_t cur_chunk = 0, cur_off = 0;
    // cached tensor-path tile list
    void *tile_buf = nullptr;
    size_t tile_buf_bytes = 0;
    int64_t tile_key[5] = {-1, -1, -1, -1, -1};
    int32_t tile_count = 0;
    // fused top-k: band tile lists (all waves, concatenated) and the per-row selection state
    void *band_buf = nullptr;
    size_t band_buf_bytes = 0;
    int64_t band_key[4] = {-1, -1, -1, -1};
    std::vector<int64_t> band_off;
    unsigned long long *thr_key = nullptr;   // [n_left] k-th best key of the row's running list (0 = not full)
    int32_t *thr_id = nullptr;
    int32_t *cand_cnt = nullptr;             // [n_left] candidates appended since the last merge
    int32_t *cand_id = nullptr;              // [n_left][cand_cap]
    double *cand_sim = nullptr;
    int32_t cand_cap = 0;
    int32_t *row_flag = nullptr;             // [n_left] the row's buffer overflowed in the current band

    // CSR of the left rows, entries ascending by right id (core/data.go:236-243)
    int64_t *l_ptr = nullptr;
    int32_t *l_col = nullptr;
    double *l_val = nullptr;
    uint8_t *l_code = nullptr;
    double *ld_val = nullptr;  // left rows in DATASET order (means / std accumulate in it)
    // CSR of the right rows, entries ascending by left id (Predict candidates)
    int64_t *r_ptr = nullptr;
    int32_t *r_col = nullptr;
    double *r_val = nullptr;

    double *means = nullptr;      // KNN.Means    (dataset order sum / count)
    double *stddevs = nullptr;    // KNN.StdDevs
    double *pmeans = nullptr;     // Pearson's own row means (sorted-order sum, core/sim.go:49-62)
    double *left_bias = nullptr;  // KNN.Bias
    int32_t *row_cnt = nullptr;   // ratings per left row and their integer sum (tensor path, Pearson)
    int32_t *row_sum = nullptr;
    int64_t max_row_cnt = 0;
    int32_t max_right_len = 0;    // longest right row = most candidates one prediction can have
    double triples = 0.0;         // co-rated triples of the full matrix: sum over right rows of cnt*(cnt-1)/2
    double *right_bias = nullptr;
    // Slope One: the right rows (users) in DATASET order — left ids only — and their means
    int32_t *rd_col = nullptr;
    double *right_means = nullptr;

    // stream path: b-side term of every rating in right-CSR order (value, value - row mean, ...),
    // chunk pointers cp[right][Q+1] into each right row, and l2r: left-CSR entry -> right-CSR index
    double *r_dev = nullptr;
    int32_t *cp = nullptr;
    int32_t n_chunks = 0;
    int32_t stream_jc = 256;
    bool stream_lower = false;     // full-matrix stream Fit computes j < i (else j > i); chosen per Fit from the lookup counts
    double inc_upper = 0.0, inc_lower = 0.0;   // (entry, chunk) lookups of the two triangles
    int64_t *l2r = nullptr;
    int32_t *perm_lr = nullptr, *perm_rl = nullptr, *perm_tmp = nullptr;  // CSR position -> input row (arena, valid until the next Fit)
    int32_t *row_order = nullptr;  // left rows sorted by descending length
    int32_t *row_heavy = nullptr;  // heavy rows split off that order (sim_stream.cu: sim_stream_heavy_kernel)
    int32_t n_heavy = 0;
    // popular columns (sim_stream.cu: sim_pop_kernel): the n_pop longest rows, longest first.  Their ratings are
    // taken out of the right CSR the column walk reads (walk CSR: w_ptr / w_col / w_dev, `cp` and `l2r` refer to
    // it) and kept as a dense table pop_dense[n_right][pop_ld] instead (one byte per cell when every rating is a
    // small integer, else the b-side double; 0 / NaN = no rating).
    int32_t n_pop = 0, pop_ld = 0;
    bool pop_u8 = false;
    int32_t *pop_idx = nullptr;    // [n_left] index in the popular list or -1
    int32_t *pop_items = nullptr;  // [pop_ld] row id (-1 beyond n_pop)
    uint8_t *pop_blk = nullptr;    // [ceil(n_left / 32)] the block of 32 rows holds a popular row
    void *pop_dense = nullptr;
    int64_t *w_ptr = nullptr;
    int32_t *w_col = nullptr;
    double *w_dev = nullptr;
    int32_t *row_all = nullptr;    // every row of the shard, longest first (the dense pass visits them all)
    int64_t n_all_rows = 0;
    // int8 planes X^2, M, X of the left matrix, [3][k_pad / 256][n_pad][256] (tensor path)
    int8_t *planes = nullptr;
    int64_t tc_npad = 0, tc_kpad = 0;

    // outputs
    double *sims = nullptr;  // (row_end-row_begin) x ld_s, NaN = unset
    int64_t ld_s = 0;
    int32_t *topk_idx = nullptr;  // RS_STORE_TOPK: rows x topk
    double *topk_sim = nullptr;

    int32_t *d_flags = nullptr;  // small device scratch for validation flags
    void *ovf = nullptr;         // predict: overflow list (count + indices)
    size_t ovf_bytes = 0;
    std::vector<void *> scratch;         // device staging of the host-pointer entry points
    std::vector<size_t> scratch_bytes;
};

// ---- devmem.cu: process-wide cache of device allocations ----
// Estimator copies are created and destroyed per cross-validation fold (core/eval.go:29-35);
// freed device blocks are parked here and handed to the next handle on the same device, so a
// create/Fit/Predict/destroy cycle performs no cudaMalloc/cudaFree after the first one.
int32_t rs_cached_malloc(int device, void **out, size_t bytes, size_t *got);
void rs_cached_free(int device, void *p, size_t bytes);
void rs_cache_trim(void);

// ---- api.cu: grow-only per-handle device scratch (slot-indexed) ----
extern "C" int32_t rs_scratch_get(rs_knn *h, int slot, size_t bytes, void **out);

// ---- prep.cu ----
int32_t rs_prep_build(rs_knn *h, const int32_t *d_left, const int32_t *d_right, const double *d_rating,
                      const double *d_left_bias, const double *d_right_bias);
int32_t rs_dev_alloc(rs_knn *h, void **out, size_t bytes);
template <typename T>
inline int32_t rs_alloc(rs_knn *h, T **out, size_t count) {
    return rs_dev_alloc(h, reinterpret_cast<void **>(out), count * sizeof(T));
}
int32_t rs_prep_rt(rs_knn *h);
int32_t rs_prep_planes(rs_knn *h);

// ---- sim_stream.cu ----
int32_t rs_sim_stream_launch(rs_knn *h);
int32_t rs_symmetrize_launch(rs_knn *h);
int32_t rs_mirror_launch(rs_knn *h);

// ---- sim_tensor.cu ----
int32_t rs_sim_tensor_launch(rs_knn *h, int32_t *d_cosums, int64_t cos_row0, int64_t cos_nrows);
int32_t rs_tensor_band_count(const rs_knn *h);
int32_t rs_sim_tensor_band_launch(rs_knn *h, int32_t wave);
bool rs_tensor_topk_fused(const rs_knn *h);

// ---- predict.cu ----
int32_t rs_predict_launch(rs_knn *h, const int32_t *d_left, const int32_t *d_right, int64_t n, double *d_out,
                          int32_t *d_nb_ids, double *d_nb_sims, int32_t *d_nb_count, int32_t nb_cap, int32_t foreign_zero = 0);
int32_t rs_topk_launch(rs_knn *h, int32_t k, int32_t *d_idx, double *d_sim);
int32_t rs_slope_predict_launch(rs_knn *h, const int32_t *d_left, const int32_t *d_right, int64_t n, double *d_out);
int32_t rs_topk_slab_launch(rs_knn *h, int64_t g0, int32_t m, double *tbuf, int64_t ld_t, int32_t k);
int32_t rs_topk_compact_launch(rs_knn *h, int32_t k, int32_t *d_overflow, int rerun);
int32_t rs_topk_mask_launch(rs_knn *h, int32_t k, int restore);

// order-preserving map double -> uint64 (larger similarity = larger key); -0.0 folded to +0.0
__host__ __device__ inline uint64_t rs_sim_key(double s) {
    s = s + 0.0;  // -0.0 -> +0.0, every other value unchanged
#ifdef __CUDA_ARCH__
    uint64_t b = (uint64_t)__double_as_longlong(s);
#else
    uint64_t b;
    __builtin_memcpy(&b, &s, 8);
#endif
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__host__ __device__ inline double rs_key_sim(uint64_t k) {
    uint64_t b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)b);
#else
    double s;
    __builtin_memcpy(&s, &b, 8);
    return s;
#endif
}
