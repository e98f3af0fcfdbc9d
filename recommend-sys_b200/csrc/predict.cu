// predict.cu — batched KNN.Predict (core/knn.go:75-141) and per-row top-k.
//
// Predict: one warp per (left,right) test pair.
//   1. candidates = entries of the right row (core/knn.go:95-99), similarity gathered from
//      the left row of the HBM-resident matrix; NaN filtered; count <= mink -> GlobalMean.
//   2. the candidates are ordered by (similarity desc, inner id asc) — the canonical tie
//      policy — with a warp-level bitonic sort over (key,pos) records in shared memory
//      (right rows are id-sorted, so position order == id order);
//   3. the first min(k,count) are accumulated SEQUENTIALLY in that order exactly as
//      core/knn.go:116-130 does, so the prediction is bit-identical to the restated
//      reference under the canonical policy.
// HBM-bound by design: per prediction C*(4 B id + 8 B rating + 8 B similarity [+ 8 B mean/bias]).
#include <cstdlib>

#include "common.cuh"

namespace {

constexpr int PRED_WARPS = 4;
constexpr int PRED_CAP = 512;   // (key,pos) records per warp

struct Rec {
    uint64_t key;
    uint32_t pos;
};

// a ranks before b: larger key first, then smaller position (= smaller id)
__device__ __forceinline__ bool rec_before(uint64_t ka, uint32_t pa, uint64_t kb, uint32_t pb) {
    return ka > kb || (ka == kb && pa < pb);
}

// Bitonic sort of n_pad (power of two, >= 32) records into "before" order by one warp.
__device__ void warp_bitonic(uint64_t *keys, uint32_t *pos, int n_pad, int lane) {
    for (int size = 2; size <= n_pad; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncwarp();
            for (int t = lane; t < (n_pad >> 1); t += 32) {
                int lo = 2 * t - (t & (stride - 1));
                int hi = lo + stride;
                bool up = ((lo & size) == 0);  // this run sorts into "before" order
                uint64_t ka = keys[lo], kb = keys[hi];
                uint32_t pa = pos[lo], pb = pos[hi];
                bool swap = up ? rec_before(kb, pb, ka, pa) : rec_before(ka, pa, kb, pb);
                if (swap) { keys[lo] = kb; keys[hi] = ka; pos[lo] = pb; pos[hi] = pa; }
            }
        }
    }
    __syncwarp();
}


struct PredArgs {
    const int32_t *left, *right;
    int64_t n;
    double *out;
    const int64_t *r_ptr;
    const int32_t *r_col;
    const double *r_val;
    const double *sims;
    int64_t ld_s;
    int64_t row_begin, row_end;
    const double *means, *stddevs, *bias;
    double global_mean;
    int32_t n_right;
    int32_t k, min_k, knn_type;
    int32_t *nb_ids;
    double *nb_sims;
    int32_t *nb_count;
    int32_t nb_cap;
};

// =====================================================================================
// Main predict kernel: selection, not sorting.  One warp per prediction:
//   pass A  gather the similarity of every candidate, count the non-NaN ones (core/knn.go:95-99)
//           and find their min / max;
//   pass B  (only when more candidates than the register capacity CAP = 32*R) 256-bucket
//           histogram of a monotone linear bucket of the similarity in shared memory, suffix
//           scan -> boundary bucket: everything above it is certainly in the top k, everything
//           below certainly not; refined (up to 3 levels) inside the boundary bucket when it is
//           too full; genuine ties are taken in scan order = ascending inner id (canonical);
//   pass C  compaction of the surviving <= CAP candidates into registers (R per lane);
//   sort    one warp-wide bitonic network over CAP (key,pos) records in registers
//           (strides < R in-register, >= R via __shfl_xor);
//   reduce  the first min(k,valid) neighbours are accumulated sequentially in sorted order,
//           exactly as core/knn.go:116-130 does.
// A prediction whose boundary bucket still overflows after 3 levels is appended to an
// overflow list and finished by the generic shared-memory kernel below.
// =====================================================================================
constexpr int SEL_WARPS = 8;

template <int R>
__device__ __forceinline__ void ce_regs(uint64_t (&key)[R], uint32_t (&pos)[R], int i, int j, bool up) {
    const bool j_before_i = rec_before(key[j], pos[j], key[i], pos[i]);
    if (j_before_i == up) {
        const uint64_t tk = key[i]; key[i] = key[j]; key[j] = tk;
        const uint32_t tp = pos[i]; pos[i] = pos[j]; pos[j] = tp;
    }
}

template <int R, int STRIDE>
__device__ __forceinline__ void stage_regs(uint64_t (&key)[R], uint32_t (&pos)[R], int lane, int size) {
#pragma unroll
    for (int x = 0; x < R; x++) {
        if ((x & STRIDE) == 0) {
            const bool up = (((lane * R + x) & size) == 0);
            ce_regs<R>(key, pos, x, x | STRIDE, up);
        }
    }
}

// blocked layout: element e = lane*R + x
template <int R>
__device__ __forceinline__ void warp_sort_regs(uint64_t (&key)[R], uint32_t (&pos)[R], int lane) {
#pragma unroll 1
    for (int size = 2; size <= 32 * R; size <<= 1) {
#pragma unroll 1
        for (int stride = size >> 1; stride >= R; stride >>= 1) {
            const int lm = stride / R;
            const bool lower = (lane & lm) == 0;
#pragma unroll
            for (int x = 0; x < R; x++) {
                const bool up = (((lane * R + x) & size) == 0);
                const uint64_t ok = __shfl_xor_sync(0xffffffffu, key[x], lm);
                const uint32_t op = __shfl_xor_sync(0xffffffffu, pos[x], lm);
                const bool other_before = rec_before(ok, op, key[x], pos[x]);
                if (other_before == (up == lower)) { key[x] = ok; pos[x] = op; }
            }
        }
        if (R >= 16 && size >= 16) stage_regs<R, (R >= 16 ? 8 : 0)>(key, pos, lane, size);
        if (R >= 8 && size >= 8) stage_regs<R, (R >= 8 ? 4 : 0)>(key, pos, lane, size);
        if (R >= 4 && size >= 4) stage_regs<R, (R >= 4 ? 2 : 0)>(key, pos, lane, size);
        if (R >= 2) stage_regs<R, (R >= 2 ? 1 : 0)>(key, pos, lane, size);
    }
}

__device__ __forceinline__ int sel_bucket(double s, double lo, double scale) {
    int b = (int)((s - lo) * scale);   // monotone non-decreasing in s
    return b > 255 ? 255 : (b < 0 ? 0 : b);
}

template <int R>
__global__ void __launch_bounds__(SEL_WARPS * 32) predict_select_kernel(PredArgs a, int32_t *overflow_list,
                                                                        int32_t *overflow_count, int scap) {
    constexpr int CAP = 32 * R;
    // per-warp staging of the gathered similarities: pass A gathers each candidate's similarity
    // from HBM/L2 once (4 independent gathers per lane in flight) and parks the first `scap` of
    // them here; the selection passes read them back instead of gathering again
    extern __shared__ double s_stage[];
    __shared__ uint32_t s_hist[SEL_WARPS][256];
    __shared__ uint64_t s_key[SEL_WARPS][CAP];
    __shared__ uint32_t s_pos[SEL_WARPS][CAP];
    __shared__ double s_av[SEL_WARPS][CAP];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t *hist = s_hist[warp];
    uint64_t *ckey = s_key[warp];
    uint32_t *cpos = s_pos[warp];
    double *sbuf = s_stage + (size_t)warp * scap;
    const int64_t n_warps = (int64_t)gridDim.x * SEL_WARPS;
    const double nan_v = __longlong_as_double(0x7ff8000000000001ll);
    const uint32_t lt_mask = (1u << lane) - 1u;

    for (int64_t p = (int64_t)blockIdx.x * SEL_WARPS + warp; p < a.n; p += n_warps) {
        const int32_t l = a.left[p], r = a.right[p];
        if (a.nb_count && lane == 0) *a.nb_count = 0;
        if (l < 0 || r < 0 || r >= a.n_right) {            // core/knn.go:89-91 (newID)
            if (lane == 0) a.out[p] = a.global_mean;
            continue;
        }
        if (l < a.row_begin || l >= a.row_end) {           // not in this shard
            if (lane == 0) a.out[p] = nan_v;
            continue;
        }
        const double *row = a.sims + (l - a.row_begin) * a.ld_s;
        const int64_t cb = a.r_ptr[r];
        const int cnt = (int)(a.r_ptr[r + 1] - cb);        // a right row has at most n_left entries
        const int32_t *ids = a.r_col + cb;

        // ---- pass A: gather once, count + range ----
        double lo = __longlong_as_double(0x7ff0000000000000ll), hi = -lo;
        int valid = 0;
        __syncwarp();   // the previous prediction has finished reading sbuf
        for (int e0 = 0; e0 < cnt; e0 += 128) {
            int32_t idv[4];
            double sv4[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int e = e0 + u * 32 + lane;
                idv[u] = e < cnt ? ids[e] : -1;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) sv4[u] = idv[u] >= 0 ? row[idv[u]] : nan_v;
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int e = e0 + u * 32 + lane;
                const double s = sv4[u];
                if (e < scap) sbuf[e] = s;
                if (s == s) { valid++; lo = fmin(lo, s); hi = fmax(hi, s); }
            }
        }
        __syncwarp();
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            valid += __shfl_xor_sync(0xffffffffu, valid, o);
            lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if (valid <= a.min_k) {                            // core/knn.go:102-104 (note <=)
            if (lane == 0) a.out[p] = a.global_mean;
            continue;
        }
        const int num = a.k < valid ? a.k : valid;         // core/knn.go:111-114

        // ---- pass B: narrow to <= CAP candidates that contain the top `num` ----
        // selected  <=>  s > hi_sure  ||  (lo <= s <= hi && bucket(s) >= T)   [+ scan-order ties]
        double hi_sure = __longlong_as_double(0x7ff0000000000000ll);  // +inf: nothing above yet
        double scale = 0.0;
        int T = 0;
        int sure = 0;          // candidates already known to be in the top `num`
        int tie_take = -1;     // >= 0: the interval is one value; take this many in scan order
        bool overflow = false;
        if (valid > CAP) {
            for (int level = 0;; level++) {
                for (int x = lane; x < 256; x += 32) hist[x] = 0;
                __syncwarp();
                scale = (hi > lo) ? 256.0 / (hi - lo) : 0.0;
                for (int e = lane; e < cnt; e += 32) {
                    const double s = e < scap ? sbuf[e] : row[ids[e]];
                    if (s >= lo && s <= hi) atomicAdd(&hist[sel_bucket(s, lo, scale)], 1u);
                }
                __syncwarp();
                // suffix counts: lane owns buckets [8*lane, 8*lane+8)
                uint32_t h[8], mine = 0;
#pragma unroll
                for (int x = 0; x < 8; x++) { h[x] = hist[8 * lane + x]; mine += h[x]; }
                uint32_t above = mine;   // inclusive suffix over lanes
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t v = __shfl_down_sync(0xffffffffu, above, o);
                    if (lane + o < 32) above += v;
                }
                above -= mine;           // candidates in buckets of higher lanes
                const int need = num - sure;
                // the boundary bucket is the highest T with count(bucket >= T) >= need
                int myT = -1;
                uint32_t run = above, sure_here = 0, bd_here = 0;
#pragma unroll
                for (int x = 7; x >= 0; x--) {
                    if (myT < 0 && run + h[x] >= (uint32_t)need) { myT = 8 * lane + x; sure_here = run; bd_here = h[x]; }
                    run += h[x];
                }
                const uint32_t has = __ballot_sync(0xffffffffu, myT >= 0);
                const int src = 31 - __clz(has);           // highest lane that found it
                T = __shfl_sync(0xffffffffu, myT, src);
                const int sure_lvl = (int)__shfl_sync(0xffffffffu, sure_here, src);
                const int bd = (int)__shfl_sync(0xffffffffu, bd_here, src);
                if (sure + sure_lvl + bd <= CAP) break;    // compaction fits
                // too many in the boundary bucket: refine inside it
                double lo2 = __longlong_as_double(0x7ff0000000000000ll), hi2 = -lo2;
                for (int e = lane; e < cnt; e += 32) {
                    const double s = e < scap ? sbuf[e] : row[ids[e]];
                    if (s >= lo && s <= hi && sel_bucket(s, lo, scale) == T) { lo2 = fmin(lo2, s); hi2 = fmax(hi2, s); }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    lo2 = fmin(lo2, __shfl_xor_sync(0xffffffffu, lo2, o));
                    hi2 = fmax(hi2, __shfl_xor_sync(0xffffffffu, hi2, o));
                }
                sure += sure_lvl;
                hi_sure = hi2;            // everything above the boundary bucket is certain
                lo = lo2; hi = hi2; T = 0; scale = 0.0;
                if (lo2 == hi2) { tie_take = num - sure; break; }   // genuine ties: scan order
                if (level == 2) { overflow = true; break; }
            }
        }
        const bool take_all = valid <= CAP;   // every valid candidate fits in the registers
        if (overflow) {
            if (lane == 0) overflow_list[atomicAdd(overflow_count, 1)] = (int32_t)p;
            continue;
        }

        // ---- pass C: compaction (scan order = ascending inner id) ----
        int have = 0, ties_taken = 0;
        for (int base = 0; base < cnt; base += 32) {
            const int e = base + lane;
            bool take = false, tie = false;
            uint64_t key = 0;
            if (e < cnt) {
                const double s = e < scap ? sbuf[e] : row[ids[e]];
                if (s == s) {
                    key = rs_sim_key(s);
                    if (take_all || s > hi_sure) take = true;
                    else if (s >= lo && s <= hi) {
                        if (tie_take >= 0) tie = true;
                        else take = sel_bucket(s, lo, scale) >= T;
                    }
                }
            }
            if (tie_take >= 0) {
                const uint32_t tm = __ballot_sync(0xffffffffu, tie);
                if (tie && ties_taken + __popc(tm & lt_mask) < tie_take) take = true;
                ties_taken += __popc(tm);
            }
            const uint32_t m = __ballot_sync(0xffffffffu, take);
            if (take) {
                const int slot = have + __popc(m & lt_mask);
                ckey[slot] = key;
                cpos[slot] = (uint32_t)e;
            }
            have += __popc(m);
        }
        __syncwarp();

        // ---- sort the survivors in registers ----
        uint64_t key[R];
        uint32_t pos[R];
#pragma unroll
        for (int x = 0; x < R; x++) {
            const int e = lane * R + x;
            key[x] = e < have ? ckey[e] : 0ull;
            pos[x] = e < have ? cpos[e] : 0xffffffffu;
        }
        __syncwarp();
        warp_sort_regs<R>(key, pos, lane);

        // ---- weighted mean over the first `num`, sequential in sorted order ----
        // the sorted (similarity, adjusted rating) pairs go through shared memory so that the
        // serial accumulation is two broadcast loads and three FP64 operations per neighbour
        double *w_s = reinterpret_cast<double *>(ckey), *w_a = s_av[warp];
#pragma unroll
        for (int x = 0; x < R; x++) {
            const int e = lane * R + x;
            if (e < num) {
                const int32_t id = ids[pos[x]];
                const double s = row[id];
                double rating = a.r_val[cb + pos[x]];
                if (a.knn_type == RS_KNN_CENTERED) rating -= a.means[id];                       // core/knn.go:121
                else if (a.knn_type == RS_KNN_ZSCORE) rating = (rating - a.means[id]) / a.stddevs[id];
                else if (a.knn_type == RS_KNN_BASELINE) rating -= a.bias[id];
                w_s[e] = s;
                w_a[e] = rating;
                if (a.nb_ids && e < a.nb_cap) { a.nb_ids[e] = id; a.nb_sims[e] = s; }
            }
        }
        __syncwarp();
        double wsum = 0.0, wrat = 0.0;
#pragma unroll 4
        for (int e = 0; e < num; e++) {
            const double sq = w_s[e], aq = w_a[e];
            wsum += sq;                                        // core/knn.go:117
            wrat += sq * aq;                                   // core/knn.go:127
        }
        __syncwarp();
        if (lane == 0) {
            double pred = wrat / wsum;                         // core/knn.go:131
            if (a.knn_type == RS_KNN_CENTERED) pred += a.means[l];
            else if (a.knn_type == RS_KNN_BASELINE) pred += a.bias[l];
            else if (a.knn_type == RS_KNN_ZSCORE) { pred *= a.stddevs[l]; pred += a.means[l]; }
            a.out[p] = pred;
            if (a.nb_count) *a.nb_count = num < a.nb_cap ? num : a.nb_cap;
        }
    }
}

// Generic shared-memory kernel: finishes the (rare) predictions on the overflow list.
__global__ void __launch_bounds__(PRED_WARPS * 32) predict_kernel(PredArgs a, const int32_t *__restrict__ list,
                                                                  const int32_t *__restrict__ list_count) {
    __shared__ uint64_t s_keys[PRED_WARPS][PRED_CAP];
    __shared__ uint32_t s_pos[PRED_WARPS][PRED_CAP];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t *keys = s_keys[warp];
    uint32_t *pos = s_pos[warp];
    const int64_t n_warps = (int64_t)gridDim.x * PRED_WARPS;
    const double nan_v = __longlong_as_double(0x7ff8000000000001ll);

    const int64_t n_work = list ? (int64_t)*list_count : a.n;
    for (int64_t w = (int64_t)blockIdx.x * PRED_WARPS + warp; w < n_work; w += n_warps) {
        const int64_t p = list ? (int64_t)list[w] : w;
        const int32_t l = a.left[p], r = a.right[p];
        if (a.nb_count && lane == 0) *a.nb_count = 0;
        if (l < 0 || r < 0 || r >= a.n_right) {            // core/knn.go:89-91 (newID)
            if (lane == 0) a.out[p] = a.global_mean;
            continue;
        }
        if (l < a.row_begin || l >= a.row_end) {           // not in this shard: flagged as NaN
            if (lane == 0) a.out[p] = nan_v;
            continue;
        }
        const double *row = a.sims + (l - a.row_begin) * a.ld_s;
        const int64_t cb = a.r_ptr[r], ce = a.r_ptr[r + 1];
        const int keep = a.k < PRED_CAP / 2 ? a.k : PRED_CAP / 2;

        // ---- gather + filter; keep the best `keep` so far in keys[0..have) ----
        int have = 0;        // records currently buffered (warp-uniform)
        int64_t valid = 0;   // non-NaN candidates seen (core/knn.go:95-99)
        uint64_t thr_key = 0;  // once a sort has happened: records not before (thr) are dropped
        uint32_t thr_pos = 0xffffffffu;
        bool have_thr = false;
        for (int64_t base = cb; base < ce; base += 32) {
            const int64_t x = base + lane;
            bool ok = false;
            uint64_t key = 0;
            if (x < ce) {
                const double s = row[a.r_col[x]];
                ok = (s == s);
                key = rs_sim_key(s);
            }
            const uint32_t okmask = __ballot_sync(0xffffffffu, ok);
            valid += __popc(okmask);
            const uint32_t pp = (uint32_t)(x - cb);
            bool take = ok && (!have_thr || rec_before(key, pp, thr_key, thr_pos));
            const uint32_t tmask = __ballot_sync(0xffffffffu, take);
            if (take) {
                int slot = have + __popc(tmask & ((1u << lane) - 1u));
                keys[slot] = key;
                pos[slot] = pp;
            }
            have += __popc(tmask);
            if (have > PRED_CAP - 32) {
                // buffer nearly full: sort, keep the best `keep`, remember the threshold
                int n_pad = PRED_CAP;
                __syncwarp();
                for (int t = have + lane; t < n_pad; t += 32) { keys[t] = 0; pos[t] = 0xffffffffu; }
                warp_bitonic(keys, pos, n_pad, lane);
                have = keep;
                thr_key = keys[keep - 1];
                thr_pos = pos[keep - 1];
                have_thr = true;
                __syncwarp();
            }
        }
        if (valid <= (int64_t)a.min_k) {                   // core/knn.go:102-104 (note <=)
            if (lane == 0) a.out[p] = a.global_mean;
            continue;
        }
        int n_pad = 32;
        while (n_pad < have) n_pad <<= 1;
        __syncwarp();
        for (int t = have + lane; t < n_pad; t += 32) { keys[t] = 0; pos[t] = 0xffffffffu; }
        warp_bitonic(keys, pos, n_pad, lane);

        int num = a.k;                                     // core/knn.go:111-114
        if ((int64_t)num > valid) num = (int)valid;
        if (num > have) num = have;                        // only when k > PRED_CAP/2 (rejected by the host)

        // ---- weighted mean over the first `num`, sequential in sorted order ----
        double wsum = 0.0, wrat = 0.0;
        for (int b0 = 0; b0 < num; b0 += 32) {
            const int t = b0 + lane;
            double s = 0.0, adj = 0.0;
            int32_t id = -1;
            if (t < num) {
                const int64_t x = cb + pos[t];
                id = a.r_col[x];
                s = row[id];
                double rating = a.r_val[x];
                if (a.knn_type == RS_KNN_CENTERED) rating -= a.means[id];                       // core/knn.go:121
                else if (a.knn_type == RS_KNN_ZSCORE) rating = (rating - a.means[id]) / a.stddevs[id];
                else if (a.knn_type == RS_KNN_BASELINE) rating -= a.bias[id];
                adj = rating;
                if (a.nb_ids && t < a.nb_cap) { a.nb_ids[t] = id; a.nb_sims[t] = s; }
            }
            const int lim = (num - b0) < 32 ? (num - b0) : 32;
            for (int q = 0; q < lim; q++) {
                const double sq = __shfl_sync(0xffffffffu, s, q);
                const double aq = __shfl_sync(0xffffffffu, adj, q);
                wsum += sq;                                // core/knn.go:117
                wrat += sq * aq;                           // core/knn.go:127
            }
        }
        if (lane == 0) {
            double pred = wrat / wsum;                     // core/knn.go:131
            if (a.knn_type == RS_KNN_CENTERED) pred += a.means[l];
            else if (a.knn_type == RS_KNN_BASELINE) pred += a.bias[l];
            else if (a.knn_type == RS_KNN_ZSCORE) { pred *= a.stddevs[l]; pred += a.means[l]; }
            a.out[p] = pred;
            if (a.nb_count) *a.nb_count = num < a.nb_cap ? num : a.nb_cap;
        }
    }
}

// ---------------- per-row top-k from the resident matrix ----------------
constexpr int TOPK_THREADS = 256;
constexpr int TOPK_CAP = 2048;

__device__ void block_bitonic(uint64_t *keys, uint32_t *pos, int n_pad) {
    for (int size = 2; size <= n_pad; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < (n_pad >> 1); t += blockDim.x) {
                int lo = 2 * t - (t & (stride - 1));
                int hi = lo + stride;
                bool up = ((lo & size) == 0);
                uint64_t ka = keys[lo], kb = keys[hi];
                uint32_t pa = pos[lo], pb = pos[hi];
                bool swap = up ? rec_before(kb, pb, ka, pa) : rec_before(ka, pa, kb, pb);
                if (swap) { keys[lo] = kb; keys[hi] = ka; pos[lo] = pb; pos[hi] = pa; }
            }
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(TOPK_THREADS) topk_rows_kernel(const double *__restrict__ sims, int64_t ld_s,
                                                                 int32_t n, int32_t k, int32_t *__restrict__ idx,
                                                                 double *__restrict__ sim) {
    __shared__ uint64_t keys[TOPK_CAP];
    __shared__ uint32_t pos[TOPK_CAP];
    __shared__ int s_have;
    __shared__ uint64_t s_thr_key;
    __shared__ uint32_t s_thr_pos;
    const int64_t r = blockIdx.x;
    const double *row = sims + r * ld_s;
    if (threadIdx.x == 0) { s_have = 0; s_thr_key = 0; s_thr_pos = 0xffffffffu; }
    __syncthreads();
    bool have_thr = false;
    for (int32_t base = 0; base < n; base += TOPK_THREADS) {
        const int32_t j = base + threadIdx.x;
        bool take = false;
        uint64_t key = 0;
        if (j < n) {
            const double s = row[j];
            if (s == s) {
                key = rs_sim_key(s);
                take = !have_thr || rec_before(key, (uint32_t)j, s_thr_key, s_thr_pos);
            }
        }
        if (take) {
            int slot = atomicAdd(&s_have, 1);
            keys[slot] = key;
            pos[slot] = (uint32_t)j;
        }
        __syncthreads();
        if (s_have > TOPK_CAP - TOPK_THREADS) {
            const int have = s_have;
            for (int t = have + threadIdx.x; t < TOPK_CAP; t += blockDim.x) { keys[t] = 0; pos[t] = 0xffffffffu; }
            block_bitonic(keys, pos, TOPK_CAP);
            if (threadIdx.x == 0) {
                s_have = k;
                s_thr_key = keys[k - 1];
                s_thr_pos = pos[k - 1];
            }
            have_thr = true;
            __syncthreads();
        }
    }
    const int have = s_have;
    int n_pad = 32;
    while (n_pad < have) n_pad <<= 1;
    for (int t = have + threadIdx.x; t < n_pad; t += blockDim.x) { keys[t] = 0; pos[t] = 0xffffffffu; }
    block_bitonic(keys, pos, n_pad);
    for (int t = threadIdx.x; t < k; t += blockDim.x) {
        const int64_t o = r * k + t;
        if (t < have) { idx[o] = (int32_t)pos[t]; sim[o] = rs_key_sim(keys[t]); }
        else { idx[o] = -1; sim[o] = __longlong_as_double(0x7ff8000000000001ll); }
    }
}

}  // namespace

template <int R>
static int32_t launch_select(const PredArgs &a, int32_t *ovf, unsigned blocks, size_t smem, int scap, cudaStream_t st) {
    auto kern = predict_select_kernel<R>;
    RS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<blocks, SEL_WARPS * 32, smem, st>>>(a, ovf + 1, ovf, scap);
    return RS_OK;
}

int32_t rs_predict_launch(rs_knn *h, const int32_t *d_left, const int32_t *d_right, int64_t n, double *d_out,
                          int32_t *d_nb_ids, double *d_nb_sims, int32_t *d_nb_count, int32_t nb_cap) {
    if (n <= 0) return RS_OK;
    if (h->p.k > PRED_CAP / 2) {
        rs_set_error("k=%d exceeds the %d neighbours the predict kernel supports", h->p.k, PRED_CAP / 2);
        return RS_ERR_UNSUPPORTED;
    }
    if (n >= (1ll << 31)) {
        rs_set_error("at most 2^31-1 predictions per call");
        return RS_ERR_UNSUPPORTED;
    }
    PredArgs a{};
    a.left = d_left; a.right = d_right; a.n = n; a.out = d_out;
    a.r_ptr = h->r_ptr; a.r_col = h->r_col; a.r_val = h->r_val;
    a.sims = h->sims; a.ld_s = h->ld_s; a.row_begin = h->row_begin; a.row_end = h->row_end;
    a.means = h->means; a.stddevs = h->stddevs; a.bias = h->left_bias;
    a.global_mean = h->global_mean; a.n_right = h->n_right;
    a.k = h->p.k; a.min_k = h->p.min_k; a.knn_type = h->p.knn_type;
    a.nb_ids = d_nb_ids; a.nb_sims = d_nb_sims; a.nb_count = d_nb_count; a.nb_cap = nb_cap;
    // overflow list: [0] = count, [1..n] = prediction indices
    if ((size_t)(n + 1) * 4 > h->ovf_bytes) {
        RS_CUDA(cudaStreamSynchronize(h->stream));
        if (h->ovf) rs_cached_free(h->device, h->ovf, h->ovf_bytes);
        h->ovf = nullptr;
        h->ovf_bytes = 0;
        const size_t want = (size_t)(n + 1) * 4 + (size_t)n;  // 25 % headroom
        size_t got = 0;
        RS_TRY(rs_cached_malloc(h->device, &h->ovf, want, &got));
        h->ovf_bytes = got;
    }
    int32_t *ovf = reinterpret_cast<int32_t *>(h->ovf);
    RS_CUDA(cudaMemsetAsync(ovf, 0, 4, h->stream));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
    int64_t blocks = (n + SEL_WARPS - 1) / SEL_WARPS;
    const int64_t cap = (int64_t)sms * 8;   // resident CTAs; warps stride over the predictions
    if (blocks > cap) blocks = cap;
    // staging capacity per warp (similarities parked in shared memory between the passes)
    int scap = 512;    // measured best on the ML-1M shape (0: 1.62 ms, 512: 1.44, 1024: 1.70, 2048: 2.54 — occupancy)
    if (const char *e = getenv("RS_KNN_PRED_SCAP")) scap = atoi(e);
    if (scap < 0) scap = 0;
    if (scap > 2048) scap = 2048;
    scap = scap / 32 * 32;
    const size_t smem = (size_t)SEL_WARPS * scap * sizeof(double);
    const int64_t per_sm = smem ? (int64_t)(200 * 1024) / (int64_t)(smem + 16 * 1024) : 8;
    if (blocks > (int64_t)sms * (per_sm > 0 ? per_sm : 1)) blocks = (int64_t)sms * (per_sm > 0 ? per_sm : 1);
    // register capacity of the selection kernel: room for k plus a boundary bucket
    if (h->p.k <= 44) RS_TRY(launch_select<2>(a, ovf, (unsigned)blocks, smem, scap, h->stream));
    else if (h->p.k <= 104) RS_TRY(launch_select<4>(a, ovf, (unsigned)blocks, smem, scap, h->stream));
    else RS_TRY(launch_select<8>(a, ovf, (unsigned)blocks, smem, scap, h->stream));
    // the generic kernel drains the overflow list (normally empty: a handful of warps exit at once)
    predict_kernel<<<(unsigned)sms, PRED_WARPS * 32, 0, h->stream>>>(a, ovf + 1, ovf);
    h->prof.predict_launches++;
    h->prof.total_launches += 2;
    RS_CUDA(cudaGetLastError());
    return RS_OK;
}

int32_t rs_topk_launch(rs_knn *h, int32_t k, int32_t *d_idx, double *d_sim) {
    const int64_t rows = h->row_end - h->row_begin;
    if (rows <= 0) return RS_OK;
    if (k < 1 || k > TOPK_CAP / 4) {
        rs_set_error("top-k supports 1 <= k <= %d (got %d)", TOPK_CAP / 4, k);
        return RS_ERR_UNSUPPORTED;
    }
    topk_rows_kernel<<<(unsigned)rows, TOPK_THREADS, 0, h->stream>>>(h->sims, h->ld_s, h->n_left, k, d_idx, d_sim);
    h->prof.total_launches++;
    RS_CUDA(cudaGetLastError());
    return RS_OK;
}
