// sim_stream.cu — exact-order FP64 similarity kernel ("stream" path).
//
// Reproduces core/sim.go (Cosine :10-25, MSD :28-44, Pearson :47-81) bit for bit.  For a left
// row i the reference walks i's entries in ascending right id c and, for every other left row
// j that also rated c, adds one term to each of three running sums — in that order.
//
// Mapping ("column walk"): one WARP owns a work item (row i, chunk of JC = 128 or 256 consecutive columns j)
// and keeps that chunk's accumulators in shared memory.  It walks row i's entries in ascending
// c; for each c the raters of c that fall in the chunk are a CONTIGUOUS slice of c's id-sorted
// list in the right CSR (chunk pointers `cp`, precomputed), so the lanes read (j, b-side term)
// coalesced and update acc[j] — every j at most once per c, and c strictly in order, so each
// accumulator receives exactly the reference's terms in the reference's order with the same
// IEEE operations (the library is built with --fmad=false).  With the full matrix only j > i
// is visited (the slice starts right after i's own position in c's list, `l2r`); the mirror
// pass fills j < i.  Work = the co-rated triples; no N x nnz term anywhere.
//
// Bound: shared-memory RMW bandwidth (6 accesses per triple, random banks) and L2 reads of
// 12 B per triple; see DESIGN.md and profiles/.
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace {

constexpr int SW = RS_STREAM_WARPS;        // warps per CTA, each an independent work item

struct StreamArgs {
    const int64_t *l_ptr;
    const int32_t *l_col;
    const double *l_val;
    const int64_t *l2r;       // left-CSR entry (i,c) -> index of (c,i) in the right CSR
    const int64_t *r_ptr;
    const int32_t *r_col;
    const double *r_dev;      // b-side term per right-CSR entry (rating, or rating - row mean, ...)
    const int32_t *row_order; // left rows of the shard, longest first (load balance)
    const int32_t *cp;        // [n_right][Q+1] offsets (relative to r_ptr[c]) of the chunk boundaries
    int32_t n_chunks;         // Q
    const double *pmeans;
    const double *left_bias, *right_bias;
    double global_bias, shrinkage;
    double *sims;
    int64_t ld_s;
    int32_t n_left;
    int64_t row_begin, row_end;
    int64_t n_rows;           // rows of row_order to walk
    int cyc_R;                // >= 2: cyclic row shards, the output row is rs_cyc_local(i)
    int symmetric;            // 0: full rows; 1: only columns j > i are computed; 2: only j < i (the mirror pass fills the rest)
    unsigned long long *counter;
};

constexpr int G = 8;   // columns whose first 32 raters are loaded together (must divide 32)

template <int SIM, bool SHRINK, int JC>
__device__ __forceinline__ void stream_update(double *acc, int j, double ra, double raa, double rb) {
    if (SIM == RS_SIM_MSD) {
        const double d = ra - rb;
        acc[j] += d * d;                 // sum += (ir-jr)^2   core/sim.go:37
        acc[JC + j] += 1.0;              // count++            core/sim.go:38
    } else {
        acc[j] += raa;                   // m += ..            core/sim.go:19 / :75
        acc[JC + j] += rb * rb;          // n += rb*rb         core/sim.go:20 / :76
        acc[2 * JC + j] += ra * rb;      // l += ra*rb         core/sim.go:21 / :77
        if (SHRINK) acc[3 * JC + j] += 1.0;
    }
}

// SYM = 1: the full matrix is being computed, only columns j > i are visited (the run starts right
// after i's own position in c's list) and the mirror pass fills j < i.  SYM = 2: only columns j < i
// (the run ends at i's own position).  Which triangle is cheaper depends on how the row lengths
// correlate with the ids: a row pays one lookup per (entry, chunk) it visits, so long rows should
// visit few chunks — with first-appearance ids the popular rows have the LOW ids and the lower
// triangle costs 3.5x fewer lookups on the MovieLens-20M shape (rs_prep_rt picks per Fit).
// SYM = 0 (row shard): the whole run is visited and the diagonal pair (i,i) is skipped by index.
// Indices into the right CSR are kept in 32 bits inside the loop (nnz < 2^32 is checked at launch)
// and the per-entry a-side terms are staged in shared memory, which halves the instructions per
// column against the first version of this kernel (profiles/r01_stream_notes.md).
template <int SIM, bool SHRINK, int SYM, int JC>
__global__ void __launch_bounds__(SW * 32) sim_stream_kernel(StreamArgs a) {
    constexpr int NACC = SHRINK ? 4 : 3;
    extern __shared__ double s_acc_all[];                    // [SW][NACC][JC] accumulators, then [SW][32] a-side terms
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *acc = s_acc_all + (size_t)warp * NACC * JC;
    double *s_ra = s_acc_all + (size_t)SW * NACC * JC + warp * 32;
    const int64_t Q = a.n_chunks;
    const int64_t n_items = a.n_rows * Q;
    const double nan_v = __longlong_as_double(0x7ff8000000000001ll);
    const int32_t *__restrict__ r_col = a.r_col;
    const double *__restrict__ r_dev = a.r_dev;

    for (;;) {
        unsigned long long item = 0;
        if (lane == 0) item = atomicAdd(a.counter, 1ull);
        item = __shfl_sync(0xffffffffu, item, 0);
        if ((int64_t)item >= n_items) break;
        // items are ordered longest row first, so the critical path starts early
        const int64_t q = (int64_t)item % Q;
        const int32_t i = a.row_order[(int64_t)item / Q];
        const int j0 = (int)(q * JC);
        if (SYM == 1 && j0 + JC <= i) continue;  // every column of the chunk is < i
        if (SYM == 2 && j0 > i) continue;        // every column of the chunk is > i

        for (int x = lane; x < NACC * JC; x += 32) acc[x] = 0.0;

        double ai = 0.0;
        if (SIM == RS_SIM_PEARSON) ai = a.pmeans[i];
        if (SIM == RS_SIM_PEARSON_BASELINE) ai = a.global_bias + a.left_bias[i];

        const int64_t eb = a.l_ptr[i], ee = a.l_ptr[i + 1];
        for (int64_t x0 = eb; x0 < ee; x0 += 32) {
            // each lane prepares one entry (c, x) of row i: its a-side term and its run of c's list
            const int64_t e = x0 + lane;
            double ra = 0.0;
            uint32_t lo = 0, self = 0xffffffffu;
            int n = 0;
            if (e < ee) {
                const int32_t c = a.l_col[e];
                const double v = a.l_val[e];
                if (SIM == RS_SIM_PEARSON) ra = v - ai;                           // core/sim.go:73
                else if (SIM == RS_SIM_PEARSON_BASELINE) { const double bb = ai + a.right_bias[c]; ra = v - bb; }
                else ra = v;
                const int64_t rp = a.r_ptr[c];
                const int32_t *cpc = a.cp + (int64_t)c * (Q + 1) + q;
                int64_t lo64 = rp + cpc[0];
                int64_t hi = rp + cpc[1];
                const int64_t self64 = a.l2r[e];
                if (SYM == 1) { if (self64 + 1 > lo64) lo64 = self64 + 1; }       // only j > i
                else if (SYM == 2) { if (self64 < hi) hi = self64; }              // only j < i
                else self = (uint32_t)self64;
                n = hi > lo64 ? (int)(hi - lo64) : 0;
                lo = (uint32_t)lo64;
            }
            __syncwarp();           // the previous batch has finished with acc and s_ra
            s_ra[lane] = ra;
            __syncwarp();
            const int lim = (ee - x0) < 32 ? (int)(ee - x0) : 32;
            // Columns are processed in order, G at a time: the (j, b-side) pairs of the first 32
            // raters of G consecutive columns are loaded up front (independent L2 requests in
            // flight), then applied column by column.
            for (int u0 = 0; u0 < lim; u0 += G) {
                int jj[G], nn[G];
                double rbv[G];
#pragma unroll
                for (int g = 0; g < G; g++) {
                    const int u = u0 + g;                                           // < 32 always (G divides 32)
                    nn[g] = __shfl_sync(0xffffffffu, n, u);                         // 0 for u >= lim
                    const uint32_t lo_u = __shfl_sync(0xffffffffu, lo, u);
                    uint32_t self_u = 0xffffffffu;
                    if (!SYM) self_u = __shfl_sync(0xffffffffu, self, u);
                    jj[g] = -1;
                    rbv[g] = 0.0;
                    if (lane < nn[g]) {
                        const uint32_t idx = lo_u + (uint32_t)lane;
                        if (SYM || idx != self_u) {                                 // the diagonal pair (i,i)
                            jj[g] = r_col[idx] - j0;
                            rbv[g] = r_dev[idx];                                    // jr | jr - meanB (core/sim.go:74)
                        }
                    }
                }
#pragma unroll
                for (int g = 0; g < G; g++) {
                    if (nn[g] == 0) continue;                                       // warp-uniform
                    const double ra_u = s_ra[u0 + g];
                    const double raa_u = ra_u * ra_u;                               // core/sim.go:19 / :75
                    if (jj[g] >= 0) stream_update<SIM, SHRINK, JC>(acc, jj[g], ra_u, raa_u, rbv[g]);
                    if (nn[g] > 32) {                                               // long run: remaining raters
                        const uint32_t lo_u = __shfl_sync(0xffffffffu, lo, u0 + g);
                        uint32_t self_u = 0xffffffffu;
                        if (!SYM) self_u = __shfl_sync(0xffffffffu, self, u0 + g);
                        for (int t = lane + 32; t < nn[g]; t += 32) {
                            const uint32_t idx = lo_u + (uint32_t)t;
                            if (!SYM && idx == self_u) continue;
                            stream_update<SIM, SHRINK, JC>(acc, r_col[idx] - j0, ra_u, raa_u, r_dev[idx]);
                        }
                    }
                    __syncwarp();   // column c is complete before column c+1 touches the same j
                }
            }
        }
        __syncwarp();

        // epilogue: JC similarities of row i, coalesced
        double *out = a.sims + (a.cyc_R > 1 ? rs_cyc_local(i, a.cyc_R) : (int64_t)(i - a.row_begin)) * a.ld_s + j0;
        for (int j = lane; j < JC; j += 32) {
            const int64_t col = (int64_t)j0 + j;
            if (col >= a.n_left) break;
            if (SYM == 1 && col < i) continue;                                    // mirror pass writes it
            if (SYM == 2 && col > i) break;
            double s;
            if (SIM == RS_SIM_MSD) s = 1.0 / (acc[j] / acc[JC + j] + 1.0);        // core/sim.go:43
            else s = acc[2 * JC + j] / (sqrt(acc[j]) * sqrt(acc[JC + j]));        // core/sim.go:24 / :80
            if (SHRINK) {
                const double cn = acc[3 * JC + j];
                s = (cn - 1.0) / (cn - 1.0 + a.shrinkage) * s;
            }
            if (col == (int64_t)i) s = nan_v;                                     // diagonal stays NaN
            out[j] = s;
        }
        __syncwarp();
    }
}


// Mirror the computed upper block-triangle into the lower one: the three similarities are
// bit-symmetric (sums and products commute), which is why the reference can write
// Sims[j][i] = Sims[i][j] (core/knn.go:205-208).  32x32 tiles through shared memory,
// coalesced on both sides.  `chunk` is the column-chunk width the producer skipped by.
// lower = 1: the producer computed j < i, the destinations are the cells right of the diagonal.
__global__ void symmetrize_kernel(double *__restrict__ s, int64_t ld, int32_t n, int lower) {
    __shared__ double tile[32][33];
    int64_t bi = blockIdx.y, bj = blockIdx.x;                // destination tile (rows bi, cols bj)
    if (bj > bi) return;                                     // one block per unordered tile pair
    if (lower) { const int64_t t = bi; bi = bj; bj = t; }
    const int64_t r0 = bi * 32, c0 = bj * 32;
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {
        int64_t sr = c0 + y, sc = r0 + threadIdx.x;          // source = transposed position
        tile[y][threadIdx.x] = (sr < n && sc < n) ? s[sr * ld + sc] : 0.0;
    }
    __syncthreads();
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {
        int64_t r = r0 + y, c = c0 + threadIdx.x;
        if (r < n && c < n && (lower ? c > r : c < r)) s[r * ld + c] = tile[threadIdx.x][y];
    }
}

// Cyclic row shards: every shard computed one triangle of ITS rows (lower = 1: the cells j < i); the
// other triangle of a row is the transpose of cells that live in the rows of other shards.  One CTA
// moves one 32 x 32 tile: it reads the source tile from the owner's matrix — peer memory over NVLink
// when the owner is another GPU, 256 contiguous bytes per row —, transposes it in shared memory and
// writes it into this shard's rows.  This is the one exchange step of the sharded Fit: half the
// matrix crosses the links once, pulled by the consumers (no staging buffer, no pack / unpack).
struct MirrorArgs {
    double *self;
    const double *peer[RS_MAX_PEERS];
    int64_t ld;
    int32_t n;
    int32_t count, index, lower;
};
__global__ void mirror_kernel(MirrorArgs a) {
    __shared__ double tile[32][33];
    const int64_t bI = blockIdx.y;                           // own block (local index)
    const int64_t gI = bI * a.count + a.index, gJ = blockIdx.x;
    if (a.lower ? gJ < gI : gJ > gI) return;                 // destination tiles: right of (left of) the diagonal
    const double *src = a.peer[gJ % a.count];
    const int64_t sJ = (gJ / a.count) * RS_CYC_B;            // first local row of block gJ at its owner
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {
        const int64_t sr = gJ * 32 + y, sc = gI * 32 + threadIdx.x;   // global (row, col) of the source cell
        tile[y][threadIdx.x] = (sr < a.n && sc < a.n) ? src[(sJ + y) * a.ld + sc] : 0.0;
    }
    __syncthreads();
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {
        const int64_t r = gI * 32 + y, c = gJ * 32 + threadIdx.x;
        if (r < a.n && c < a.n && (a.lower ? c > r : c < r)) a.self[(bI * 32 + y) * a.ld + c] = tile[threadIdx.x][y];
    }
}

}  // namespace

template <int SIM, bool SHRINK, int SYM, int JC>
static int32_t launch_stream_jc(rs_knn *h, const StreamArgs &s, int grid) {
    const int smem = SW * ((SHRINK ? 4 : 3) * JC + 32) * (int)sizeof(double);
    auto kern = sim_stream_kernel<SIM, SHRINK, SYM, JC>;
    RS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<grid, SW * 32, smem, h->stream>>>(s);
    return RS_OK;
}
template <int SIM, bool SHRINK, int SYM>
static int32_t launch_stream_sym(rs_knn *h, const StreamArgs &s, int grid) {
    if (h->stream_jc == 128) return launch_stream_jc<SIM, SHRINK, SYM, 128>(h, s, grid);
    return launch_stream_jc<SIM, SHRINK, SYM, 256>(h, s, grid);
}
template <int SIM, bool SHRINK>
static int32_t launch_stream(rs_knn *h, const StreamArgs &s, int grid) {
    if (h->nnz >= (int64_t)0xffffffffll) {
        rs_set_error("stream path indexes the ratings with 32 bits (nnz=%lld)", (long long)h->nnz);
        return RS_ERR_UNSUPPORTED;
    }
    if (s.symmetric == 2) return launch_stream_sym<SIM, SHRINK, 2>(h, s, grid);
    return s.symmetric ? launch_stream_sym<SIM, SHRINK, 1>(h, s, grid) : launch_stream_sym<SIM, SHRINK, 0>(h, s, grid);
}

int32_t rs_sim_stream_launch(rs_knn *h) {
    StreamArgs a{};
    a.l_ptr = h->l_ptr; a.l_col = h->l_col; a.l_val = h->l_val; a.l2r = h->l2r;
    a.r_ptr = h->r_ptr; a.r_col = h->r_col; a.r_dev = h->r_dev; a.cp = h->cp; a.n_chunks = h->n_chunks;
    const bool cyc = h->cyc_R > 1;
    // rs_prep_rt leaves the rows to walk (longest first) in row_order; in top-k-only
    // mode (n_work_rows < 0) it is the natural order and the current slab is a slice of it
    a.row_order = h->n_work_rows >= 0 ? h->row_order : h->row_order + h->row_begin;
    a.n_rows = h->n_work_rows >= 0 ? h->n_work_rows : h->row_end - h->row_begin;
    a.cyc_R = h->cyc_R;
    a.pmeans = h->pmeans; a.left_bias = h->left_bias; a.right_bias = h->right_bias;
    a.global_bias = h->global_bias; a.shrinkage = h->p.shrinkage;
    a.sims = h->sims; a.ld_s = h->ld_s; a.n_left = h->n_left;
    a.row_begin = h->row_begin; a.row_end = h->row_end;
    a.symmetric = h->force_sym ? 1 : ((h->row_begin == 0 && h->row_end == h->n_left) || cyc) ? (h->stream_lower ? 2 : 1) : 0;
    a.counter = reinterpret_cast<unsigned long long *>(h->d_flags + 2);
    RS_CUDA(cudaMemsetAsync(a.counter, 0, 8, h->stream));
    if (a.n_rows <= 0) return RS_OK;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
    const int64_t items = a.n_rows * (int64_t)h->n_chunks;
    int64_t grid = (int64_t)sms * 4;          // resident CTAs; warps pull work items from the counter
    if (grid > (items + SW - 1) / SW) grid = (items + SW - 1) / SW;
    switch (h->p.sim) {
    case RS_SIM_COSINE: RS_TRY((launch_stream<RS_SIM_COSINE, false>(h, a, (int)grid))); break;
    case RS_SIM_MSD: RS_TRY((launch_stream<RS_SIM_MSD, false>(h, a, (int)grid))); break;
    case RS_SIM_PEARSON: RS_TRY((launch_stream<RS_SIM_PEARSON, false>(h, a, (int)grid))); break;
    case RS_SIM_PEARSON_BASELINE:
        if (h->p.shrinkage > 0.0) RS_TRY((launch_stream<RS_SIM_PEARSON_BASELINE, true>(h, a, (int)grid)));
        else RS_TRY((launch_stream<RS_SIM_PEARSON_BASELINE, false>(h, a, (int)grid)));
        break;
    default: rs_set_error("unknown similarity %d", h->p.sim); return RS_ERR_INVALID;
    }
    h->prof.sim_launches++;
    h->prof.total_launches++;
    RS_CUDA(cudaGetLastError());
    return RS_OK;
}

int32_t rs_mirror_launch(rs_knn *h) {
    MirrorArgs a{};
    a.self = h->sims;
    for (int q = 0; q < h->cyc_R; q++) a.peer[q] = h->peer_sims[q];
    a.peer[h->cyc_r] = h->sims;
    a.ld = h->ld_s; a.n = h->n_left; a.count = h->cyc_R; a.index = h->cyc_r; a.lower = h->stream_lower ? 1 : 0;
    const int64_t nblk = ((int64_t)h->n_left + RS_CYC_B - 1) / RS_CYC_B;
    const int64_t own = (nblk - h->cyc_r + h->cyc_R - 1) / h->cyc_R;
    if (own <= 0) return RS_OK;
    dim3 grid((unsigned)nblk, (unsigned)own), block(32, 8);
    mirror_kernel<<<grid, block, 0, h->stream>>>(a);
    h->prof.total_launches++;
    RS_CUDA(cudaGetLastError());
    return RS_OK;
}

int32_t rs_symmetrize_launch(rs_knn *h) {
    if (h->cyc_R > 1) return RS_OK;                          // cyclic shards: rs_knn_mirror after the peers are attached
    if (!(h->row_begin == 0 && h->row_end == h->n_left)) return RS_OK;
    const unsigned t = (unsigned)((h->n_left + 31) / 32);
    dim3 grid(t, t), block(32, 8);
    symmetrize_kernel<<<grid, block, 0, h->stream>>>(h->sims, h->ld_s, h->n_left, h->stream_lower ? 1 : 0);
    h->prof.total_launches++;
    RS_CUDA(cudaGetLastError());
    return RS_OK;
}
