// sim_stream.cu — exact-order FP64 similarity kernel ("stream" path).
//
// Reproduces core/sim.go (Cosine :10-25, MSD :28-44, Pearson :47-81) bit for bit.  For a left
// row i the reference walks i's entries in ascending right id c and, for every other left row
// j that also rated c, adds one term to each of three running sums — in that order.  Here
//   * a thread OWNS 8 consecutive j and their accumulators (shared memory, layout
//     [sum][bit][thread] so a warp's 64-bit accesses never bank-conflict);
//   * the CTA walks row i's entries in ascending c; per entry every thread reads one 8-byte
//     {mask, prefix} word of the bit matrix  MP[c][j/32]  (bit = "row j rated c", prefix =
//     rank of the word's first entry in c's id-sorted rating list) and visits ONLY its set
//     bits; the b-side term of each visited entry (rating - row mean, core/sim.go:74, and its
//     square) was precomputed once per rating in the same IEEE operations;
//   * each accumulator therefore receives exactly the reference's terms in the reference's
//     order (the library is built with --fmad=false).
// Work is proportional to the co-rated triples plus one 8-byte word per (entry, 32 columns),
// not to N x nnz bytes.  Bound: issue slots / FP64 pipe / L2; see DESIGN.md and profiles/.
#include "common.cuh"

namespace {

constexpr int ST = RS_STREAM_THREADS;      // 256 threads
constexpr int JPT = RS_STREAM_JPT;         // 8 columns per thread
constexpr int TS = 256;                    // entries of row i staged per pass

struct StreamArgs {
    const int64_t *l_ptr;
    const int32_t *l_col;
    const double *l_val;
    const int64_t *r_ptr;
    const uint2 *mp;          // [n_right][words] {mask, prefix}
    int64_t words;            // words per right row (multiple of ST*JPT/32)
    const double *r_dev;      // b-side term per right-CSR entry
    const double *r_dev2;     // its square
    const double *pmeans;
    const double *left_bias, *right_bias;
    double global_bias, shrinkage;
    double *sims;
    int64_t ld_s;
    int32_t n_left;
    int64_t row_begin;
    int symmetric;            // 1: only columns j > i are computed (the mirror pass fills j < i)
};

template <int SIM, bool SHRINK>
__global__ void __launch_bounds__(ST) sim_stream_kernel(StreamArgs a) {
    constexpr int NACC = SHRINK ? 4 : 3;
    extern __shared__ double s_acc[];      // [NACC][JPT][ST]
    __shared__ int32_t s_c[TS];
    __shared__ int64_t s_off[TS];
    __shared__ double s_a[TS];
    __shared__ double s_aa[TS];

    const int tid = threadIdx.x;
    const int32_t i = (int32_t)(a.row_begin + blockIdx.y);
    const int64_t jb = ((int64_t)blockIdx.x * ST + tid) * JPT;          // first column of this thread
    if (a.symmetric && ((int64_t)(blockIdx.x + 1) * ST * JPT <= (int64_t)i)) return;  // every column of this CTA is < i

    // bits of this thread's byte that take part
    uint32_t keep = 0;
#pragma unroll
    for (int b = 0; b < JPT; b++) {
        const int64_t j = jb + b;
        if (j < a.n_left && j != i && (!a.symmetric || j > i)) keep |= 1u << b;
    }
    const int64_t word = jb >> 5;
    const int shift = (int)(jb & 31);
    const uint32_t below = (1u << shift) - 1u;                           // bits of the word before my byte

#pragma unroll
    for (int k = 0; k < NACC; k++)
#pragma unroll
        for (int b = 0; b < JPT; b++) s_acc[(k * JPT + b) * ST + tid] = 0.0;

    double ai = 0.0;
    if (SIM == RS_SIM_PEARSON) ai = a.pmeans[i];
    if (SIM == RS_SIM_PEARSON_BASELINE) ai = a.global_bias + a.left_bias[i];

    const int64_t eb = a.l_ptr[i], ee = a.l_ptr[i + 1];
    for (int64_t base = eb; base < ee; base += TS) {
        const int cnt = (int)((ee - base) < TS ? (ee - base) : TS);
        __syncthreads();
        for (int x = tid; x < cnt; x += ST) {
            const int32_t c = a.l_col[base + x];
            const double v = a.l_val[base + x];
            double ra;
            if (SIM == RS_SIM_PEARSON) ra = v - ai;                               // core/sim.go:73
            else if (SIM == RS_SIM_PEARSON_BASELINE) { const double bb = ai + a.right_bias[c]; ra = v - bb; }
            else ra = v;
            s_c[x] = c;
            s_off[x] = a.r_ptr[c];
            s_a[x] = ra;
            s_aa[x] = ra * ra;                                                    // core/sim.go:19 / :75
        }
        __syncthreads();
        if (keep == 0) continue;

        for (int x0 = 0; x0 < cnt; x0 += 4) {
            uint2 w[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int x = x0 + u;
                w[u] = (x < cnt) ? __ldg(a.mp + (int64_t)s_c[x] * a.words + word) : make_uint2(0u, 0u);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                uint32_t m = (w[u].x >> shift) & 0xffu & keep;
                if (m == 0) continue;
                const int x = x0 + u;
                const uint32_t mine_all = (w[u].x >> shift) & 0xffu;
                const int64_t first = s_off[x] + w[u].y + __popc(w[u].x & below);  // entry of my byte's bit 0.. in c's list
                const double ra = s_a[x], raa = s_aa[x];
                while (m) {
                    const int b = __ffs(m) - 1;
                    m &= m - 1;
                    const int64_t e = first + __popc(mine_all & ((1u << b) - 1u));
                    double *acc = s_acc + b * ST + tid;
                    if (SIM == RS_SIM_MSD) {
                        const double d = ra - a.r_dev[e];
                        acc[0] += d * d;                     // sum += (ir-jr)^2   core/sim.go:37
                        acc[JPT * ST] += 1.0;                // count++            core/sim.go:38
                    } else {
                        const double rb = a.r_dev[e];        // jr (cosine) / jr - meanB (core/sim.go:74)
                        acc[0] += raa;                       // m += ..            core/sim.go:19 / :75
                        acc[JPT * ST] += a.r_dev2[e];        // n += rb*rb         core/sim.go:20 / :76
                        acc[2 * JPT * ST] += ra * rb;        // l += ra*rb         core/sim.go:21 / :77
                        if (SHRINK) acc[3 * JPT * ST] += 1.0;
                    }
                }
            }
        }
    }

    // epilogue: row i of the shard, 64 contiguous bytes per thread
    double *out = a.sims + (int64_t)blockIdx.y * a.ld_s + jb;
#pragma unroll
    for (int b = 0; b < JPT; b++) {
        const int64_t j = jb + b;
        if (j >= a.n_left) continue;
        if (a.symmetric && j < i) continue;                                       // mirror pass writes it
        const double *acc = s_acc + b * ST + tid;
        double s;
        if (SIM == RS_SIM_MSD) s = 1.0 / (acc[0] / acc[JPT * ST] + 1.0);          // core/sim.go:43
        else s = acc[2 * JPT * ST] / (sqrt(acc[0]) * sqrt(acc[JPT * ST]));        // core/sim.go:24 / :80
        if (SHRINK) {
            const double cn = acc[3 * JPT * ST];
            s = (cn - 1.0) / (cn - 1.0 + a.shrinkage) * s;
        }
        if (j == (int64_t)i) s = __longlong_as_double(0x7ff8000000000001ll);      // diagonal stays NaN
        out[b] = s;
    }
}

// Mirror the computed upper block-triangle into the lower one: the three similarities are
// bit-symmetric (sums and products commute), which is why the reference can write
// Sims[j][i] = Sims[i][j] (core/knn.go:205-208).  32x32 tiles through shared memory,
// coalesced on both sides.  `chunk` is the column-chunk width the producer skipped by.
__global__ void symmetrize_kernel(double *__restrict__ s, int64_t ld, int32_t n, int chunk) {
    __shared__ double tile[32][33];
    const int64_t bi = blockIdx.y, bj = blockIdx.x;          // destination tile (rows bi, cols bj)
    const int64_t r0 = bi * 32, c0 = bj * 32;
    // destination (r, c) was skipped by the producer iff chunk(c) < chunk(r)
    if ((c0 / chunk) >= ((r0 + 31) / chunk)) return;  // every destination in this tile was computed
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {
        int64_t sr = c0 + y, sc = r0 + threadIdx.x;          // source = transposed position
        tile[y][threadIdx.x] = (sr < n && sc < n) ? s[sr * ld + sc] : 0.0;
    }
    __syncthreads();
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {
        int64_t r = r0 + y, c = c0 + threadIdx.x;
        if (r < n && c < n && (c / chunk) < (r / chunk)) s[r * ld + c] = tile[threadIdx.x][y];
    }
}

}  // namespace

template <int SIM, bool SHRINK>
static int32_t launch_stream(rs_knn *h, const StreamArgs &s, dim3 g) {
    const int smem = (SHRINK ? 4 : 3) * JPT * ST * (int)sizeof(double);
    auto kern = sim_stream_kernel<SIM, SHRINK>;
    RS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<g, ST, smem, h->stream>>>(s);
    return RS_OK;
}

int32_t rs_sim_stream_launch(rs_knn *h) {
    StreamArgs a{};
    a.l_ptr = h->l_ptr; a.l_col = h->l_col; a.l_val = h->l_val; a.r_ptr = h->r_ptr;
    a.mp = h->mp; a.words = h->mp_words; a.r_dev = h->r_dev; a.r_dev2 = h->r_dev2;
    a.pmeans = h->pmeans; a.left_bias = h->left_bias; a.right_bias = h->right_bias;
    a.global_bias = h->global_bias; a.shrinkage = h->p.shrinkage;
    a.sims = h->sims; a.ld_s = h->ld_s; a.n_left = h->n_left; a.row_begin = h->row_begin;
    const int64_t rows = h->row_end - h->row_begin;
    a.symmetric = (h->row_begin == 0 && h->row_end == h->n_left) ? 1 : 0;
    if (rows <= 0) return RS_OK;
    const unsigned gx = (unsigned)(h->mp_words * 32 / (ST * JPT));
    for (int64_t r0 = 0; r0 < rows; r0 += 65535) {
        StreamArgs s = a;
        const int64_t nr = rows - r0 < 65535 ? rows - r0 : 65535;
        s.row_begin = h->row_begin + r0;
        s.sims = h->sims + r0 * h->ld_s;
        dim3 g(gx, (unsigned)nr);
        switch (h->p.sim) {
        case RS_SIM_COSINE: RS_TRY((launch_stream<RS_SIM_COSINE, false>(h, s, g))); break;
        case RS_SIM_MSD: RS_TRY((launch_stream<RS_SIM_MSD, false>(h, s, g))); break;
        case RS_SIM_PEARSON: RS_TRY((launch_stream<RS_SIM_PEARSON, false>(h, s, g))); break;
        case RS_SIM_PEARSON_BASELINE:
            if (h->p.shrinkage > 0.0) RS_TRY((launch_stream<RS_SIM_PEARSON_BASELINE, true>(h, s, g)));
            else RS_TRY((launch_stream<RS_SIM_PEARSON_BASELINE, false>(h, s, g)));
            break;
        default: rs_set_error("unknown similarity %d", h->p.sim); return RS_ERR_INVALID;
        }
        h->prof.sim_launches++;
        h->prof.total_launches++;
    }
    RS_CUDA(cudaGetLastError());
    return RS_OK;
}

int32_t rs_symmetrize_launch(rs_knn *h) {
    if (!(h->row_begin == 0 && h->row_end == h->n_left)) return RS_OK;
    const unsigned t = (unsigned)((h->n_left + 31) / 32);
    dim3 grid(t, t), block(32, 8);
    symmetrize_kernel<<<grid, block, 0, h->stream>>>(h->sims, h->ld_s, h->n_left, 1);
    h->prof.total_launches++;
    RS_CUDA(cudaGetLastError());
    return RS_OK;
}
