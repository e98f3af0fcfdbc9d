// sim_stream.cu — exact-order FP64 similarity kernel ("stream" path).
//
// Reproduces core/sim.go (Cosine :10-25, MSD :28-44, Pearson :47-81) bit for bit: for a left
// row i the reference walks i's entries in ascending right id and, for every other left row
// j that also rated that id, accumulates the three running sums in that order.  Here one
// thread OWNS a group of 8 consecutive j and keeps their accumulators in registers; the CTA
// walks row i's entries in ascending right id c and every thread reads the 8 rating bytes
// RT[c][j..j+8) of the transposed byte matrix (coalesced: 128 threads x 8 B = 1 KB per c).
// Each accumulator therefore receives exactly the reference's terms in the reference's
// order, with the same IEEE operations (the library is built with --fmad=false).
//
// Bound: FP64 pipe + L2 (RT is read |row i| times per CTA column chunk); see DESIGN.md.
#include "common.cuh"

namespace {

constexpr int TS = 256;  // entries of row i staged in shared memory per pass

struct StreamArgs {
    const int64_t *l_ptr;
    const int32_t *l_col;
    const double *l_val;
    const uint8_t *rt;
    int64_t ld_rt;
    const double *lut;
    const double *pmeans;      // Pearson row means
    const double *left_bias;   // PearsonBaseline
    const double *right_bias;
    double global_bias;
    double shrinkage;
    double *sims;
    int64_t ld_s;
    int32_t n_left;
    int64_t row_begin;
    int symmetric;             // 1: only column chunks >= the row's chunk are computed
};

template <int SIM>
__global__ void __launch_bounds__(RS_STREAM_THREADS) sim_stream_kernel(StreamArgs a) {
    const int32_t i = (int32_t)(a.row_begin + blockIdx.y);
    const int64_t j0 = (int64_t)blockIdx.x * RS_STREAM_JC;
    if (a.symmetric && (int64_t)(blockIdx.x + 1) * RS_STREAM_JC <= (int64_t)i) return;

    __shared__ int32_t s_c[TS];
    __shared__ double s_a[TS];
    __shared__ double s_aa[TS];
    __shared__ double s_lut[256];

    const int tid = threadIdx.x;
    const int64_t jb = j0 + (int64_t)tid * RS_STREAM_JPT;
    for (int x = tid; x < 256; x += RS_STREAM_THREADS) s_lut[x] = a.lut[x];

    double accm[RS_STREAM_JPT], accn[RS_STREAM_JPT], accl[RS_STREAM_JPT];
    double bj[RS_STREAM_JPT];  // per-j constant: Pearson mean of row j / baseline of row j
#pragma unroll
    for (int b = 0; b < RS_STREAM_JPT; b++) {
        accm[b] = 0.0; accn[b] = 0.0; accl[b] = 0.0;
        int64_t j = jb + b;
        bj[b] = 0.0;
        if (j < a.n_left) {
            if (SIM == RS_SIM_PEARSON) bj[b] = a.pmeans[j];
            if (SIM == RS_SIM_PEARSON_BASELINE) bj[b] = a.global_bias + a.left_bias[j];
        }
    }
    double ai = 0.0;
    if (SIM == RS_SIM_PEARSON) ai = a.pmeans[i];
    if (SIM == RS_SIM_PEARSON_BASELINE) ai = a.global_bias + a.left_bias[i];

    const int64_t eb = a.l_ptr[i], ee = a.l_ptr[i + 1];
    const uint8_t *rt = a.rt + jb;

    for (int64_t base = eb; base < ee; base += TS) {
        const int cnt = (int)((ee - base) < TS ? (ee - base) : TS);
        __syncthreads();
        for (int x = tid; x < cnt; x += RS_STREAM_THREADS) {
            int32_t c = a.l_col[base + x];
            double v = a.l_val[base + x];
            double ra;
            if (SIM == RS_SIM_PEARSON) ra = v - ai;                               // core/sim.go:73
            else if (SIM == RS_SIM_PEARSON_BASELINE) { double bb = ai + a.right_bias[c]; ra = v - bb; }
            else ra = v;
            s_c[x] = c;
            s_a[x] = ra;
            s_aa[x] = ra * ra;                                                    // core/sim.go:19 / :75
        }
        __syncthreads();

        for (int x0 = 0; x0 < cnt; x0 += 4) {
            uint2 w[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                int x = x0 + u;
                if (x < cnt) w[u] = __ldg(reinterpret_cast<const uint2 *>(rt + (int64_t)s_c[x] * a.ld_rt));
                else w[u] = make_uint2(0u, 0u);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int x = (x0 + u < cnt) ? x0 + u : cnt - 1;
                const uint32_t any = __any_sync(0xffffffffu, (w[u].x | w[u].y) != 0u);
                if (!any) continue;
                const double ra = s_a[x], raa = s_aa[x];
                double rbias = 0.0;
                if (SIM == RS_SIM_PEARSON_BASELINE) rbias = a.right_bias[s_c[x]];
#pragma unroll
                for (int b = 0; b < RS_STREAM_JPT; b++) {
                    const uint32_t code = ((b < 4 ? w[u].x : w[u].y) >> (8 * (b & 3))) & 0xffu;
                    if (code) {
                        const double y = s_lut[code];
                        if (SIM == RS_SIM_COSINE) {
                            accm[b] += raa;                  // m += ir*ir      core/sim.go:19
                            accn[b] += y * y;                // n += jr*jr      core/sim.go:20
                            accl[b] += ra * y;               // l += ir*jr      core/sim.go:21
                        } else if (SIM == RS_SIM_MSD) {
                            const double d = ra - y;
                            accm[b] += d * d;                // sum += (ir-jr)^2 core/sim.go:37
                            accn[b] += 1.0;                  // count++          core/sim.go:38
                        } else if (SIM == RS_SIM_PEARSON) {
                            const double rb = y - bj[b];     // core/sim.go:74
                            accm[b] += raa;                  // core/sim.go:75
                            accn[b] += rb * rb;              // core/sim.go:76
                            accl[b] += ra * rb;              // core/sim.go:77
                        } else {
                            const double bb = bj[b] + rbias;
                            const double rb = y - bb;
                            accm[b] += raa;
                            accn[b] += rb * rb;
                            accl[b] += ra * rb;
                        }
                    }
                }
            }
        }
    }

    // epilogue: one similarity per owned j; row i of the shard, coalesced 64 B per thread
    double *out = a.sims + (int64_t)blockIdx.y * a.ld_s + jb;
#pragma unroll
    for (int b = 0; b < RS_STREAM_JPT; b++) {
        const int64_t j = jb + b;
        if (j >= a.n_left) continue;
        double s;
        if (SIM == RS_SIM_MSD) s = 1.0 / (accm[b] / accn[b] + 1.0);               // core/sim.go:43
        else s = accl[b] / (sqrt(accm[b]) * sqrt(accn[b]));                       // core/sim.go:24 / :80
        if (j == (int64_t)i) s = __longlong_as_double(0x7ff8000000000001ll);      // diagonal stays NaN
        out[b] = s;
    }
}

// PearsonBaseline with shrinkage needs the co-rating count as a fourth accumulator; kept
// as a separate kernel so the three common similarities stay at 24 accumulators.
__global__ void __launch_bounds__(RS_STREAM_THREADS) sim_stream_pb_shrink_kernel(StreamArgs a) {
    const int32_t i = (int32_t)(a.row_begin + blockIdx.y);
    const int64_t j0 = (int64_t)blockIdx.x * RS_STREAM_JC;
    if (a.symmetric && (int64_t)(blockIdx.x + 1) * RS_STREAM_JC <= (int64_t)i) return;
    __shared__ double s_lut[256];
    const int tid = threadIdx.x;
    const int64_t jb = j0 + (int64_t)tid * RS_STREAM_JPT;
    for (int x = tid; x < 256; x += RS_STREAM_THREADS) s_lut[x] = a.lut[x];
    __syncthreads();
    double accm[RS_STREAM_JPT], accn[RS_STREAM_JPT], accl[RS_STREAM_JPT], accc[RS_STREAM_JPT], bj[RS_STREAM_JPT];
#pragma unroll
    for (int b = 0; b < RS_STREAM_JPT; b++) {
        accm[b] = accn[b] = accl[b] = accc[b] = 0.0;
        int64_t j = jb + b;
        bj[b] = (j < a.n_left) ? a.global_bias + a.left_bias[j] : 0.0;
    }
    const double ai = a.global_bias + a.left_bias[i];
    const uint8_t *rt = a.rt + jb;
    for (int64_t e = a.l_ptr[i]; e < a.l_ptr[i + 1]; e++) {
        const int32_t c = a.l_col[e];
        const double rbias = a.right_bias[c];
        const double ba = ai + rbias;
        const double ra = a.l_val[e] - ba;
        const double raa = ra * ra;
        const uint2 w = __ldg(reinterpret_cast<const uint2 *>(rt + (int64_t)c * a.ld_rt));
#pragma unroll
        for (int b = 0; b < RS_STREAM_JPT; b++) {
            const uint32_t code = ((b < 4 ? w.x : w.y) >> (8 * (b & 3))) & 0xffu;
            if (code) {
                const double bb = bj[b] + rbias;
                const double rb = s_lut[code] - bb;
                accm[b] += raa;
                accn[b] += rb * rb;
                accl[b] += ra * rb;
                accc[b] += 1.0;
            }
        }
    }
    double *out = a.sims + (int64_t)blockIdx.y * a.ld_s + jb;
#pragma unroll
    for (int b = 0; b < RS_STREAM_JPT; b++) {
        const int64_t j = jb + b;
        if (j >= a.n_left) continue;
        double s = accl[b] / (sqrt(accm[b]) * sqrt(accn[b]));
        s = (accc[b] - 1.0) / (accc[b] - 1.0 + a.shrinkage) * s;
        if (j == (int64_t)i) s = __longlong_as_double(0x7ff8000000000001ll);
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

int32_t rs_sim_stream_launch(rs_knn *h) {
    StreamArgs a{};
    a.l_ptr = h->l_ptr; a.l_col = h->l_col; a.l_val = h->l_val;
    a.rt = h->rt; a.ld_rt = h->ld_rt; a.lut = h->lut; a.pmeans = h->pmeans;
    a.left_bias = h->left_bias; a.right_bias = h->right_bias; a.global_bias = h->global_bias;
    a.shrinkage = h->p.shrinkage;
    a.sims = h->sims; a.ld_s = h->ld_s; a.n_left = h->n_left; a.row_begin = h->row_begin;
    const int64_t rows = h->row_end - h->row_begin;
    a.symmetric = (h->row_begin == 0 && h->row_end == h->n_left) ? 1 : 0;
    dim3 grid((unsigned)(h->ld_rt / RS_STREAM_JC), (unsigned)rows);
    if (rows <= 0) return RS_OK;
    for (int64_t r0 = 0; r0 < rows; r0 += 65535) {
        StreamArgs s = a;
        int64_t nr = rows - r0 < 65535 ? rows - r0 : 65535;
        s.row_begin = h->row_begin + r0;
        s.sims = h->sims + r0 * h->ld_s;
        dim3 g(grid.x, (unsigned)nr);
        switch (h->p.sim) {
        case RS_SIM_COSINE: sim_stream_kernel<RS_SIM_COSINE><<<g, RS_STREAM_THREADS, 0, h->stream>>>(s); break;
        case RS_SIM_MSD: sim_stream_kernel<RS_SIM_MSD><<<g, RS_STREAM_THREADS, 0, h->stream>>>(s); break;
        case RS_SIM_PEARSON: sim_stream_kernel<RS_SIM_PEARSON><<<g, RS_STREAM_THREADS, 0, h->stream>>>(s); break;
        case RS_SIM_PEARSON_BASELINE:
            if (h->p.shrinkage > 0.0) sim_stream_pb_shrink_kernel<<<g, RS_STREAM_THREADS, 0, h->stream>>>(s);
            else sim_stream_kernel<RS_SIM_PEARSON_BASELINE><<<g, RS_STREAM_THREADS, 0, h->stream>>>(s);
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
    symmetrize_kernel<<<grid, block, 0, h->stream>>>(h->sims, h->ld_s, h->n_left, RS_STREAM_JC);
    h->prof.total_launches++;
    RS_CUDA(cudaGetLastError());
    return RS_OK;
}
