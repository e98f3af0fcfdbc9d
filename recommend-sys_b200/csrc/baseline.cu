// baseline.cu — ALS baseline estimates on the device (EXTENSION).
//
// BASELINE.json config 3 names "ALS baselines"; the reference only has BaseLine.Fit, a strictly
// sequential SGD (core/base.go:135-163) that the host keeps running unchanged (rs_host_baseline_sgd)
// and whose 20 passes over the ratings dominate Fit at MovieLens-20M size (1.8 s of 2.0 s).
// The ALS form is embarrassingly parallel per row:
//     repeat n_epochs:  b_i = sum_{u in R(i)} (r_ui - mu - b_u) / (reg_i + |R(i)|)   every item
//                       b_u = sum_{i in R(u)} (r_ui - mu - b_i) / (reg_u + |R(u)|)   every user
// One warp per row; the summation order is fixed (terms in dataset order dealt round-robin to
// the 32 lanes, then an xor butterfly) and is the one the CPU checker (or_baseline_als)
// restates, so the biases are bit-identical to the checker.  PARITY UNPINNED: there is no
// reference implementation of ALS baselines.
#include <cub/cub.cuh>

#include "common.cuh"

namespace {

constexpr int T = 256;
inline unsigned blocks_for(int64_t n) { return (unsigned)((n + T - 1) / T); }

__global__ void iota_count_kernel(const int32_t *__restrict__ users, const int32_t *__restrict__ items, int64_t nnz,
                                  int32_t n_users, int32_t n_items, int32_t *idx, unsigned long long *ucount,
                                  unsigned long long *icount, int32_t *bad) {
    const int64_t x = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (x >= nnz) return;
    idx[x] = (int32_t)x;
    const int32_t u = users[x], i = items[x];
    if (u < 0 || u >= n_users || i < 0 || i >= n_items) { *bad = 1; return; }
    atomicAdd(&ucount[u], 1ull);
    atomicAdd(&icount[i], 1ull);
}

__global__ void gather_side_kernel(const int32_t *__restrict__ perm, const int32_t *__restrict__ other,
                                   const double *__restrict__ rating, int64_t nnz, int32_t *o_other, double *o_val) {
    const int64_t x = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (x >= nnz) return;
    const int32_t p = perm[x];
    o_other[x] = other[p];
    o_val[x] = rating[p];
}

// one warp per row: out[row] = sum_t ((val - mu) - other_bias[other]) / (reg + len)
__global__ void als_side_kernel(const unsigned long long *__restrict__ ptr, const int32_t *__restrict__ other,
                                const double *__restrict__ val, int32_t n_rows, double mu, double reg,
                                const double *__restrict__ other_bias, double *__restrict__ out) {
    const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n_rows) return;
    const unsigned long long b = ptr[row], e = ptr[row + 1];
    double part = 0.0;
    for (unsigned long long t = b + lane; t < e; t += 32) {
        const double term = (val[t] - mu) - other_bias[other[t]];
        part += term;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) out[row] = part / (reg + (double)(e - b));
}

struct Blocks {
    int device;
    std::vector<std::pair<void *, size_t>> v;
    cudaStream_t drain = nullptr;   // synchronised before the blocks are released
    ~Blocks() {
        if (drain) cudaStreamSynchronize(drain);
        for (auto &b : v) rs_cached_free(device, b.first, b.second);
    }
    template <typename U> int32_t get(U **out, size_t count) {
        void *p = nullptr;
        size_t got = 0;
        RS_TRY(rs_cached_malloc(device, &p, count * sizeof(U) + 256, &got));
        v.emplace_back(p, got);
        *out = reinterpret_cast<U *>(p);
        return RS_OK;
    }
};

}  // namespace

extern "C" int32_t rs_baseline_als(int32_t device, const int32_t *users, const int32_t *items, const double *ratings,
                                   int64_t nnz, int32_t n_users, int32_t n_items, double global_mean, double reg_u,
                                   double reg_i, int32_t n_epochs, double *user_bias, double *item_bias) {
    if (!users || !items || !ratings || !user_bias || !item_bias || nnz <= 0 || n_users <= 0 || n_items <= 0 ||
        n_epochs < 0) {
        rs_set_error("rs_baseline_als: empty or null input");
        return RS_ERR_INVALID;
    }
    if (nnz >= (1ll << 31)) {
        rs_set_error("rs_baseline_als: nnz %lld exceeds the 2^31-1 entries supported", (long long)nnz);
        return RS_ERR_UNSUPPORTED;
    }
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        rs_set_error("rs_baseline_als: no CUDA device (this library has no CPU fallback)");
        return RS_ERR_CUDA;
    }
    if (device < 0) RS_CUDA(cudaGetDevice(&device));
    RS_CUDA(cudaSetDevice(device));
    cudaStream_t st;
    RS_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    // destroyed AFTER `mem` (declared first): the stream is drained before its blocks go back to the shared
    // cache, so an error return can never hand out memory that queued work still touches
    struct StreamGuard { cudaStream_t s; ~StreamGuard() { cudaStreamSynchronize(s); cudaStreamDestroy(s); } } guard{st};
    Blocks mem{device, {}};
    mem.drain = st;

    int32_t *d_u, *d_i, *idx, *keys, *perm, *u_other, *i_other, *bad;
    double *d_r, *u_val, *i_val, *bu, *bi;
    unsigned long long *ucount, *icount, *uptr, *iptr;
    RS_TRY(mem.get(&d_u, nnz)); RS_TRY(mem.get(&d_i, nnz)); RS_TRY(mem.get(&d_r, nnz));
    RS_TRY(mem.get(&idx, nnz)); RS_TRY(mem.get(&keys, nnz)); RS_TRY(mem.get(&perm, nnz));
    RS_TRY(mem.get(&u_other, nnz)); RS_TRY(mem.get(&i_other, nnz));
    RS_TRY(mem.get(&u_val, nnz)); RS_TRY(mem.get(&i_val, nnz));
    RS_TRY(mem.get(&bu, (size_t)n_users)); RS_TRY(mem.get(&bi, (size_t)n_items));
    RS_TRY(mem.get(&ucount, (size_t)n_users + 1)); RS_TRY(mem.get(&icount, (size_t)n_items + 1));
    RS_TRY(mem.get(&uptr, (size_t)n_users + 1)); RS_TRY(mem.get(&iptr, (size_t)n_items + 1));
    RS_TRY(mem.get(&bad, 1));

    RS_CUDA(cudaMemcpyAsync(d_u, users, (size_t)nnz * 4, cudaMemcpyHostToDevice, st));
    RS_CUDA(cudaMemcpyAsync(d_i, items, (size_t)nnz * 4, cudaMemcpyHostToDevice, st));
    RS_CUDA(cudaMemcpyAsync(d_r, ratings, (size_t)nnz * 8, cudaMemcpyHostToDevice, st));
    RS_CUDA(cudaMemsetAsync(ucount, 0, ((size_t)n_users + 1) * 8, st));
    RS_CUDA(cudaMemsetAsync(icount, 0, ((size_t)n_items + 1) * 8, st));
    RS_CUDA(cudaMemsetAsync(bu, 0, (size_t)n_users * 8, st));
    RS_CUDA(cudaMemsetAsync(bi, 0, (size_t)n_items * 8, st));
    RS_CUDA(cudaMemsetAsync(bad, 0, 4, st));
    iota_count_kernel<<<blocks_for(nnz), T, 0, st>>>(d_u, d_i, nnz, n_users, n_items, idx, ucount, icount, bad);

    int32_t h_bad = 0;
    RS_CUDA(cudaMemcpyAsync(&h_bad, bad, 4, cudaMemcpyDeviceToHost, st));
    RS_CUDA(cudaStreamSynchronize(st));
    if (h_bad) {
        rs_set_error("rs_baseline_als: inner ids outside [0,n_users) x [0,n_items)");
        return RS_ERR_INVALID;
    }
    // CUB temp storage sized for the largest request
    size_t need = 0, need2 = 0;
    RS_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, need, d_u, keys, idx, perm, (int)nnz, 0, 32, st));
    RS_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, need2, ucount, uptr, (n_users > n_items ? n_users : n_items) + 1, st));
    if (need2 > need) need = need2;
    char *tmp;
    RS_TRY(mem.get(&tmp, need));

    auto bits_for = [](int32_t n) { int b = 1; while ((1ll << b) < (long long)n && b < 31) b++; return b; };
    // stable sorts keep the dataset order inside a row (the order the oracle sums in)
    RS_CUDA(cub::DeviceRadixSort::SortPairs(tmp, need, d_u, keys, idx, perm, (int)nnz, 0, bits_for(n_users), st));
    gather_side_kernel<<<blocks_for(nnz), T, 0, st>>>(perm, d_i, d_r, nnz, u_other, u_val);
    RS_CUDA(cub::DeviceRadixSort::SortPairs(tmp, need, d_i, keys, idx, perm, (int)nnz, 0, bits_for(n_items), st));
    gather_side_kernel<<<blocks_for(nnz), T, 0, st>>>(perm, d_u, d_r, nnz, i_other, i_val);
    RS_CUDA(cub::DeviceScan::ExclusiveSum(tmp, need, ucount, uptr, n_users + 1, st));
    RS_CUDA(cub::DeviceScan::ExclusiveSum(tmp, need, icount, iptr, n_items + 1, st));

    for (int ep = 0; ep < n_epochs; ep++) {
        als_side_kernel<<<blocks_for((int64_t)n_items * 32), T, 0, st>>>(iptr, i_other, i_val, n_items, global_mean,
                                                                         reg_i, bu, bi);
        als_side_kernel<<<blocks_for((int64_t)n_users * 32), T, 0, st>>>(uptr, u_other, u_val, n_users, global_mean,
                                                                         reg_u, bi, bu);
    }
    RS_CUDA(cudaMemcpyAsync(user_bias, bu, (size_t)n_users * 8, cudaMemcpyDeviceToHost, st));
    RS_CUDA(cudaMemcpyAsync(item_bias, bi, (size_t)n_items * 8, cudaMemcpyDeviceToHost, st));
    RS_CUDA(cudaStreamSynchronize(st));
    RS_CUDA(cudaGetLastError());
    return RS_OK;
}
