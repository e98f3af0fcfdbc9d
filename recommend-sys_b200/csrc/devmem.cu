// devmem.cu — see common.cuh.  A small best-fit cache of device blocks per device.
#include <cstdlib>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace {
struct Block { int device; void *p; size_t bytes; };
std::mutex g_mu;
std::vector<Block> g_free;
}  // namespace

int32_t rs_cached_malloc(int device, void **out, size_t bytes, size_t *got) {
    {
        std::lock_guard<std::mutex> lk(g_mu);
        int best = -1;
        for (int i = 0; i < (int)g_free.size(); i++) {
            const Block &b = g_free[i];
            if (b.device != device || b.bytes < bytes) continue;
            if (b.bytes > 2 * bytes + (64u << 20)) continue;   // do not waste a huge block on a small request
            if (best < 0 || b.bytes < g_free[best].bytes) best = i;
        }
        if (best >= 0) {
            *out = g_free[best].p;
            if (got) *got = g_free[best].bytes;
            g_free.erase(g_free.begin() + best);
            return RS_OK;
        }
    }
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaErrorMemoryAllocation) {   // give cached blocks back to the driver and retry once
        (void)cudaGetLastError();
        rs_cache_trim();
        e = cudaMalloc(&p, bytes);
    }
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        rs_set_error("cudaMalloc(%zu bytes): %s", bytes, cudaGetErrorString(e));
        return e == cudaErrorMemoryAllocation ? RS_ERR_OOM : RS_ERR_CUDA;
    }
    *out = p;
    if (got) *got = bytes;
    return RS_OK;
}

// Parked bytes are capped (RS_KNN_CACHE_BYTES, default 96 GiB per process — one Netflix-shape arena is
// 30 GB, and a 16 GiB cap made every e2e step re-allocate it: 499 -> 1050 ms): beyond the cap the oldest
// parked blocks go back to the driver, so a co-resident allocator (torch, another library) is not
// starved forever by arenas of estimators that no longer exist.
static size_t cache_cap() {
    static size_t cap = [] {
        const char *e = getenv("RS_KNN_CACHE_BYTES");
        return e ? (size_t)strtoull(e, nullptr, 10) : ((size_t)96 << 30);
    }();
    return cap;
}

void rs_cached_free(int device, void *p, size_t bytes) {
    if (!p) return;
    std::vector<Block> drop;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        g_free.push_back({device, p, bytes});
        size_t total = 0;
        for (const Block &b : g_free) total += b.bytes;
        while (total > cache_cap() && !g_free.empty()) {     // oldest first
            total -= g_free.front().bytes;
            drop.push_back(g_free.front());
            g_free.erase(g_free.begin());
        }
    }
    if (!drop.empty()) {
        int cur = 0;
        cudaGetDevice(&cur);
        for (const Block &b : drop) {
            cudaSetDevice(b.device);
            cudaFree(b.p);
        }
        cudaSetDevice(cur);
    }
}

void rs_cache_trim(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    int cur = 0;
    cudaGetDevice(&cur);
    for (const Block &b : g_free) {
        cudaSetDevice(b.device);
        cudaFree(b.p);
    }
    g_free.clear();
    cudaSetDevice(cur);
}
