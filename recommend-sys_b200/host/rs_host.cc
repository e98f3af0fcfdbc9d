// rs_host.cc — see rs_host.h.  Host code of the drop-in (not the oracle, not the device path).
#include "rs_host.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <string>
#include <cstring>
#include <unordered_map>
#include <thread>
#include <vector>

extern "C" {

int64_t rs_host_inner_ids(const int64_t *raw, int64_t n, int32_t *inner_out) {
    std::unordered_map<int64_t, int32_t> ids;
    ids.reserve((size_t)std::min<int64_t>(n, 1 << 22));
    int32_t count = 0;
    for (int64_t i = 0; i < n; i++) {
        auto it = ids.find(raw[i]);
        if (it == ids.end()) {
            ids.emplace(raw[i], count);
            inner_out[i] = count++;
        } else {
            inner_out[i] = it->second;
        }
    }
    return count;
}

void rs_host_baseline_sgd(const int32_t *iu, const int32_t *ii, const double *rating, int64_t n, int32_t n_users,
                          int32_t n_items, double reg, double lr, int32_t n_epochs, double *user_bias,
                          double *item_bias, double *global_bias_out) {
    std::fill(user_bias, user_bias + n_users, 0.0);
    std::fill(item_bias, item_bias + n_items, 0.0);
    double mu = 0.0;
    for (int32_t ep = 0; ep < n_epochs; ep++) {
        for (int64_t x = 0; x < n; x++) {
            double *bu = &user_bias[iu[x]], *bi = &item_bias[ii[x]];
            const double ubias = *bu, ibias = *bi;
            double est = mu;      // BaseLine.Predict: globalBias, then += userBias, then += itemBias
            est += ubias;
            est += ibias;
            const double diff = est - rating[x];
            mu -= lr * diff;
            *bu -= lr * (diff + reg * ubias);
            *bi -= lr * (diff + reg * ibias);
        }
    }
    if (global_bias_out) *global_bias_out = mu;
}

// ---------------- synthetic data ----------------
namespace {
struct Rng {  // splitmix64-seeded xoshiro256**
    uint64_t s[4];
    static uint64_t splitmix(uint64_t &x) {
        uint64_t z = (x += 0x9e3779b97f4a7c15ull);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
        return z ^ (z >> 31);
    }
    explicit Rng(uint64_t seed) { for (auto &v : s) v = splitmix(seed); }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() {
        uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
        return r;
    }
    double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    double normal() {
        double u1 = uniform(), u2 = uniform();
        if (u1 < 1e-300) u1 = 1e-300;
        return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
    }
    uint64_t below(uint64_t n) { return (uint64_t)(uniform() * (double)n) % n; }
};
}  // namespace

int64_t rs_host_synth_ratings(int32_t n_users, int32_t n_items, int64_t nnz_target, uint64_t seed, int64_t *users,
                              int64_t *items, double *ratings) {
    Rng rng(seed);
    const int64_t cap = (int64_t)n_users * (int64_t)n_items;
    if (nnz_target > cap / 2) nnz_target = cap / 2;
    // user activity: log-normal, floor 20 (MovieLens keeps users with >= 20 ratings), scaled to nnz
    std::vector<double> act((size_t)n_users);
    double tot = 0.0;
    for (auto &a : act) { a = std::exp(1.0 * rng.normal()); tot += a; }
    std::vector<int64_t> deg((size_t)n_users);
    const int64_t floor_deg = std::min<int64_t>(20, std::max<int64_t>(1, nnz_target / n_users));
    int64_t assigned = 0;
    for (int32_t u = 0; u < n_users; u++) {
        int64_t d = (int64_t)std::llround(act[u] / tot * (double)nnz_target);
        d = std::max<int64_t>(d, floor_deg);
        d = std::min<int64_t>(d, (int64_t)n_items * 6 / 10);
        deg[u] = d;
        assigned += d;
    }
    // fix the total up/down deterministically
    for (int32_t u = 0; assigned != nnz_target; u = (u + 1) % n_users) {
        if (assigned < nnz_target && deg[u] < (int64_t)n_items * 6 / 10) { deg[u]++; assigned++; }
        else if (assigned > nnz_target && deg[u] > floor_deg) { deg[u]--; assigned--; }
        else if (assigned > nnz_target && floor_deg <= 1 && deg[u] > 1) { deg[u]--; assigned--; }
    }
    // item popularity: Zipf-like via a cumulative table over a shuffled item order
    std::vector<double> cum((size_t)n_items);
    double z = 0.0;
    for (int32_t i = 0; i < n_items; i++) { z += 1.0 / std::pow((double)(i + 8), 0.9); cum[i] = z; }
    std::vector<int32_t> item_perm((size_t)n_items);
    for (int32_t i = 0; i < n_items; i++) item_perm[i] = i;
    for (int32_t i = n_items - 1; i > 0; i--) std::swap(item_perm[i], item_perm[rng.below((uint64_t)i + 1)]);
    std::vector<double> ub((size_t)n_users), ib((size_t)n_items);
    for (auto &b : ub) b = 0.45 * rng.normal();
    for (auto &b : ib) b = 0.55 * rng.normal();

    std::vector<uint8_t> seen((size_t)n_items, 0);
    std::vector<int32_t> mine;
    int64_t w = 0;
    for (int32_t u = 0; u < n_users; u++) {
        mine.clear();
        const int64_t d = deg[u];
        while ((int64_t)mine.size() < d) {
            const double x = rng.uniform() * z;
            int32_t rank = (int32_t)(std::lower_bound(cum.begin(), cum.end(), x) - cum.begin());
            if (rank >= n_items) rank = n_items - 1;
            int32_t it = item_perm[rank];
            // popularity collisions: probe linearly in rank order (keeps the tail heavy)
            while (seen[it]) { rank = (rank + 1) % n_items; it = item_perm[rank]; }
            seen[it] = 1;
            mine.push_back(it);
        }
        for (int32_t it : mine) {
            seen[it] = 0;
            double r = 3.53 + ub[u] + ib[it] + 0.95 * rng.normal();
            long q = std::lround(r);
            q = std::max(1l, std::min(5l, q));
            users[w] = (int64_t)u + 1;        // raw ids are 1-based like MovieLens
            items[w] = (int64_t)it + 1;
            ratings[w] = (double)q;
            w++;
        }
    }
    // seeded shuffle of the row order (exercises first-appearance inner ids)
    for (int64_t i = w - 1; i > 0; i--) {
        const int64_t j = (int64_t)rng.below((uint64_t)i + 1);
        std::swap(users[i], users[j]);
        std::swap(items[i], items[j]);
        std::swap(ratings[i], ratings[j]);
    }
    return w;
}

// core/data.go:134 `stat.Mean(rowSet.Ratings, nil)`: sum / n.  The CPU restatement used by the tests sums
// sequentially; for integer ratings every order gives the same double, for other ratings gonum's
// summation order is not pinned by any reference test (DESIGN.md "parity unpinned").
double rs_host_mean_seq(const double *x, int64_t n) {
    double sum = 0.0;
    for (int64_t i = 0; i < n; i++) sum += x[i];
    return sum / (double)n;
}

static void convert_dense_range(const int32_t *table, int64_t n_table, const int64_t *raw, int64_t x0, int64_t x1,
                                int32_t *inner_out) {
    for (int64_t x = x0; x < x1; x++) {
        const int64_t r = raw[x];
        inner_out[x] = (r >= 0 && r < n_table) ? table[r] : -1;     // -1 = newID (core/data.go:129)
    }
}

void rs_host_convert_dense(const int32_t *table, int64_t n_table, const int64_t *raw, int64_t n, int32_t *inner_out) {
    convert_dense_range(table, n_table, raw, 0, n, inner_out);
}

// The same over `threads` host threads (a test set of 4 M pairs: 6.5 ms per id column on one core — under a
// Fit sharded over 8 GPUs that is longer than the similarity kernel it is meant to hide behind).
void rs_host_convert_dense_mt(const int32_t *table, int64_t n_table, const int64_t *raw, int64_t n, int32_t *inner_out,
                              int32_t threads) {
    if (threads > 16) threads = 16;
    if (threads < 2 || n < (int64_t)1 << 18) { convert_dense_range(table, n_table, raw, 0, n, inner_out); return; }
    std::vector<std::thread> pool;
    const int64_t per = (n + threads - 1) / threads;
    for (int32_t t = 1; t < threads; t++) {
        const int64_t x0 = per * t < n ? per * t : n, x1 = per * (t + 1) < n ? per * (t + 1) : n;
        pool.emplace_back(convert_dense_range, table, n_table, raw, x0, x1, inner_out);
    }
    convert_dense_range(table, n_table, raw, 0, per < n ? per : n, inner_out);
    for (auto &th : pool) th.join();
}

// Routing of test pairs to the cyclic row shards of a multi-GPU Fit (csrc/common.cuh RS_CYC_B): the pair goes
// to the shard that owns its left row, block b of `block` rows -> shard b % world; pairs with an unknown left id
// (cold start, answered with GlobalMean by any shard) are spread round-robin.  One counting-sort pass:
// order[] = pair indices grouped by shard (stable), counts[world].  4 M pairs: ~10 ms (numpy: 150 ms).
void rs_host_route_pairs(const int32_t *left_inner, int64_t n, int32_t world, int32_t block, int64_t *order,
                         int64_t *counts) {
    // owner of every block of rows, by table (two integer divisions per pair cost more than the rest of the pass)
    int32_t max_left = -1;
    for (int64_t i = 0; i < n; i++) max_left = left_inner[i] > max_left ? left_inner[i] : max_left;
    std::vector<uint8_t> own_of_row((size_t)(max_left + 1) + 1);
    for (int32_t l = 0; l <= max_left; l++) own_of_row[(size_t)l] = (uint8_t)((l / block) % world);
    std::vector<uint8_t> own((size_t)n);
    std::vector<int64_t> off((size_t)world + 1, 0);
    int32_t cold = 0;
    for (int64_t i = 0; i < n; i++) {
        const int32_t l = left_inner[i];
        uint8_t o;
        if (l < 0) { o = (uint8_t)cold; cold = cold + 1 == world ? 0 : cold + 1; }
        else o = own_of_row[(size_t)l];
        own[(size_t)i] = o;
        off[(size_t)o + 1]++;
    }
    for (int32_t r = 0; r < world; r++) { counts[r] = off[(size_t)r + 1]; off[(size_t)r + 1] += off[r]; }
    for (int64_t i = 0; i < n; i++) order[off[own[(size_t)i]]++] = i;
}

// ---- on-disk neighbour lists and a rating loader that keeps half-stars ----
// File format (little endian):  magic "RSKNNL01" | int64 n_rows | int32 k | int32 reserved |
// uint64 checksum (FNV-1a 64 of the payload) | int32 idx[n_rows*k] | float64 sim[n_rows*k].
// idx -1 / sim NaN = unused slot (rs_knn_topk).  Return codes: 0 ok, -1 io error, -2 bad file.
static uint64_t fnv1a(const void *p, size_t n, uint64_t h) {
    const unsigned char *b = (const unsigned char *)p;
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}

int32_t rs_host_save_neighbors(const char *path, int64_t n_rows, int32_t k, const int32_t *idx, const double *sim) {
    FILE *f = fopen(path, "wb");
    if (!f) return -1;
    const size_t cells = (size_t)n_rows * (size_t)k;
    uint64_t sum = fnv1a(idx, cells * 4, 1469598103934665603ull);
    sum = fnv1a(sim, cells * 8, sum);
    const int32_t reserved = 0;
    bool ok = fwrite("RSKNNL01", 1, 8, f) == 8 && fwrite(&n_rows, 8, 1, f) == 1 && fwrite(&k, 4, 1, f) == 1 &&
              fwrite(&reserved, 4, 1, f) == 1 && fwrite(&sum, 8, 1, f) == 1 &&
              fwrite(idx, 4, cells, f) == cells && fwrite(sim, 8, cells, f) == cells;
    ok = (fclose(f) == 0) && ok;
    return ok ? 0 : -1;
}

// Two-step load: header first (idx == NULL: only n_rows / k are returned), then the payload.
int32_t rs_host_load_neighbors(const char *path, int64_t *n_rows, int32_t *k, int32_t *idx, double *sim) {
    FILE *f = fopen(path, "rb");
    if (!f) return -1;
    char magic[8];
    int32_t reserved = 0;
    uint64_t sum = 0;
    if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "RSKNNL01", 8) != 0 || fread(n_rows, 8, 1, f) != 1 ||
        fread(k, 4, 1, f) != 1 || fread(&reserved, 4, 1, f) != 1 || fread(&sum, 8, 1, f) != 1 || *n_rows < 0 || *k < 1) {
        fclose(f);
        return -2;
    }
    if (!idx || !sim) { fclose(f); return 0; }
    const size_t cells = (size_t)*n_rows * (size_t)*k;
    const bool ok = fread(idx, 4, cells, f) == cells && fread(sim, 8, cells, f) == cells;
    fclose(f);
    if (!ok) return -2;
    uint64_t got = fnv1a(idx, cells * 4, 1469598103934665603ull);
    got = fnv1a(sim, cells * 8, got);
    return got == sum ? 0 : -2;
}

// Rating file -> COO.  core/data.go:287-310 pushes field 2 through strconv.Atoi, so "3.5" and header
// lines silently become 0; `float_ratings` != 0 parses the rating with strtod instead (half-stars kept)
// and `skip_header` drops the first line.  Fields 0 and 1 keep the reference's Atoi semantics (0 when
// not an integer).  Pass users == NULL to count the rows.  Returns the number of rows, -1 on io error.
static long long atoi_go(const char *b, const char *e) {      // strconv.Atoi: optional sign + digits only, else 0
    const char *p = b;
    bool neg = false;
    if (p < e && (*p == '+' || *p == '-')) { neg = *p == '-'; p++; }
    if (p == e) return 0;
    long long v = 0;
    for (; p < e; p++) {
        if (*p < '0' || *p > '9') return 0;
        v = v * 10 + (*p - '0');
    }
    return neg ? -v : v;
}

int64_t rs_host_load_ratings(const char *path, const char *sep, int32_t float_ratings, int32_t skip_header,
                             int64_t *users, int64_t *items, double *ratings, int64_t cap) {
    FILE *f = fopen(path, "rb");
    if (!f) return -1;
    const size_t sl = strlen(sep);
    std::string line;
    int64_t n = 0;
    bool first = true;
    int ch;
    auto flush = [&]() {
        if (first && skip_header) { first = false; line.clear(); return; }
        first = false;
        const char *b = line.data(), *e = b + line.size();
        if (e > b && e[-1] == '\r') e--;
        const char *fb[3], *fe[3];
        int nf = 0;
        const char *p = b;
        while (nf < 3) {
            const char *q = sl ? std::search(p, e, sep, sep + sl) : e;
            fb[nf] = p; fe[nf] = q; nf++;
            if (q == e) break;
            p = q + sl;
        }
        if (nf == 3) {        // the reference indexes fields[2] and would panic on a shorter line: such lines are skipped
            if (users && n < cap) {
                users[n] = atoi_go(fb[0], fe[0]);
                items[n] = atoi_go(fb[1], fe[1]);
                if (float_ratings) {
                    std::string t(fb[2], fe[2]);
                    char *end = nullptr;
                    const double v = strtod(t.c_str(), &end);
                    ratings[n] = (end && end != t.c_str()) ? v : 0.0;
                } else {
                    ratings[n] = (double)atoi_go(fb[2], fe[2]);
                }
            }
            n++;
        }
        line.clear();
    };
    while ((ch = fgetc(f)) != EOF) {
        if (ch == '\n') flush(); else line.push_back((char)ch);
    }
    if (!line.empty()) flush();
    fclose(f);
    return n;
}

}  // extern "C"
