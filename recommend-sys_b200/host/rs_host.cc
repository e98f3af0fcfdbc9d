// rs_host.cc — see rs_host.h.  Host code of the drop-in (not the oracle, not the device path).
#include "rs_host.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <unordered_map>
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

void rs_host_convert_dense(const int32_t *table, int64_t n_table, const int64_t *raw, int64_t n, int32_t *inner_out) {
    for (int64_t x = 0; x < n; x++) {
        const int64_t r = raw[x];
        inner_out[x] = (r >= 0 && r < n_table) ? table[r] : -1;     // -1 = newID (core/data.go:129)
    }
}

}  // extern "C"
