// benchmark.cc — the compiled counterpart of the reference's benchmark.go (benchmark.go:12-52)
// for the path in scope: Slope One and the four KNN estimators through 5-fold cross-validation with
// params = nil (user-based MSD, k = 40, core/knn.go:79-81,145-148), one table row each:
// Name / RMSE / MAE / Time — plus the hot-path device times the library measured.
//
//   benchmark <ratings file: user<TAB>item<TAB>rating per line>     (e.g. ml-100k u.data)
//   benchmark --synthetic USERS ITEMS NNZ
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

#include "core.hpp"

static core::DataSet load_file(const char *path) {  // core/data.go:287-310 (Atoi per field)
    core::DataSet d;
    std::ifstream f(path);
    if (!f) { std::fprintf(stderr, "cannot open %s\n", path); std::exit(1); }
    std::string line;
    while (std::getline(f, line)) {
        std::stringstream ss(line);
        std::string a, b, c;
        std::getline(ss, a, '\t'); std::getline(ss, b, '\t'); std::getline(ss, c, '\t');
        d.Users.push_back(std::atoll(a.c_str()));
        d.Items.push_back(std::atoll(b.c_str()));
        d.Ratings.push_back((double)std::atoll(c.c_str()));
    }
    return d;
}

int main(int argc, char **argv) {
    core::DataSet set;
    if (argc >= 5 && !std::strcmp(argv[1], "--synthetic")) {
        const int users = std::atoi(argv[2]), items = std::atoi(argv[3]);
        const int64_t nnz = std::atoll(argv[4]);
        set.Users.resize(nnz); set.Items.resize(nnz); set.Ratings.resize(nnz);
        const int64_t w = rs_host_synth_ratings(users, items, nnz, 0x5EED0000, set.Users.data(), set.Items.data(),
                                                set.Ratings.data());
        set.Users.resize(w); set.Items.resize(w); set.Ratings.resize(w);
    } else if (argc >= 2) {
        set = load_file(argv[1]);
    } else {
        std::fprintf(stderr, "usage: %s <ratings.tsv> | --synthetic USERS ITEMS NNZ\n", argv[0]);
        return 2;
    }
    struct Row { const char *name; std::unique_ptr<core::Estimator> algo; };
    Row rows[5] = {{"Slope One", core::NewSlopeOne()},      // benchmark.go:26
                   {"KNN", core::NewKNN()}, {"Centered K-NN", core::NewKNNWithMean()},
                   {"K-NN Baseline", core::NewKNNBaseLine()}, {"K-NN Z-Score", core::NewKNNWithZScore()}};
    std::printf("%-16s %10s %10s %12s\n", "Name", "RMSE", "MAE", "Time");
    for (auto &r : rows) {
        const auto t0 = std::chrono::steady_clock::now();
        const auto out = core::CrossValidate(*r.algo, set, {core::RMSE, core::MAE}, 5, 0, core::Parameters{});
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        double rmse = 0, mae = 0;
        for (double v : out[0].Tests) rmse += v / 5;
        for (double v : out[1].Tests) mae += v / 5;
        std::printf("%-16s %10.6f %10.6f %10.1fms\n", r.name, rmse, mae, ms);
    }
    return 0;
}
