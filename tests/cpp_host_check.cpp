// Compile-and-link check of the C++ host mirror (the-algorithm_b200/host/cpp/b200ann.hpp) against libb200ann.so.
// Without a GPU every compute entry point must fail loudly: the constructor throws AnnError(ANN_ERR_NO_DEVICE).
// With a GPU it runs a three-row known-answer query (InnerProduct distances 1 - a.b).
#include <cuda_runtime.h>

#include <cstdio>

#include "../the-algorithm_b200/host/cpp/b200ann.hpp"

int main() {
    using namespace ann;
    if (metric_from_string("Cosine") != Metric::Cosine) return 10;
    if (!(Distance{-0.0f} < Distance{0.0f}) || !(Distance{1e30f} < Distance{std::numeric_limits<float>::quiet_NaN()})) return 11;
    try {
        BruteForceIndex<int64_t> ix(Metric::InnerProduct, 2);
        const int64_t ids[3] = {7, 8, 9};
        const float rows[6] = {1, 0, 0, 1, 2, 0};
        ix.appendBatch(ids, rows, 3);
        auto res = ix.queryWithDistance({1.0f, 0.0f}, 2).get();
        if (res.size() != 2 || res[0].neighbor != 9 || res[0].distance.distance != -1.0f || res[1].neighbor != 7) return 12;
        auto only = ix.query({1.0f, 0.0f}, 0).get();
        if (!only.empty()) return 13;
        if (distance(Metric::L2, {0.0f, 0.0f}, {3.0f, 4.0f}).distance != 5.0f) return 15;           // Metric.scala:89-94
        if (distance(Metric::InnerProduct, {1.0f, 2.0f, 3.0f}, {4.0f, 5.0f, 6.0f}).distance != -31.0f) return 16;
        auto unit = norm({3.0f, 4.0f});                                                               // Metric.scala:285-289
        if (unit.size() != 2 || unit[0] != 0.6f || unit[1] != 0.8f) return 17;
        // one process, several shards (two GPUs when the box has them, else two shards on one): ComposedQueryable semantics,
        // ShardApi.scala:72-86 -- the merged answer equals the single index's, ties by id
        {
            int ndev = 0;
            cudaGetDeviceCount(&ndev);
            ShardedBruteForceIndex sx(Metric::L2, 3, {0, ndev > 1 ? 1 : 0});
            const int n = 5000;
            std::vector<int64_t> sids(n);
            std::vector<float> srows((size_t)n * 3);
            for (int i = 0; i < n; ++i) {
                sids[i] = 100000 - i;
                srows[(size_t)i * 3 + 0] = (float)(i % 71);
                srows[(size_t)i * 3 + 1] = (float)(i % 13);
                srows[(size_t)i * 3 + 2] = (float)(i / 1000);
            }
            sx.appendBatch(sids.data(), srows.data(), n);
            BruteForceIndex<int64_t> one(Metric::L2, 3);
            one.appendBatch(sids.data(), srows.data(), n);
            const float q[6] = {3.f, 4.f, 1.f, 70.f, 12.f, 4.f};
            int64_t a_ids[20], b_ids[20];
            float a_d[20], b_d[20];
            int32_t a_c[2], b_c[2];
            sx.batchQueryWithDistance(q, 2, 10, a_ids, a_d, a_c);
            one.batchQueryWithDistance(q, 2, 10, b_ids, b_d, b_c);
            if (sx.size() != n || a_c[0] != 10 || a_c[1] != 10) return 18;
            for (int i = 0; i < 20; ++i)
                if (a_ids[i] != b_ids[i] || a_d[i] != b_d[i]) return 19;
            if (sx.query({3.f, 4.f, 1.f}, 1).get().at(0) != a_ids[0]) return 20;
        }
        std::printf("gpu ok\n");
        return 0;
    } catch (const AnnError& e) {
        if (e.code == ANN_ERR_NO_DEVICE || e.code == ANN_ERR_CUDA) {
            std::printf("no device: %s\n", e.what());
            return 0;
        }
        std::printf("unexpected: %s\n", e.what());
        return 14;
    }
}
