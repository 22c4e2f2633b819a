// loadtest.cu -- concurrent single-vector load against one handle, measured inside the library.
//
// The reference measures its query servers with a load generator that issues one embedding per request from many workers
// and records average / p50 / p90 / p99 latency in microseconds and the achieved rate
// (ann/src/main/scala/com/twitter/ann/service/loadtest/AnnLoadTestWorker.scala:92-115; LoadTestRecorder.scala:114-187).
// This is that measurement for the C ABI: `threads` host threads each issue `calls_per_thread` one-vector
// ann_query_batch calls (what QueryIndexThriftController.scala:39-90 does per RPC), optionally checking every answer
// against expected neighbour ids.  It exercises the micro-batcher of index.cu exactly the way a thread-pooled server does.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/b200ann.h"
#include "kernels.h"

using namespace b200ann;

extern "C" int ann_loadtest(ann_index* ix, const float* queries, int32_t nq, int32_t dim, int32_t k, int32_t threads,
                            int32_t calls_per_thread, const int64_t* expect_ids, ann_load_stats* out) {
    if (!ix || !queries || !out) return report_error(ANN_ERR_NULL_POINTER, "ann_loadtest: NULL argument");
    if (nq < 1 || k < 1 || threads < 1 || threads > 1024 || calls_per_thread < 1)
        return report_error(ANN_ERR_INVALID_ARGUMENT, "ann_loadtest: nq, k, calls_per_thread >= 1 and threads in [1, 1024]");
    memset(out, 0, sizeof(*out));
    int64_t b0 = 0, c0 = 0;
    (void)ann_get_stat(ix, "coalesced_batches", &b0);
    (void)ann_get_stat(ix, "coalesced_calls", &c0);
    std::vector<std::vector<float>> lat((size_t)threads);
    std::atomic<int> first_rc{0};
    std::atomic<long long> mismatches{0};
    std::atomic<int> ready{0};
    std::atomic<bool> go{false};
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([&, t] {
            std::vector<int64_t> ids((size_t)k);
            std::vector<float> dist((size_t)k);
            int32_t cnt = 0;
            lat[(size_t)t].reserve((size_t)calls_per_thread);
            ready.fetch_add(1);
            while (!go.load(std::memory_order_acquire)) std::this_thread::yield();
            for (int c = 0; c < calls_per_thread; ++c) {
                const int q = (int)(((long long)t * calls_per_thread + c) % nq);
                const auto t0 = std::chrono::steady_clock::now();
                const int rc = ann_query_batch(ix, queries + (size_t)q * dim, 1, dim, k, ids.data(), dist.data(), &cnt);
                const auto t1 = std::chrono::steady_clock::now();
                if (rc) {
                    int zero = 0;
                    first_rc.compare_exchange_strong(zero, rc);
                    return;
                }
                lat[(size_t)t].push_back(std::chrono::duration<float, std::micro>(t1 - t0).count());
                if (expect_ids && memcmp(ids.data(), expect_ids + (size_t)q * k, (size_t)k * sizeof(int64_t)) != 0) mismatches.fetch_add(1);
            }
        });
    while (ready.load() < threads) std::this_thread::yield();
    const auto w0 = std::chrono::steady_clock::now();
    go.store(true, std::memory_order_release);
    for (auto& th : pool) th.join();
    const auto w1 = std::chrono::steady_clock::now();
    if (first_rc.load()) return report_error(first_rc.load(), "ann_loadtest: a query call failed");
    std::vector<float> all;
    for (auto& v : lat) all.insert(all.end(), v.begin(), v.end());
    std::sort(all.begin(), all.end());
    const double wall = std::chrono::duration<double>(w1 - w0).count();
    double sum = 0;
    for (float v : all) sum += v;
    auto pct = [&](double p) { return all.empty() ? 0.0 : (double)all[std::min(all.size() - 1, (size_t)(p * all.size()))]; };
    out->calls = (int64_t)all.size();
    out->qps = wall > 0 ? all.size() / wall : 0;
    out->avg_us = all.empty() ? 0 : sum / all.size();
    out->p50_us = pct(0.50);
    out->p90_us = pct(0.90);
    out->p99_us = pct(0.99);
    out->mismatches = mismatches.load();
    int64_t b1 = 0, c1 = 0;
    (void)ann_get_stat(ix, "coalesced_batches", &b1);
    (void)ann_get_stat(ix, "coalesced_calls", &c1);
    out->device_batches = b1 - b0;
    out->wall_seconds = wall;
    return ANN_OK;
}
