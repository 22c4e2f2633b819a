// knn_join.cu -- ann_knn_join: all-pairs exact top-k of a query set against a corpus, host buffers in, host buffers out.
//
// Replaces, for one GPU, what the reference runs on Hadoop: KnnHelper.findNearestNeighbours /
// findNearestNeighboursWithIndexingStrategy (ann/src/main/scala/com/twitter/ann/scalding/offline/KnnHelper.scala:168-215,
// 248-347) builds an index over (a random group of) the search space, queries every query embedding against it and
// keeps k per query, merging the per-group lists with sortedTake (:330-346).  This is the caller that really presents
// query batches in the thousands, so it is written as a pipeline around the hot path rather than as a loop of blocking
// calls:
//   * the corpus is cut into tiles that fit the device (one tile = one index = the reference's "search group");
//   * queries stream through in tiles of `query_tile`; three streams overlap the H2D copy of tile t+1 and the D2H copy
//     of tile t-1 with the query kernels of tile t (double-buffered pinned staging, events between the streams);
//   * with more than one corpus tile the running best of every query travels with its tile and is merged on the device
//     with the K5 merge kernel (canonical (distance, id) order, so the result equals a single index over everything).
// It uses only the public C ABI for the index itself (ann_create / ann_append_batch / ann_query_batch_device /
// ann_merge_topk_device), i.e. it is the first in-tree client of include/b200ann.h.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/b200ann.h"
#include "common.cuh"
#include "kernels.h"

namespace {

struct Slot {
    float* pin_q = nullptr;          // pinned staging: one query tile
    unsigned char* pin_res = nullptr;  // pinned staging: [ids | dist | count | this corpus tile's own counts] of one query tile
    float* d_q = nullptr;
    int64_t* d_ids = nullptr;        // [2][qt][k]: shard 0 = running best from earlier corpus tiles, shard 1 = this tile
    float* d_dist = nullptr;
    int32_t* d_cnt = nullptr;        // [2][qt]
    int64_t* d_mids = nullptr;       // merged [qt][k]
    float* d_mdist = nullptr;
    int32_t* d_mcnt = nullptr;
    cudaEvent_t h2d = nullptr, done = nullptr, d2h = nullptr;
    long long tile = -1;             // query tile whose results sit in pin_res (not yet copied out)
};

struct Join {
    Slot slot[2];
    cudaStream_t s_in = nullptr, s_run = nullptr, s_out = nullptr;
    ann_index* ix = nullptr;
    ~Join() {
        if (ix) ann_destroy(ix);
        for (Slot& s : slot) {
            if (s.pin_q) cudaFreeHost(s.pin_q);
            if (s.pin_res) cudaFreeHost(s.pin_res);
            cudaFree(s.d_q);
            cudaFree(s.d_ids);
            cudaFree(s.d_dist);
            cudaFree(s.d_cnt);
            cudaFree(s.d_mids);
            cudaFree(s.d_mdist);
            cudaFree(s.d_mcnt);
            if (s.h2d) cudaEventDestroy(s.h2d);
            if (s.done) cudaEventDestroy(s.done);
            if (s.d2h) cudaEventDestroy(s.d2h);
        }
        if (s_in) cudaStreamDestroy(s_in);
        if (s_run) cudaStreamDestroy(s_run);
        if (s_out) cudaStreamDestroy(s_out);
    }
};

#define JOIN_CUDA(expr)                                  \
    do {                                                 \
        cudaError_t e_ = (expr);                         \
        if (e_ != cudaSuccess)                           \
            return b200ann::report_error(e_ == cudaErrorMemoryAllocation ? ANN_ERR_OUT_OF_MEMORY : ANN_ERR_CUDA, cudaGetErrorString(e_)); \
    } while (0)

}  // namespace

extern "C" int ann_knn_join(const ann_config* cfg, const int64_t* corpus_ids, const float* corpus_rows, int64_t n,
                            const float* queries, int64_t nq, int32_t k, int64_t corpus_tile_rows, int32_t query_tile,
                            int64_t* out_ids, float* out_dist, int32_t* out_count) {
    using b200ann::report_error;
    if (!cfg) return report_error(ANN_ERR_NULL_POINTER, "ann_knn_join: config is NULL");
    if (n < 0 || nq < 0 || cfg->dim <= 0) return report_error(ANN_ERR_INVALID_ARGUMENT, "ann_knn_join: n < 0, nq < 0 or dim <= 0");
    if (k < 0) return report_error(ANN_ERR_NEGATIVE_K, "ann_knn_join: k < 0");
    if (nq == 0) return ANN_OK;
    if ((n > 0 && (!corpus_ids || !corpus_rows)) || !queries || (k > 0 && (!out_ids || !out_dist)))
        return report_error(ANN_ERR_NULL_POINTER, "ann_knn_join: NULL buffer");
    const int dim = cfg->dim;
    const int kk = std::max(k, 1);
    if (query_tile <= 0) query_tile = 4096;
    query_tile = (int)std::min<int64_t>(query_tile, nq);
    JOIN_CUDA(cudaSetDevice(cfg->device));
    if (k == 0 || n == 0) {   // BruteForceIndex.scala:66-91: k <= 0 or an empty index give empty lists
        for (int64_t i = 0; i < nq * (int64_t)k; ++i) {
            out_ids[i] = -1;
            out_dist[i] = __builtin_inff();
        }
        if (out_count) std::fill(out_count, out_count + nq, 0);
        return ANN_OK;
    }
    if (corpus_tile_rows <= 0) {
        // rows + bf16 shadow + ids + norms per row, 15 % headroom, query scratch (~0.6 GB at 4096 x top-256) set aside
        size_t free_b = 0, total_b = 0;
        JOIN_CUDA(cudaMemGetInfo(&free_b, &total_b));
        const size_t per_row = (size_t)((dim + 3) / 4 * 4) * 4 + (size_t)((dim + 3 + 7) / 8 * 8) * 2 + 24;
        const size_t usable = free_b > ((size_t)3 << 30) ? (size_t)((double)(free_b - ((size_t)2 << 30)) * 0.85) : free_b / 2;
        corpus_tile_rows = (int64_t)std::max<size_t>(1024, usable / per_row);
    }
    corpus_tile_rows = std::min<int64_t>(corpus_tile_rows, n);
    const int64_t n_ctiles = (n + corpus_tile_rows - 1) / corpus_tile_rows;
    const int64_t n_qtiles = (nq + query_tile - 1) / query_tile;
    const size_t qt = (size_t)query_tile;
    const size_t res_bytes = qt * kk * 12 + qt * 8;

    const bool dbg = getenv("B200ANN_DEBUG") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_start = now();
    Join J;
    JOIN_CUDA(cudaStreamCreateWithFlags(&J.s_in, cudaStreamNonBlocking));
    JOIN_CUDA(cudaStreamCreateWithFlags(&J.s_run, cudaStreamNonBlocking));
    JOIN_CUDA(cudaStreamCreateWithFlags(&J.s_out, cudaStreamNonBlocking));
    for (Slot& s : J.slot) {
        JOIN_CUDA(cudaHostAlloc(&s.pin_q, qt * dim * sizeof(float), cudaHostAllocDefault));
        JOIN_CUDA(cudaHostAlloc(&s.pin_res, res_bytes, cudaHostAllocDefault));
        JOIN_CUDA(cudaMalloc(&s.d_q, qt * dim * sizeof(float)));
        JOIN_CUDA(cudaMalloc(&s.d_ids, 2 * qt * kk * sizeof(int64_t)));
        JOIN_CUDA(cudaMalloc(&s.d_dist, 2 * qt * kk * sizeof(float)));
        JOIN_CUDA(cudaMalloc(&s.d_cnt, 2 * qt * sizeof(int32_t)));
        JOIN_CUDA(cudaMalloc(&s.d_mids, qt * kk * sizeof(int64_t)));
        JOIN_CUDA(cudaMalloc(&s.d_mdist, qt * kk * sizeof(float)));
        JOIN_CUDA(cudaMalloc(&s.d_mcnt, qt * sizeof(int32_t)));
        JOIN_CUDA(cudaEventCreateWithFlags(&s.h2d, cudaEventDisableTiming));
        JOIN_CUDA(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
        JOIN_CUDA(cudaEventCreateWithFlags(&s.d2h, cudaEventDisableTiming));
    }
    std::vector<int32_t> cnt_tmp;   // running counts when the caller does not want them
    int32_t* counts = out_count;
    if (!counts) {
        cnt_tmp.assign((size_t)nq, 0);
        counts = cnt_tmp.data();
    }

    // Copy a finished tile's results from its pinned staging into the caller's arrays.  A query the bounded selector
    // could not answer (massive ties, NaN query: include/b200ann.h ANN_ERR_CANDIDATE_OVERFLOW) comes back with an empty
    // list from the asynchronous device call; such a tile is NOT copied out (the caller's arrays keep the running best)
    // and is redone through the blocking host call, which answers those queries with the exact fallback.
    std::vector<int64_t> redo;
    int32_t expect = 0;   // results a healthy query returns for the current corpus tile: min(k, rows)
    auto drain = [&](Slot& s) -> int {
        if (s.tile < 0) return ANN_OK;
        JOIN_CUDA(cudaEventSynchronize(s.d2h));
        const int64_t q0 = s.tile * query_tile;
        const size_t m = (size_t)std::min<int64_t>(query_tile, nq - q0);
        const int32_t* own = reinterpret_cast<const int32_t*>(s.pin_res + qt * kk * 12 + qt * 4);
        bool healthy = true;
        for (size_t i = 0; i < m; ++i) healthy = healthy && own[i] == expect;
        if (!healthy) redo.push_back(s.tile);
        else {
            std::memcpy(out_ids + (size_t)q0 * k, s.pin_res, m * k * sizeof(int64_t));
            std::memcpy(out_dist + (size_t)q0 * k, s.pin_res + qt * kk * 8, m * k * sizeof(float));
            std::memcpy(counts + q0, s.pin_res + qt * kk * 12, m * sizeof(int32_t));
        }
        s.tile = -1;
        return ANN_OK;
    };
    // host merge of two canonical lists (the running best and one corpus tile's answer) for the redo path
    std::vector<int64_t> h_ids;
    std::vector<float> h_dist;
    std::vector<int32_t> h_cnt;
    auto redo_tile = [&](int64_t t, bool first_corpus_tile) -> int {
        const int64_t q0 = t * query_tile;
        const int m = (int)std::min<int64_t>(query_tile, nq - q0);
        h_ids.resize((size_t)m * k);
        h_dist.resize((size_t)m * k);
        h_cnt.resize((size_t)m);
        int rc = ann_query_batch(J.ix, queries + (size_t)q0 * dim, m, dim, k, h_ids.data(), h_dist.data(), h_cnt.data());
        if (rc) return rc;
        std::vector<int64_t> mi((size_t)k);
        std::vector<float> md((size_t)k);
        for (int i = 0; i < m; ++i) {
            int64_t* oi = out_ids + (size_t)(q0 + i) * k;
            float* od = out_dist + (size_t)(q0 + i) * k;
            const int64_t* ni = h_ids.data() + (size_t)i * k;
            const float* nd = h_dist.data() + (size_t)i * k;
            const int cp = first_corpus_tile ? 0 : counts[q0 + i], cn = h_cnt[i];
            int a = 0, b2 = 0, o = 0;
            while (o < k && (a < cp || b2 < cn)) {
                bool take_prev = b2 >= cn;
                if (a < cp && b2 < cn) {
                    const uint32_t ka = b200ann::float_order_key(od[a]), kb = b200ann::float_order_key(nd[b2]);
                    take_prev = ka < kb || (ka == kb && oi[a] <= ni[b2]);
                }
                if (take_prev) {
                    mi[o] = oi[a];
                    md[o++] = od[a++];
                } else {
                    mi[o] = ni[b2];
                    md[o++] = nd[b2++];
                }
            }
            counts[q0 + i] = o;
            for (; o < k; ++o) {
                mi[o] = -1;
                md[o] = __builtin_inff();
            }
            std::memcpy(oi, mi.data(), (size_t)k * sizeof(int64_t));
            std::memcpy(od, md.data(), (size_t)k * sizeof(float));
        }
        return ANN_OK;
    };

    const double t_setup = now();
    for (int64_t c = 0; c < n_ctiles; ++c) {
        const double t_c0 = now();
        const int64_t r0 = c * corpus_tile_rows, rows = std::min<int64_t>(corpus_tile_rows, n - r0);
        ann_config tc = *cfg;
        tc.capacity_hint = rows;
        int rc = ann_create(&tc, &J.ix);
        if (rc) return rc;
        for (int64_t a0 = 0; a0 < rows; a0 += 1 << 20) {   // ShardedAppendable-sized batches: one H2D copy + one K1 launch each
            const int64_t m = std::min<int64_t>(1 << 20, rows - a0);
            rc = ann_append_batch(J.ix, corpus_ids + r0 + a0, corpus_rows + (size_t)(r0 + a0) * dim, m);
            if (rc) return rc;
        }
        expect = (int32_t)std::min<int64_t>(k, rows);
        redo.clear();
        const double t_built = now();
        for (int64_t t = 0; t < n_qtiles; ++t) {
            Slot& s = J.slot[t & 1];
            rc = drain(s);   // the slot's previous tile (t - 2): wait for its D2H, hand the results over
            if (rc) return rc;
            const int64_t q0 = t * query_tile;
            const int m = (int)std::min<int64_t>(query_tile, nq - q0);
            std::memcpy(s.pin_q, queries + (size_t)q0 * dim, (size_t)m * dim * sizeof(float));
            JOIN_CUDA(cudaMemcpyAsync(s.d_q, s.pin_q, (size_t)m * dim * sizeof(float), cudaMemcpyHostToDevice, J.s_in));
            if (c > 0) {   // the running best of these queries goes in as shard 0
                std::memcpy(s.pin_res, out_ids + (size_t)q0 * k, (size_t)m * k * sizeof(int64_t));
                std::memcpy(s.pin_res + qt * kk * 8, out_dist + (size_t)q0 * k, (size_t)m * k * sizeof(float));
                std::memcpy(s.pin_res + qt * kk * 12, counts + q0, (size_t)m * sizeof(int32_t));
                JOIN_CUDA(cudaMemcpyAsync(s.d_ids, s.pin_res, (size_t)m * k * sizeof(int64_t), cudaMemcpyHostToDevice, J.s_in));
                JOIN_CUDA(cudaMemcpyAsync(s.d_dist, s.pin_res + qt * kk * 8, (size_t)m * k * sizeof(float), cudaMemcpyHostToDevice, J.s_in));
                JOIN_CUDA(cudaMemcpyAsync(s.d_cnt, s.pin_res + qt * kk * 12, (size_t)m * sizeof(int32_t), cudaMemcpyHostToDevice, J.s_in));
            }
            JOIN_CUDA(cudaEventRecord(s.h2d, J.s_in));
            JOIN_CUDA(cudaStreamWaitEvent(J.s_run, s.h2d, 0));
            // this tile's answer is shard 1 of the [2][m][k] pair the merge kernel reads
            int64_t* t_ids = s.d_ids + (size_t)m * k;
            float* t_dist = s.d_dist + (size_t)m * k;
            int32_t* t_cnt = s.d_cnt + m;
            rc = ann_query_batch_device(J.ix, s.d_q, m, dim, k, t_ids, t_dist, t_cnt, J.s_run);
            if (rc) return rc;
            const int64_t* r_ids = t_ids;
            const float* r_dist = t_dist;
            const int32_t* r_cnt = t_cnt;
            if (c > 0) {
                rc = ann_merge_topk_device(cfg->device, s.d_ids, s.d_dist, s.d_cnt, 2, m, k, s.d_mids, s.d_mdist, s.d_mcnt, J.s_run);
                if (rc) return rc;
                r_ids = s.d_mids;
                r_dist = s.d_mdist;
                r_cnt = s.d_mcnt;
            }
            JOIN_CUDA(cudaEventRecord(s.done, J.s_run));
            JOIN_CUDA(cudaStreamWaitEvent(J.s_out, s.done, 0));
            JOIN_CUDA(cudaMemcpyAsync(s.pin_res, r_ids, (size_t)m * k * sizeof(int64_t), cudaMemcpyDeviceToHost, J.s_out));
            JOIN_CUDA(cudaMemcpyAsync(s.pin_res + qt * kk * 8, r_dist, (size_t)m * k * sizeof(float), cudaMemcpyDeviceToHost, J.s_out));
            JOIN_CUDA(cudaMemcpyAsync(s.pin_res + qt * kk * 12, r_cnt, (size_t)m * sizeof(int32_t), cudaMemcpyDeviceToHost, J.s_out));
            JOIN_CUDA(cudaMemcpyAsync(s.pin_res + qt * kk * 12 + qt * 4, t_cnt, (size_t)m * sizeof(int32_t), cudaMemcpyDeviceToHost, J.s_out));
            JOIN_CUDA(cudaEventRecord(s.d2h, J.s_out));
            s.tile = t;
        }
        for (Slot& s : J.slot) {
            rc = drain(s);
            if (rc) return rc;
        }
        const double t_queried = now();
        if (!redo.empty()) {   // the asynchronous calls left a sticky overflow status behind: read it away, then redo exactly
            JOIN_CUDA(cudaStreamSynchronize(J.s_run));
            int64_t pending = 0;
            (void)ann_get_stat(J.ix, "pending_error", &pending);
            for (int64_t t : redo) {
                rc = redo_tile(t, c == 0);
                if (rc) return rc;
            }
        }
        ann_destroy(J.ix);
        J.ix = nullptr;
        if (dbg)
            fprintf(stderr, "[b200ann] knn_join corpus tile %lld: setup %.1f ms, build %.1f ms, %lld query tiles %.1f ms, redo %zu tiles + destroy %.1f ms\n",
                    (long long)c, t_setup - t_start, t_built - t_c0, (long long)n_qtiles, t_queried - t_built, redo.size(), now() - t_queried);
    }
    return ANN_OK;
}
