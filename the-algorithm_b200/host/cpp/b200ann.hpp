// b200ann.hpp -- header-only C++17 mirror of the reference interface over the C ABI (include/b200ann.h).
//
// The reference path is Scala (compiled, JVM); this image has no JVM, so the host side a native caller would use is
// restated here in C++ with the reference's names and argument meaning:
//   ann::Metric / ann::Distance ............ ann/src/main/scala/com/twitter/ann/common/Metric.scala:17-36, 76-185
//   ann::EntityEmbedding<T> ................ .../common/Api.scala:21
//   ann::NeighborWithDistance<T> ........... .../common/Api.scala:97
//   ann::Queryable<T> / ann::Appendable<T> . .../common/Api.scala:24-51, 133-145
//   ann::BruteForceIndex<T> ................ .../brute_force/BruteForceIndex.scala:26-92 (apply, append, toQueryable,
//                                            query, queryWithDistance) + batch entry points of the device index
// com.twitter.util.Future becomes std::future; a failed C-ABI call becomes an exception stored in the future
// (futurePool { ... } semantics, BruteForceIndex.scala:49,71).  T must be convertible to/from int64_t (TwHIN uses Long,
// tweet-mixer AnnQueryableByIdModule.scala:38); other id types keep a slot table on the caller's side.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <future>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/b200ann.h"

namespace ann {

enum class Metric : int32_t { L2 = ANN_METRIC_L2, Cosine = ANN_METRIC_COSINE, InnerProduct = ANN_METRIC_INNER_PRODUCT };

// Metric.fromString, Metric.scala:63-73 (EditDistance is a string metric: not on this path)
inline Metric metric_from_string(const std::string& name) {
    if (name == "Cosine") return Metric::Cosine;
    if (name == "L2") return Metric::L2;
    if (name == "InnerProduct") return Metric::InnerProduct;
    throw std::invalid_argument("No Metric with the name " + name);
}

// Ordering.Float.compare (java.lang.Float.compare): -0.0 < +0.0, NaN greatest, all NaN equal.  Metric.scala:22-36
inline int float_compare(float a, float b) {
    auto key = [](float f) -> uint32_t {
        if (f != f) return 0xFFFFFFFFu;
        uint32_t u;
        std::memcpy(&u, &f, 4);
        return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    };
    uint32_t ka = key(a), kb = key(b);
    return (ka > kb) - (ka < kb);
}

struct Distance {
    float distance;
    int compare(const Distance& that) const { return float_compare(distance, that.distance); }
    bool operator<(const Distance& that) const { return compare(that) < 0; }
};

template <typename T>
struct EntityEmbedding {
    T id;
    std::vector<float> embedding;
};

template <typename T>
struct NeighborWithDistance {
    T neighbor;
    Distance distance;
};

struct BruteForceRuntimeParams {};  // object BruteForceRuntimeParams extends RuntimeParams, BruteForceIndex.scala:24

class AnnError : public std::runtime_error {
public:
    AnnError(int code, const std::string& msg) : std::runtime_error("b200ann error " + std::to_string(code) + ": " + msg), code(code) {}
    int code;
};

inline void check(int rc) {
    if (rc != ANN_OK) throw AnnError(rc, ann_last_error());
}

// Metric.distance(embedding1, embedding2) -- Metric.scala:76-86 (L2 :89-94, Cosine :120-125, InnerProduct :153-158) -- on the
// device, with the arithmetic of the query path (ann_distance_pairs).  `absoluteDistance` is the same number here.
inline Distance distance(Metric metric, const std::vector<float>& embedding1, const std::vector<float>& embedding2, int device = 0) {
    if (embedding1.size() != embedding2.size()) throw AnnError(ANN_ERR_DIMENSION_MISMATCH, "embeddings differ in dimension");
    float out = 0.f;
    check(ann_distance_pairs(static_cast<int32_t>(metric), 0u, static_cast<int32_t>(embedding1.size()), embedding1.data(),
                             embedding2.data(), 1, &out, device));
    return Distance{out};
}

// MetricUtil.norm -- Metric.scala:285-289 -- the embedding scaled to unit L2 norm (ann_normalize_rows)
inline std::vector<float> norm(const std::vector<float>& embedding, int device = 0) {
    std::vector<float> out(embedding.size());
    check(ann_normalize_rows(static_cast<int32_t>(embedding.size()), embedding.data(), 1, out.data(), device));
    return out;
}

template <typename T>
class Queryable {
public:
    virtual ~Queryable() = default;
    virtual std::future<std::vector<T>> query(const std::vector<float>& embedding, int numOfNeighbors, BruteForceRuntimeParams = {}) = 0;
    virtual std::future<std::vector<NeighborWithDistance<T>>> queryWithDistance(const std::vector<float>& embedding, int numOfNeighbors,
                                                                                 BruteForceRuntimeParams = {}) = 0;
};

template <typename T>
class Appendable {
public:
    virtual ~Appendable() = default;
    virtual std::future<void> append(const EntityEmbedding<T>& entity) = 0;
    virtual Queryable<T>& toQueryable() = 0;
};

template <typename T = int64_t>
class BruteForceIndex : public Appendable<T>, public Queryable<T> {
public:
    static constexpr const char* DataFileName = "BruteForceFileData";

    // BruteForceIndex.apply(metric, futurePool, initialEmbeddings): the pool is the caller's std::launch policy
    BruteForceIndex(Metric metric, int dim, std::launch policy = std::launch::deferred, int device = 0, int64_t capacity_hint = 0)
        : metric_(metric), dim_(dim), policy_(policy) {
        ann_config cfg{static_cast<int32_t>(metric), dim, capacity_hint, device, 0u};
        check(ann_create(&cfg, &h_));
    }
    ~BruteForceIndex() override { ann_destroy(h_); }
    BruteForceIndex(const BruteForceIndex&) = delete;
    BruteForceIndex& operator=(const BruteForceIndex&) = delete;

    std::future<void> append(const EntityEmbedding<T>& entity) override {
        return std::async(policy_, [this, entity] {
            if (static_cast<int>(entity.embedding.size()) != dim_) throw AnnError(ANN_ERR_DIMENSION_MISMATCH, "embedding dimension");
            int64_t id = static_cast<int64_t>(entity.id);
            check(ann_append_batch(h_, &id, entity.embedding.data(), 1));
        });
    }
    // batched Appendable path: n rows in one host->device copy + one norms/shadow kernel
    void appendBatch(const int64_t* ids, const float* rows, int64_t n) { check(ann_append_batch(h_, ids, rows, n)); }

    Queryable<T>& toQueryable() override { return *this; }

    int64_t size() const {
        int64_t n = 0;
        check(ann_size(h_, &n));
        return n;
    }

    std::future<std::vector<NeighborWithDistance<T>>> queryWithDistance(const std::vector<float>& embedding, int numOfNeighbors,
                                                                         BruteForceRuntimeParams = {}) override {
        return std::async(policy_, [this, embedding, numOfNeighbors] {
            std::vector<NeighborWithDistance<T>> out;
            if (numOfNeighbors <= 0) return out;  // every push is popped again, BruteForceIndex.scala:83-85
            std::vector<int64_t> ids(numOfNeighbors);
            std::vector<float> dist(numOfNeighbors);
            int32_t cnt = 0;
            check(ann_query_batch(h_, embedding.data(), 1, static_cast<int32_t>(embedding.size()), numOfNeighbors, ids.data(), dist.data(), &cnt));
            for (int j = 0; j < cnt; ++j) out.push_back({static_cast<T>(ids[j]), Distance{dist[j]}});
            return out;
        });
    }

    std::future<std::vector<T>> query(const std::vector<float>& embedding, int numOfNeighbors, BruteForceRuntimeParams = {}) override {
        return std::async(policy_, [this, embedding, numOfNeighbors] {
            std::vector<T> out;
            for (auto& n : queryWithDistance(embedding, numOfNeighbors).get()) out.push_back(n.neighbor);
            return out;
        });
    }

    // b queries at once (row-major [b][dim]); outputs [b][k], counts [b]
    void batchQueryWithDistance(const float* queries, int32_t b, int32_t k, int64_t* out_ids, float* out_dist, int32_t* out_count) {
        check(ann_query_batch(h_, queries, b, dim_, k, out_ids, out_dist, out_count));
    }

    // Updatable.update (Api.scala:148-150), batched: overwrite the rows stored at insertion slots `slots` (ann_update_batch)
    void updateBatch(const int64_t* slots, const float* rows, int64_t n) { check(ann_update_batch(h_, slots, rows, n)); }

    // SerializableBruteForceIndex.toDirectory (BruteForceIndex.scala:142-161): BruteForceFileData thrift stream + _SUCCESS
    void toDirectory(const std::string& directory) { check(ann_save_directory(h_, directory.c_str(), ANN_ID_INT64_BE, ANN_LAYOUT_FLOAT_TENSOR)); }

    ann_index* handle() { return h_; }

private:
    Metric metric_;
    int dim_;
    std::launch policy_;
    ann_index* h_ = nullptr;
};

// ShardedAppendable + ComposedQueryable (common/ShardApi.scala:34-48, 58-87) over several GPUs of ONE process: the native
// composed handle (ann_sharded_*).  T = int64 ids only (the merge orders ties by id).
class ShardedBruteForceIndex : public Appendable<int64_t>, public Queryable<int64_t> {
public:
    ShardedBruteForceIndex(Metric metric, int dim, const std::vector<int32_t>& devices, std::launch policy = std::launch::deferred,
                           int64_t capacity_hint = 0)
        : dim_(dim), policy_(policy) {
        ann_config cfg{static_cast<int32_t>(metric), dim, capacity_hint, 0, 0u};
        check(ann_sharded_create(&cfg, devices.data(), static_cast<int32_t>(devices.size()), &h_));
    }
    ~ShardedBruteForceIndex() override { ann_sharded_destroy(h_); }
    ShardedBruteForceIndex(const ShardedBruteForceIndex&) = delete;
    ShardedBruteForceIndex& operator=(const ShardedBruteForceIndex&) = delete;

    std::future<void> append(const EntityEmbedding<int64_t>& entity) override {
        return std::async(policy_, [this, entity] {
            if (static_cast<int>(entity.embedding.size()) != dim_) throw AnnError(ANN_ERR_DIMENSION_MISMATCH, "embedding dimension");
            check(ann_sharded_append_batch(h_, &entity.id, entity.embedding.data(), 1));
        });
    }
    void appendBatch(const int64_t* ids, const float* rows, int64_t n) { check(ann_sharded_append_batch(h_, ids, rows, n)); }
    Queryable<int64_t>& toQueryable() override { return *this; }
    int64_t size() const {
        int64_t n = 0;
        check(ann_sharded_size(h_, &n));
        return n;
    }
    void batchQueryWithDistance(const float* queries, int32_t b, int32_t k, int64_t* out_ids, float* out_dist, int32_t* out_count) {
        check(ann_sharded_query_batch(h_, queries, b, dim_, k, out_ids, out_dist, out_count));
    }
    std::future<std::vector<NeighborWithDistance<int64_t>>> queryWithDistance(const std::vector<float>& embedding, int numOfNeighbors,
                                                                               BruteForceRuntimeParams = {}) override {
        return std::async(policy_, [this, embedding, numOfNeighbors] {
            std::vector<NeighborWithDistance<int64_t>> out;
            if (numOfNeighbors <= 0) return out;
            std::vector<int64_t> ids(numOfNeighbors);
            std::vector<float> dist(numOfNeighbors);
            int32_t cnt = 0;
            check(ann_sharded_query_batch(h_, embedding.data(), 1, static_cast<int32_t>(embedding.size()), numOfNeighbors, ids.data(),
                                          dist.data(), &cnt));
            for (int j = 0; j < cnt; ++j) out.push_back({ids[j], Distance{dist[j]}});
            return out;
        });
    }
    std::future<std::vector<int64_t>> query(const std::vector<float>& embedding, int numOfNeighbors, BruteForceRuntimeParams = {}) override {
        return std::async(policy_, [this, embedding, numOfNeighbors] {
            std::vector<int64_t> out;
            for (auto& n : queryWithDistance(embedding, numOfNeighbors).get()) out.push_back(n.neighbor);
            return out;
        });
    }
    // ShardedSerialization.toDirectory (ShardedSerialization.scala:28-38): shard_<i>/BruteForceFileData, thrift stream
    void toDirectory(const std::string& directory) { check(ann_sharded_save_directory(h_, directory.c_str(), ANN_ID_INT64_BE, ANN_LAYOUT_FLOAT_TENSOR)); }
    ann_sharded_index* handle() { return h_; }

private:
    int dim_;
    std::launch policy_;
    ann_sharded_index* h_ = nullptr;
};

}  // namespace ann
