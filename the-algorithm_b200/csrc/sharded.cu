// sharded.cu -- ONE process, R CUDA devices: ShardedAppendable + ComposedQueryable as a single C-ABI handle.
//
// The reference shards an index inside one JVM: ShardedAppendable routes every appended row to one of S sub-indices and
// ComposedQueryable fans a query out to all of them with Future.collect, flattens the S*k results, sorts and takes k
// (ann/src/main/scala/com/twitter/ann/common/ShardApi.scala:34-48, 58-87).  This is that object for the GPUs of one box, the
// only form a single-JVM host can call: `ann_sharded_create` builds one shard (an ordinary ann_index) per device and enables
// peer access between them, appends are cut into R contiguous parts that travel to their devices in parallel over separate
// PCIe links, and a query batch runs the three-phase sharded query of index.cu on every device at once:
//
//     worker s (one host thread per device, SPMD):
//       H2D queries -> seed (publish k bounds from a prefix)            record E1[s]
//       -- host barrier; stream s waits for every E1[t] --
//       filter (global seed threshold, tcgen05 chunks, publish k best)  record E2[s]
//       -- host barrier; stream s waits for every E2[t] --
//       rescore (exact fp64 finalize of this shard's share of the global survivors -> local result block)  record E3[s]
//       -- host barrier; stream s waits for every E3[t] --
//       exchange+merge of this worker's 1/R slice of the batch (P2P loads from every shard's block) -> D2H of the slice
//
// Cross-device ordering is CUDA events (cudaStreamWaitEvent works across devices of one process): no spinning kernels, no
// NCCL, no host synchronisation until the slice has landed in the caller's buffers.  The per-device code is the same C ABI
// the one-process-per-GPU route (ann/distributed.py under torchrun) drives; only the barrier differs.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/b200ann.h"
#include "kernels.h"

using namespace b200ann;

namespace {

// sense-reversing barrier for the R worker threads: the waits only cover the few microseconds between the workers'
// enqueue calls (never GPU completion), so spin first and yield after a while
class SpinBarrier {
public:
    explicit SpinBarrier(int n) : n_(n) {}
    void wait() {
        const int gen = gen_.load(std::memory_order_acquire);
        if (count_.fetch_add(1, std::memory_order_acq_rel) + 1 == n_) {
            count_.store(0, std::memory_order_relaxed);
            gen_.store(gen + 1, std::memory_order_release);
            return;
        }
        int spins = 0;
        while (gen_.load(std::memory_order_acquire) == gen)
            if (++spins > 2000) std::this_thread::yield();
    }

private:
    const int n_;
    std::atomic<int> count_{0};
    std::atomic<int> gen_{0};
};

// R persistent threads, one per device; run(fn) executes fn(s) on every worker and returns when all are done
class WorkerPool {
public:
    explicit WorkerPool(int n) : n_(n) {
        for (int s = 0; s < n; ++s) threads_.emplace_back([this, s] { loop(s); });
    }
    ~WorkerPool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
            ++gen_;
        }
        cv_.notify_all();
        for (auto& t : threads_) t.join();
    }
    void run(const std::function<void(int)>& fn) {
        {
            std::lock_guard<std::mutex> lk(mu_);
            fn_ = &fn;
            pending_ = n_;
            ++gen_;
        }
        cv_.notify_all();
        std::unique_lock<std::mutex> lk(mu_);
        done_cv_.wait(lk, [this] { return pending_ == 0; });
        fn_ = nullptr;
    }

private:
    void loop(int s) {
        long long seen = 0;
        for (;;) {
            const std::function<void(int)>* fn;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
                fn = fn_;
            }
            (*fn)(s);
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (--pending_ == 0) done_cv_.notify_one();
            }
        }
    }
    const int n_;
    std::vector<std::thread> threads_;
    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    const std::function<void(int)>* fn_ = nullptr;
    long long gen_ = 0;
    int pending_ = 0;
    bool stop_ = false;
};

struct Shard {
    int device = 0;
    ann_index* ix = nullptr;
    cudaStream_t st = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};   // phases 0..2 of the query, 3 = query batch gathered
    // per-(b, k) scratch on this device
    float* d_q = nullptr;
    uint32_t* d_seed = nullptr;          // receive buffer [world][b*k]: block s holds shard s's seed bounds (pushed by shard s)
    uint32_t* d_kth = nullptr;           // likewise for the k best bounds of the second round
    unsigned char* d_local = nullptr;     // result block of this shard's candidates: [ids b*k][dist b*k][count b]
    int64_t* d_sl_ids = nullptr;          // merged slice
    float* d_sl_dist = nullptr;
    int32_t* d_sl_cnt = nullptr;
    size_t cap_q = 0, cap_bk = 0, cap_b = 0, cap_slice = 0;
    int64_t n = 0;
};

}  // namespace

struct ann_sharded_index {
    ann_config cfg{};
    int world = 0;
    std::vector<Shard> sh;
    std::mutex mu;            // one append / query at a time on the composed handle (its shards are busy together anyway)
    WorkerPool* pool = nullptr;
    SpinBarrier* bar = nullptr;
    bool peer_ok = true;      // every pair of devices can map each other; else the merge is staged through the host
    int two_round = 1;        // 1 = share the k best bounds after the last chunk (slice-sized rescoring), 0 = seed round only
    int sliced_seeds = 1;     // two_round only: 1 = every shard seeds ITS slice of the batch over R x the rows and delivers one
                              // bound per query (ann_query_seed_slice_push_device), 0 = every shard seeds every query and
                              // delivers k bounds per query (ann_query_seed_push_device)
    long long queries = 0, fallback_batches = 0;
};

namespace {

#define SH_TRY(expr)                                                                                          \
    do {                                                                                                      \
        cudaError_t _e = (expr);                                                                              \
        if (_e != cudaSuccess) {                                                                              \
            char _b[384];                                                                                     \
            snprintf(_b, sizeof(_b), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            (void)cudaGetLastError();                                                                         \
            return report_error(_e == cudaErrorMemoryAllocation ? ANN_ERR_OUT_OF_MEMORY : ANN_ERR_CUDA, _b);  \
        }                                                                                                     \
    } while (0)

template <typename T>
cudaError_t regrow(T** p, size_t* cap, size_t want) {
    if (want <= *cap) return cudaSuccess;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    cudaError_t e = cudaMalloc(p, want * sizeof(T));
    if (e == cudaSuccess) *cap = want;
    return e;
}

int ensure_scratch(Shard& s, int b, int dim, int k, int q_count, int world) {
    SH_TRY(cudaSetDevice(s.device));
    const size_t bk = (size_t)b * std::max(k, 1);
    SH_TRY(regrow(&s.d_q, &s.cap_q, (size_t)b * dim));
    if (bk > s.cap_bk) {
        if (s.d_seed) cudaFree(s.d_seed);
        if (s.d_kth) cudaFree(s.d_kth);
        if (s.d_local) cudaFree(s.d_local);
        s.d_seed = s.d_kth = nullptr;
        s.d_local = nullptr;
        s.cap_bk = 0;
        SH_TRY(cudaMalloc(&s.d_seed, bk * 4 * world));
        SH_TRY(cudaMalloc(&s.d_kth, bk * 4 * world));
        SH_TRY(cudaMalloc(&s.d_local, bk * 12 + (size_t)b * 4 + 256));
        s.cap_bk = bk;
    }
    const size_t sl = (size_t)std::max(q_count, 1) * std::max(k, 1);
    if (sl > s.cap_slice) {
        if (s.d_sl_ids) cudaFree(s.d_sl_ids);
        if (s.d_sl_dist) cudaFree(s.d_sl_dist);
        if (s.d_sl_cnt) cudaFree(s.d_sl_cnt);
        s.d_sl_ids = nullptr;
        s.d_sl_dist = nullptr;
        s.d_sl_cnt = nullptr;
        s.cap_slice = 0;
        SH_TRY(cudaMalloc(&s.d_sl_ids, sl * 8));
        SH_TRY(cudaMalloc(&s.d_sl_dist, sl * 4));
        SH_TRY(cudaMalloc(&s.d_sl_cnt, (size_t)std::max(q_count, 1) * 4 + 16));
        s.cap_slice = sl;
    }
    return ANN_OK;
}

inline void slice_of(int s, int world, int b, int* q0, int* qn) {
    const long long lo = (long long)s * b / world, hi = (long long)(s + 1) * b / world;
    *q0 = (int)lo;
    *qn = (int)(hi - lo);
}

}  // namespace

extern "C" {

int ann_sharded_create(const ann_config* cfg, const int32_t* device_ids, int32_t n_devices, ann_sharded_index** out) {
    if (!cfg || !out) return report_error(ANN_ERR_NULL_POINTER, "ann_sharded_create: cfg/out is NULL");
    *out = nullptr;
    if (n_devices < 1 || n_devices > kMaxPeers) return report_error(ANN_ERR_INVALID_ARGUMENT, "ann_sharded_create: n_devices must be in [1, 16]");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        (void)cudaGetLastError();
        return report_error(ANN_ERR_NO_DEVICE, "ann_sharded_create: no CUDA device (this engine has no CPU fallback)");
    }
    std::vector<int> devs(n_devices);
    for (int s = 0; s < n_devices; ++s) {
        devs[s] = device_ids ? device_ids[s] : s;
        if (devs[s] < 0 || devs[s] >= ndev) return report_error(ANN_ERR_INVALID_ARGUMENT, "ann_sharded_create: bad device ordinal");
    }
    ann_sharded_index* sx = new (std::nothrow) ann_sharded_index();
    if (!sx) return report_error(ANN_ERR_OUT_OF_MEMORY, "ann_sharded_create: host allocation failed");
    sx->cfg = *cfg;
    sx->world = n_devices;
    sx->sh.resize(n_devices);
    auto cleanup = [&](int rc) {
        ann_sharded_destroy(sx);
        return rc;
    };
    for (int s = 0; s < n_devices; ++s) {
        Shard& sh = sx->sh[s];
        sh.device = devs[s];
        ann_config c = *cfg;
        c.device = devs[s];
        c.capacity_hint = cfg->capacity_hint > 0 ? (cfg->capacity_hint + n_devices - 1) / n_devices : 0;
        int rc = ann_create(&c, &sh.ix);
        if (rc) return cleanup(rc);
        if (cudaSetDevice(sh.device) != cudaSuccess || cudaStreamCreateWithFlags(&sh.st, cudaStreamNonBlocking) != cudaSuccess)
            return cleanup(report_error(ANN_ERR_CUDA, "ann_sharded_create: stream creation failed"));
        for (auto& e : sh.ev)
            if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess)
                return cleanup(report_error(ANN_ERR_CUDA, "ann_sharded_create: event creation failed"));
    }
    // every device maps every other one (NVLink / NVSwitch peer access); without it the exchange is staged through the host
    for (int s = 0; s < n_devices && sx->peer_ok; ++s) {
        cudaSetDevice(devs[s]);
        for (int t = 0; t < n_devices; ++t) {
            if (t == s || devs[t] == devs[s]) continue;   // a device may be listed twice (several shards on one GPU: tests)
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, devs[s], devs[t]) != cudaSuccess || !can) {
                sx->peer_ok = false;
                break;
            }
            cudaError_t e = cudaDeviceEnablePeerAccess(devs[t], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                sx->peer_ok = false;
                break;
            }
            (void)cudaGetLastError();
        }
    }
    (void)cudaGetLastError();
    if (n_devices > 1) {
        sx->pool = new WorkerPool(n_devices);
        sx->bar = new SpinBarrier(n_devices);
    }
    *out = sx;
    return ANN_OK;
}

void ann_sharded_destroy(ann_sharded_index* sx) {
    if (!sx) return;
    delete sx->pool;
    delete sx->bar;
    for (Shard& s : sx->sh) {
        cudaSetDevice(s.device);
        if (s.st) cudaStreamSynchronize(s.st);
        cudaFree(s.d_q);
        cudaFree(s.d_seed);
        cudaFree(s.d_kth);
        cudaFree(s.d_local);
        cudaFree(s.d_sl_ids);
        cudaFree(s.d_sl_dist);
        cudaFree(s.d_sl_cnt);
        for (auto& e : s.ev)
            if (e) cudaEventDestroy(e);
        if (s.st) cudaStreamDestroy(s.st);
        ann_destroy(s.ix);
    }
    (void)cudaGetLastError();
    delete sx;
}

int ann_sharded_size(const ann_sharded_index* sx, int64_t* n) {
    if (!sx || !n) return report_error(ANN_ERR_NULL_POINTER, "ann_sharded_size: NULL argument");
    int64_t t = 0;
    for (const Shard& s : sx->sh) t += s.n;
    *n = t;
    return ANN_OK;
}

int ann_sharded_shard(ann_sharded_index* sx, int32_t shard, ann_index** out, int64_t* rows) {
    if (!sx || !out) return report_error(ANN_ERR_NULL_POINTER, "ann_sharded_shard: NULL argument");
    if (shard < 0 || shard >= sx->world) return report_error(ANN_ERR_INVALID_ARGUMENT, "ann_sharded_shard: no such shard");
    *out = sx->sh[shard].ix;
    if (rows) *rows = sx->sh[shard].n;
    return ANN_OK;
}

int ann_sharded_set_option(ann_sharded_index* sx, const char* name, int64_t value) {
    if (!sx || !name) return report_error(ANN_ERR_NULL_POINTER, "ann_sharded_set_option: NULL argument");
    std::lock_guard<std::mutex> lk(sx->mu);
    if (!strcmp(name, "sliced_seeds")) {
        sx->sliced_seeds = value ? 1 : 0;
        return ANN_OK;
    }
    if (!strcmp(name, "two_round")) {
        sx->two_round = value ? 1 : 0;
        return ANN_OK;
    }
    for (Shard& s : sx->sh) {   // everything else is a per-shard option
        int rc = ann_set_option(s.ix, name, value);
        if (rc) return rc;
    }
    return ANN_OK;
}

int ann_sharded_get_stat(const ann_sharded_index* sx, const char* name, int64_t* value) {
    if (!sx || !name || !value) return report_error(ANN_ERR_NULL_POINTER, "ann_sharded_get_stat: NULL argument");
    if (!strcmp(name, "shards")) *value = sx->world;
    else if (!strcmp(name, "peer_access")) *value = sx->peer_ok ? 1 : 0;
    else if (!strcmp(name, "fallback_batches")) *value = sx->fallback_batches;
    else if (!strcmp(name, "queries")) *value = sx->queries;
    else if (!strcmp(name, "dim")) *value = sx->cfg.dim;   // one dimension, not a sum over the shards
    else {   // sums over the shards ("launches", "row_bytes", ...)
        int64_t t = 0;
        for (const Shard& s : sx->sh) {
            int64_t v = 0;
            int rc = ann_get_stat(s.ix, name, &v);
            if (rc) return rc;
            t += v;
        }
        *value = t;
    }
    return ANN_OK;
}

// ShardedAppendable.append (ShardApi.scala:34-48), batched: the batch is cut into `world` contiguous parts, part s goes to
// shard s (deterministic and balanced, where the reference draws Random.nextInt(shards) per row, ShardApi.scala:21-25).
int ann_sharded_append_batch(ann_sharded_index* sx, const int64_t* ids, const float* rows, int64_t n) {
    if (!sx) return report_error(ANN_ERR_NULL_POINTER, "ann_sharded_append_batch: index is NULL");
    if (n < 0) return report_error(ANN_ERR_INVALID_ARGUMENT, "ann_sharded_append_batch: n < 0");
    if (n == 0) return ANN_OK;
    if (!rows) return report_error(ANN_ERR_NULL_POINTER, "ann_sharded_append_batch: rows is NULL");
    std::lock_guard<std::mutex> lk(sx->mu);
    const int R = sx->world, dim = sx->cfg.dim;
    int64_t base = 0;
    for (const Shard& s : sx->sh) base += s.n;
    std::vector<int> rcs(R, ANN_OK);
    std::vector<std::string> msgs(R);
    auto part = [&](int s) {
        const int64_t lo = (int64_t)s * n / R, hi = (int64_t)(s + 1) * n / R;
        if (hi <= lo) return;
        std::vector<int64_t> gen;
        const int64_t* pid = ids ? ids + lo : nullptr;
        if (!ids) {   // ids default to the insertion index of the COMPOSED index, not of the shard
            gen.resize((size_t)(hi - lo));
            for (int64_t i = lo; i < hi; ++i) gen[(size_t)(i - lo)] = base + i;
            pid = gen.data();
        }
        rcs[s] = ann_append_batch(sx->sh[s].ix, pid, rows + (size_t)lo * dim, hi - lo);
        if (rcs[s]) msgs[s] = ann_last_error();
        else sx->sh[s].n += hi - lo;
    };
    if (R == 1) part(0);
    else sx->pool->run(part);
    for (int s = 0; s < R; ++s)
        if (rcs[s]) return report_error(rcs[s], msgs[s].c_str());
    return ANN_OK;
}

// ComposedQueryable.queryWithDistance (ShardApi.scala:72-86) for b queries at once, host buffers in and out.
int ann_sharded_query_batch(ann_sharded_index* sx, const float* queries, int32_t b, int32_t dim, int32_t k, int64_t* out_ids,
                            float* out_dist, int32_t* out_count) {
    if (!sx) return report_error(ANN_ERR_NULL_POINTER, "ann_sharded_query_batch: index is NULL");
    if (b < 0) return report_error(ANN_ERR_INVALID_ARGUMENT, "ann_sharded_query_batch: b < 0");
    if (k < 0) return report_error(ANN_ERR_NEGATIVE_K, "ann_sharded_query_batch: k < 0");
    if (dim != sx->cfg.dim) return report_error(ANN_ERR_DIMENSION_MISMATCH, "ann_sharded_query_batch: query dimension != index dimension");
    if (b == 0) return ANN_OK;
    if (!queries || (k > 0 && (!out_ids || !out_dist))) return report_error(ANN_ERR_NULL_POINTER, "ann_sharded_query_batch: NULL buffer");
    std::lock_guard<std::mutex> lk(sx->mu);
    const int R = sx->world;
    sx->queries += b;
    if (R == 1) return ann_query_batch(sx->sh[0].ix, queries, b, dim, k, out_ids, out_dist, out_count);
    if (k == 0) {
        if (out_count) std::fill(out_count, out_count + b, 0);
        return ANN_OK;
    }
    if (!sx->peer_ok) return report_error(ANN_ERR_CUDA, "ann_sharded_query_batch: the devices cannot map each other's memory (no P2P)");

    std::vector<int> rcs(R, ANN_OK);
    std::vector<std::string> msgs(R);
    std::atomic<int> failed{0}, flagged{0};
    std::vector<int32_t> cnt_host((size_t)b);
    int32_t* cnt_out = out_count ? out_count : cnt_host.data();
    std::vector<uint32_t*> recv_seed(R), recv_kth(R);   // every shard's receive buffers (peer-visible)
    std::vector<const void*> local_ptrs(R);
    std::vector<float*> q_ptrs(R);                       // every shard's query batch buffer
    // The batch crosses PCIe ONCE: every shard copies its slice of the queries from the caller's buffer and pushes it into
    // the peers' batch buffers over NVLink (the caller's memory is usually pageable -- a JVM direct buffer -- and R staged
    // copies of the whole batch cost R times the host-side staging).  Needs 16-byte aligned slices; else every shard copies
    // the whole batch.
    bool gather_q = R > 1 && sx->peer_ok;
    for (int t = 0; t < R && gather_q; ++t) {
        int t0, tn;
        slice_of(t, R, b, &t0, &tn);
        if (((size_t)t0 * dim * sizeof(float)) % 16 || ((size_t)tn * dim * sizeof(float)) % 16) gather_q = false;
    }

    // exact == true: every shard answers with its own exact top-k (device_fallback on), no threshold sharing -- the route for
    // batches in which some shard's bounded selector flagged a query
    auto run = [&](bool exact) {
        auto worker = [&](int s) {
            Shard& sh = sx->sh[s];
            int q0, qn;
            slice_of(s, R, b, &q0, &qn);
            auto step = [&](int rc) {
                if (rc && !rcs[s]) {
                    rcs[s] = rc;
                    msgs[s] = ann_last_error();
                    failed.store(1);
                }
            };
            auto cuda_step = [&](cudaError_t e, const char* what) {
                if (e != cudaSuccess && !rcs[s]) {
                    rcs[s] = ANN_ERR_CUDA;
                    msgs[s] = std::string(what) + ": " + cudaGetErrorString(e);
                    failed.store(1);
                    (void)cudaGetLastError();
                }
            };
            step(ensure_scratch(sh, b, dim, k, qn, R));
            recv_seed[s] = sh.d_seed;
            recv_kth[s] = sh.d_kth;
            local_ptrs[s] = sh.d_local;
            q_ptrs[s] = sh.d_q;
            int64_t* l_ids = reinterpret_cast<int64_t*>(sh.d_local);
            float* l_dist = reinterpret_cast<float*>(sh.d_local + (size_t)b * k * 8);
            int32_t* l_cnt = reinterpret_cast<int32_t*>(sh.d_local + (size_t)b * k * 12);
            sx->bar->wait();   // pointer tables complete; a failed allocation is visible to everyone
            const bool ok0 = !failed.load();
            cudaStream_t st = sh.st;
            auto sync_all = [&](int phase) {   // my phase is enqueued -> everyone's is -> my stream waits for all of them
                if (ok0) cuda_step(cudaEventRecord(sh.ev[phase], st), "cudaEventRecord");
                sx->bar->wait();
                if (ok0)
                    for (int t = 0; t < R; ++t)
                        if (t != s) cuda_step(cudaStreamWaitEvent(st, sx->sh[t].ev[phase], 0), "cudaStreamWaitEvent");
            };
            if (ok0) cuda_step(cudaSetDevice(sh.device), "cudaSetDevice");
            if (gather_q) {
                if (ok0 && qn > 0) {
                    const size_t off = (size_t)q0 * dim, bytes = (size_t)qn * dim * sizeof(float);
                    cuda_step(cudaMemcpyAsync(sh.d_q + off, queries + off, bytes, cudaMemcpyHostToDevice, st), "H2D query slice");
                    std::vector<void*> dst;
                    for (int t = 0; t < R; ++t)
                        if (t != s) dst.push_back(q_ptrs[t] + off);
                    cuda_step(launch_peer_push(sh.d_q + off, dst.data(), (int)dst.size(), bytes, st), "query slice push");
                }
                sync_all(3);
            } else if (ok0) {
                cuda_step(cudaMemcpyAsync(sh.d_q, queries, (size_t)b * dim * sizeof(float), cudaMemcpyHostToDevice, st), "H2D queries");
            }
            if (exact) {
                if (ok0 && !rcs[s]) {
                    step(ann_set_option(sh.ix, "device_fallback", 1));
                    step(ann_query_batch_device(sh.ix, sh.d_q, b, dim, k, l_ids, l_dist, l_cnt, st));
                    step(ann_set_option(sh.ix, "device_fallback", 0));
                }
                sync_all(2);
            } else {
                // bounds travel by PUSH: shard s writes its [b*k] block into block s of every shard's receive buffer, and
                // every consumer reads its own buffer (local HBM) after the phase's events
                const size_t blk = (size_t)b * k;
                std::vector<uint32_t*> seed_dst(R), kth_dst(R);
                std::vector<const uint32_t*> seed_src(R), kth_src(R);
                for (int t = 0; t < R; ++t) {
                    seed_dst[t] = recv_seed[t] + (size_t)s * blk;
                    kth_dst[t] = recv_kth[t] + (size_t)s * blk;
                    seed_src[t] = sh.d_seed + (size_t)t * blk;
                    kth_src[t] = sh.d_kth + (size_t)t * blk;
                }
                const bool sliced = sx->two_round && sx->sliced_seeds;
                if (sliced) {
                    // one bound per query, written by the query's slice owner into the head of every shard's seed buffer
                    if (ok0 && !failed.load())
                        step(ann_query_seed_slice_push_device(sh.ix, sh.d_q, b, dim, k, q0, qn, R, recv_seed.data(), R, st));
                } else if (ok0 && !failed.load()) {
                    step(ann_query_seed_push_device(sh.ix, sh.d_q, b, dim, k, seed_dst.data(), R, st));
                }
                sync_all(0);
                if (sliced) {
                    if (ok0 && !failed.load())
                        step(ann_query_filter_bounds_push_device(sh.ix, sh.d_q, b, dim, k, sh.d_seed, R, kth_dst.data(), R, st));
                    sync_all(1);
                    if (ok0 && !failed.load())
                        step(ann_query_rescore_device(sh.ix, sh.d_q, b, dim, k, kth_src.data(), R, l_ids, l_dist, l_cnt, st));
                } else if (sx->two_round) {
                    if (ok0 && !failed.load())
                        step(ann_query_filter_push_device(sh.ix, sh.d_q, b, dim, k, seed_src.data(), R, kth_dst.data(), R, st));
                    sync_all(1);
                    if (ok0 && !failed.load())
                        step(ann_query_rescore_device(sh.ix, sh.d_q, b, dim, k, kth_src.data(), R, l_ids, l_dist, l_cnt, st));
                } else if (ok0 && !failed.load()) {
                    step(ann_query_finish_device(sh.ix, sh.d_q, b, dim, k, seed_src.data(), R, l_ids, l_dist, l_cnt, st));
                }
                sync_all(2);
            }
            if (ok0 && !failed.load() && qn > 0) {
                step(ann_exchange_merge_slice_device(sh.device, local_ptrs.data(), R, b, k, q0, qn, sh.d_sl_ids, sh.d_sl_dist, sh.d_sl_cnt, st));
                cuda_step(cudaMemcpyAsync(out_ids + (size_t)q0 * k, sh.d_sl_ids, (size_t)qn * k * 8, cudaMemcpyDeviceToHost, st), "D2H ids");
                cuda_step(cudaMemcpyAsync(out_dist + (size_t)q0 * k, sh.d_sl_dist, (size_t)qn * k * 4, cudaMemcpyDeviceToHost, st), "D2H dist");
                cuda_step(cudaMemcpyAsync(cnt_out + q0, sh.d_sl_cnt, (size_t)qn * 4, cudaMemcpyDeviceToHost, st), "D2H count");
            }
            if (ok0) cuda_step(cudaStreamSynchronize(st), "cudaStreamSynchronize");
            if (!failed.load())
                for (int q = q0; q < q0 + qn; ++q)
                    if (cnt_out[q] < 0) {
                        flagged.store(1);
                        break;
                    }
            sx->bar->wait();   // nobody reuses the blocks / key arrays of this batch before every peer has finished reading them
        };
        sx->pool->run(worker);
    };
    run(false);
    if (!failed.load() && flagged.load()) {
        // some shard's bounded selector gave up on a query (massive ties, NaN / zero-norm query, pool overflow): answer the
        // batch again with every shard's exact fallback switched on.  Rare, and always exact.
        flagged.store(0);
        sx->fallback_batches++;
        for (Shard& s : sx->sh) {   // clear the sticky device word the first attempt left behind
            int64_t v;
            (void)ann_get_stat(s.ix, "pending_error", &v);
        }
        run(true);
    }
    for (int s = 0; s < R; ++s)
        if (rcs[s]) return report_error(rcs[s], msgs[s].c_str());
    if (flagged.load()) return report_error(ANN_ERR_CANDIDATE_OVERFLOW, "ann_sharded_query_batch: a query stayed flagged after the exact fallback");
    return ANN_OK;
}

}  // extern "C"
