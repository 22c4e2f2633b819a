// index.cu -- the C ABI (include/b200ann.h): device-resident index storage, batched append, query dispatch.
//
// Mirrors, for one GPU shard, com.twitter.ann.brute_force.BruteForceIndex
// (ann/src/main/scala/com/twitter/ann/brute_force/BruteForceIndex.scala:26-92):
//   storage  : ConcurrentLinkedQueue[EntityEmbedding[T]]  ->  row-major fp32 matrix [n][pitch] in HBM, 128-byte
//              aligned base, int64 ids, per-row norms, optional bf16 shadow operand for the tensor-core filter
//   append   : linkedQueue.add(e) per row                 ->  one H2D copy + one norms/shadow kernel per batch
//   query    : per-query scan + PriorityQueue             ->  approximate filter (streaming scan or tcgen05 GEMM)
//              whose survivors are rescored exactly and ordered by (Float.compare(distance), id)
// There is no CPU fallback: without a usable sm_100 device every compute entry point returns an error.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <utility>
#include <vector>

#include <chrono>
#include <condition_variable>
#include <deque>

#include "../../include/b200ann.h"
#include "kernels.h"
#include "vmm.hpp"

using namespace b200ann;

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}

}  // namespace

// for the other translation units of the library (knn_join.cu): record the message ann_last_error() returns
int b200ann::report_error(int code, const char* msg) { return fail(code, msg ? msg : ""); }

namespace {

#define CUDA_TRY(expr)                                                                                         \
    do {                                                                                                       \
        cudaError_t _e = (expr);                                                                               \
        if (_e != cudaSuccess) {                                                                               \
            char _b[512];                                                                                      \
            snprintf(_b, sizeof(_b), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            (void)cudaGetLastError();                                                                          \
            return fail(_e == cudaErrorMemoryAllocation ? ANN_ERR_OUT_OF_MEMORY : ANN_ERR_CUDA, _b);           \
        }                                                                                                      \
    } while (0)

constexpr int kMaxDim = 1024;
constexpr int kMaxK = 1024;
constexpr int kScanGroup = 64;       // queries finalized together on the scan path
constexpr int kScanPoolCap = 16384;  // pool entries per query, scan path
constexpr int kGemmPoolCap = 4096;   // pool entries per query, gemm path
constexpr int kPubStride = 256;
constexpr size_t kSmallResultBytes = 256 << 10;   // host queries whose whole result fits go through the pinned landing zone

struct DeviceScalars {
    uint32_t max_norm_bits;
    uint32_t error_flags;
    uint32_t max_resid_bits;   // running max of |a - bf16(a)| over the shadow rows (float bits): the measured rounding error
    uint32_t reserved;
    unsigned long long n_special;
    unsigned long long bad_queries;
    uint32_t special_list[kSpecialCap];   // indices of the first special rows (see common.cuh)
};
constexpr size_t kScalarsHeader = offsetof(DeviceScalars, special_list);   // the part the host polls

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;  // elements
    cudaError_t ensure(size_t want) {
        if (want <= n) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
        cudaError_t e = cudaMalloc(&p, want * sizeof(T));
        if (e == cudaSuccess) n = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
};

__global__ void iota_ids_kernel(int64_t* ids, long long row0, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        ids[row0 + i] = row0 + i;
}

__global__ void collect_flags_kernel(const QueryState* qs, int b, DeviceScalars* sc) {
    uint32_t f = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < b; i += gridDim.x * blockDim.x) f |= qs[i].flags;
    f = __reduce_or_sync(0xFFFFFFFFu, f);
    if ((threadIdx.x & 31) == 0 && f) atomicOr(&sc->error_flags, f);
}

}  // namespace

struct ann_index {
    ann_config cfg{};
    int dim = 0, pitch = 0, kp = 0, metric = 0;
    bool l2_squared = false, use_shadow = true, accum_f32 = false;
    bool unit_rows = false;   // ANN_FLAG_COSINE_UNIT_ROWS: rows stored normalised, Cosine answered as InnerProduct over unit vectors
    int device = 0, sm_count = 0;
    size_t smem_optin = 0;
    cudaStream_t stream = nullptr;
    // Locking.  `mu` guards the published row count and the per-query scratch: a query holds it while it runs.  `append_mu`
    // serialises appends among themselves; an append copies its rows and runs K1 on `append_stream` WITHOUT `mu` -- it only
    // writes rows beyond the published count, and the storage below never moves (vmm.hpp) -- and takes `mu` just to publish.
    // So appends do not wait for running queries and queries do not wait for running appends (BruteForceIndex.scala:48-52,
    // 66-71: both are lock-free on a ConcurrentLinkedQueue).  In-place updates take both.
    std::mutex mu, append_mu;
    cudaStream_t append_stream = nullptr;
    DeviceScalars* h_scalars = nullptr;   // pinned host mirror of the scalars' header (read back with the append's own sync)
    // pinned landing zone for SMALL query results (guarded by `mu`): a device->host copy into the caller's pageable memory
    // blocks the host until it has landed, so the four copies of a host query (ids, distances, counts, flag word) were four
    // serial round trips; into pinned memory they are queued back to back behind the kernels and ONE synchronisation
    // covers them, then the rows are copied to the caller on the host (a single-vector result is 1.2 KB)
    unsigned char* h_res = nullptr;

    long long cap = 0, n = 0, max_rows = 0;
    // storage: virtual ranges reserved once, physical memory mapped behind the rows as they arrive (grows in place)
    VmArray vm_rows, vm_ids, vm_inv, vm_norm, vm_shadow;
    float* rows = nullptr;
    int64_t* ids = nullptr;
    float* inv_norm = nullptr;
    float* row_norm = nullptr;
    __nv_bfloat16* shadow = nullptr;
    DeviceScalars* scalars = nullptr;
    unsigned long long n_special = 0;

    // micro-batcher of the host entry point (see Coalescer below)
    struct PendingQuery;
    struct Coalescer {
        std::mutex mu;
        std::condition_variable cv;   // ONE condition for all waiting calls: a finished batch wakes its callers with one broadcast
        std::deque<PendingQuery*> pending;
        bool leader_active = false;
        int max_batch = 2048;   // queries per merged device call; 0 switches coalescing off
        int small_b = 32;       // calls with at most this many queries are coalesced
        int linger_us = 100;    // a leader that inherits the lead waits this long for the callers of the batch that just
                                // finished to come back (closed-loop clients re-issue within microseconds); a lone caller never waits
        float* pin_q = nullptr;               // pinned staging of the merged batch
        unsigned char* pin_res = nullptr;
        size_t cap_q = 0, cap_res = 0;
        long long batches = 0, merged_calls = 0;
        std::atomic<int> n_pending{0};                              // pending.size(), readable without the lock
        int last_batch_calls = 0;                                   // host calls in the merged batch that finished last
        std::chrono::steady_clock::time_point last_done{};          // ... and when it finished
    } co;

    // scratch
    DevBuf<float> q_in, q_padded, out_dist, q_unit;
    DevBuf<__nv_bfloat16> q_shadow;
    DevBuf<QueryState> qstate;
    DevBuf<entry_t> pool;
    DevBuf<uint32_t> special_rows, pub_keys;
    DevBuf<int64_t> out_ids;
    DevBuf<int32_t> out_count;
    DevBuf<unsigned char> fb_scratch;
    DevBuf<float> upd_rows;
    DevBuf<long long> upd_slots;

    // options / stats
    int path_opt = 0, gemm_min_batch = 2, gemm_cta_group = 2, gemm_epi_warps = 0;
    int gemm_mma_warps = 2;           // MMA-issuing threads per CTA pair (2: warps 1 and 3 alternate query tiles; 1 = round 1's single issuer)
    int gemm_hit_budget = 500;        // candidates a chunk of the GEMM path may add per query (sets the chunk schedule)
    int gemm_growth_pct = 0;          // chunk growth factor in percent (0 = derived from the hit budget)
    int gemm_small_select = 1;        // GEMM path: 2048 / 1024-entry selector stages (4 CTAs per SM) instead of 4096 / 2048
    long long gemm_seed_rows = 0;     // rows of the threshold-seeding launch (0 = as many as the pool has slots for)
    bool device_fallback = false;
    bool gemm_blocked_by_update = false;   // device entry point: synchronise and run the exact fallback for flagged queries
    // optional CUDA-event timing of the dominant kernel of each path, on the launching stream (bench.py roofline)
    bool timing = false;
    // label 0 = the dominant kernel of the path (scan / tensor-core filter: the roofline's `achieved`); the others make up the
    // per-step breakdown bench.py prints with "timing" = 2
    enum { kLblMain = 0, kLblPrep, kLblCompact, kLblSeedMerge, kLblFinalize, kLblOther, kLblCount };
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pairs;
    std::vector<int> ev_label;
    size_t ev_used = 0;
    int timing_level = 0;     // 1 = label 0 only, 2 = every kernel of the query path
    double kernel_ms_total = 0.0;
    long long kernel_launches_timed = 0;
    double ms_by_label[kLblCount] = {0, 0, 0, 0, 0, 0};
    long long last_candidates = 0;
    std::atomic<long long> launches{0};
    long long last_path = 0, exact_fallback_queries = 0;
    long long last_gemm_chunks = 0;   // chunk launches (seed excluded) of the last tensor-core query

    // two-phase sharded query (ann_query_seed_device -> ann_query_finish_device): what the first phase left behind
    struct SeedSession {
        bool open = false;        // a seed call is waiting for its finish call
        bool gemm = false;        // the batch is on the tensor-core path (prep done, scratch in place)
        bool seeded = false;      // ... and a seed launch ran: this shard's bounds are published
        bool clobbered = false;   // another call touched the scratch or the rows in between
        bool filtered = false;    // ann_query_filter_device ran: only ann_query_rescore_device may follow
        bool sliced = false;      // sliced seeding: this shard seeded only its slice of the batch, over `scale` x the rows
        int b = 0, k = 0;
        long long seed_rows = 0;
    } sess;
};

namespace {

int set_device(const ann_index* ix) {
    CUDA_TRY(cudaSetDevice(ix->device));
    return ANN_OK;
}

// Map enough physical memory behind every array for `need` rows.  Nothing is copied and no pointer changes; fresh chunks
// are zeroed on `st` (pad columns of `rows` and tile tails of every array must read as 0).
int grow(ann_index* ix, long long need, cudaStream_t st) {
    if (need <= ix->cap) return ANN_OK;
    if (need > ix->max_rows) {
        char msg[200];
        snprintf(msg, sizeof(msg), "the index is limited to %lld rows on this device (address range reserved at creation)", ix->max_rows);
        return fail(ANN_ERR_OUT_OF_MEMORY, msg);
    }
    const long long target = (need + 127) / 128 * 128;
    struct Arr {
        VmArray* vm;
        size_t row_bytes;
    } arrs[5] = {{&ix->vm_rows, (size_t)ix->pitch * sizeof(float)},
                 {&ix->vm_ids, sizeof(int64_t)},
                 {&ix->vm_norm, sizeof(float)},
                 {ix->metric == kMetricCosine ? &ix->vm_inv : nullptr, sizeof(float)},
                 {ix->use_shadow ? &ix->vm_shadow : nullptr, (size_t)ix->kp * sizeof(__nv_bfloat16)}};
    long long ncap = ix->max_rows;
    for (const Arr& a : arrs) {
        if (!a.vm) continue;
        size_t off = 0, len = 0;
        cudaError_t e = a.vm->ensure((size_t)target * a.row_bytes, &off, &len);
        if (e == cudaSuccess && len) e = cudaMemsetAsync(static_cast<char*>(a.vm->ptr()) + off, 0, len, st);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            char msg[256];
            snprintf(msg, sizeof(msg), "growing the index to %lld rows failed: %s", target,
                     e == cudaErrorMemoryAllocation ? "out of device memory" : cudaGetErrorString(e));
            return fail(e == cudaErrorMemoryAllocation ? ANN_ERR_OUT_OF_MEMORY : ANN_ERR_CUDA, msg);
        }
        ncap = std::min<long long>(ncap, (long long)(a.vm->mapped / a.row_bytes));
    }
    ix->cap = ncap / 128 * 128;
    return ANN_OK;
}

// rows/ids already on the device (or staged there); place them and run K1
// Caller holds append_mu (not mu): only rows beyond the published count are written.
int append_device_core(ann_index* ix, const int64_t* d_ids, const float* d_rows, long long n_new, cudaStream_t st) {
    int rc = grow(ix, ix->n + n_new, st);
    if (rc) return rc;
    float* dst = ix->rows + (size_t)ix->n * ix->pitch;
    if (ix->pitch == ix->dim) {
        CUDA_TRY(cudaMemcpyAsync(dst, d_rows, (size_t)n_new * ix->dim * sizeof(float), cudaMemcpyDefault, st));
    } else {
        CUDA_TRY(cudaMemcpy2DAsync(dst, (size_t)ix->pitch * sizeof(float), d_rows, (size_t)ix->dim * sizeof(float),
                                   (size_t)ix->dim * sizeof(float), (size_t)n_new, cudaMemcpyDefault, st));
    }
    if (d_ids) {
        CUDA_TRY(cudaMemcpyAsync(ix->ids + ix->n, d_ids, (size_t)n_new * sizeof(int64_t), cudaMemcpyDefault, st));
    } else {
        int blocks = (int)std::min<long long>((n_new + 255) / 256, 1024);
        iota_ids_kernel<<<blocks, 256, 0, st>>>(ix->ids, ix->n, n_new);
        CUDA_TRY(cudaGetLastError());
        ix->launches++;
    }
    if (ix->unit_rows) {   // stored form = MetricUtil.norm(row) (convention C8); K1 below then sees unit rows
        CUDA_TRY(launch_normalize_rows_device(ix->rows, ix->n, n_new, ix->dim, ix->pitch, nullptr, st));
        ix->launches++;
    }
    AppendParams ap{};
    ap.rows = ix->rows;
    ap.row0 = ix->n;
    ap.n_new = n_new;
    ap.dim = ix->dim;
    ap.pitch = ix->pitch;
    ap.metric = ix->metric;
    ap.inv_norm = ix->inv_norm;
    ap.row_norm = ix->row_norm;
    ap.shadow = ix->shadow;
    ap.kp = ix->kp;
    ap.max_norm_bits = &ix->scalars->max_norm_bits;
    ap.max_resid_bits = &ix->scalars->max_resid_bits;
    ap.n_special = &ix->scalars->n_special;
    ap.special_list = ix->scalars->special_list;
    CUDA_TRY(launch_append(ap, st));
    ix->launches++;
    return ANN_OK;
}


// ---- optional per-launch CUDA-event timing (enabled with ann_set_option("timing", 1)) ----
struct TimedScope {
    ann_index* ix;
    cudaStream_t st;
    std::pair<cudaEvent_t, cudaEvent_t>* pr = nullptr;
    TimedScope(ann_index* i, cudaStream_t s, int label = ann_index::kLblMain) : ix(i), st(s) {
        if (!ix->timing || (label != ann_index::kLblMain && ix->timing_level < 2)) return;
        if (ix->ev_used == ix->ev_pairs.size()) {
            cudaEvent_t a, b;
            if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
            ix->ev_pairs.emplace_back(a, b);
            ix->ev_label.push_back(0);
        }
        ix->ev_label[ix->ev_used] = label;
        pr = &ix->ev_pairs[ix->ev_used++];
        cudaEventRecord(pr->first, st);
    }
    ~TimedScope() {
        if (pr) cudaEventRecord(pr->second, st);
    }
};

// fold finished event pairs into the totals (call after the stream has been synchronised)
void harvest_timing(ann_index* ix) {
    for (size_t i = 0; i < ix->ev_used; ++i) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ix->ev_pairs[i].first, ix->ev_pairs[i].second) == cudaSuccess) {
            ix->ms_by_label[ix->ev_label[i]] += ms;
            if (ix->ev_label[i] == ann_index::kLblMain) {
                ix->kernel_ms_total += ms;
                ix->kernel_launches_timed++;
            }
        }
    }
    ix->ev_used = 0;
    (void)cudaGetLastError();
}

struct ScanPlan {
    int qb, warps, cap, grid, r_pub, j_pub;
    size_t smem;
};

int plan_scan(const ann_index* ix, int b, int k_eff, ScanPlan* pl) {
    int cap = 1024;
    while (cap < 4 * k_eff) cap <<= 1;
    int qb = b >= 8 ? 8 : (b > 2 ? 4 : b);
    while (qb > 1 && (size_t)qb * cap * 8 > 64 * 1024) qb >>= 1;
    const size_t tile_bytes = (size_t)32 * ix->pitch * sizeof(float);
    int warps = 0;
    for (;;) {
        size_t fixed = scan_smem_bytes(qb, ix->pitch, 0, cap);
        if (fixed + tile_bytes <= ix->smem_optin) {
            warps = (int)std::min<size_t>(8, (ix->smem_optin - fixed) / tile_bytes);
            break;
        }
        if (qb == 1) break;
        qb >>= 1;
    }
    if (warps < 1) return fail(ANN_ERR_INVALID_ARGUMENT, "dimension too large for the streaming scan");
    long long n_tiles = (ix->n + 31) / 32;
    int grid = (int)std::min<long long>(ix->sm_count, std::max<long long>(1, (n_tiles + warps - 1) / warps));
    grid = std::min(grid, kPubStride);
    pl->qb = qb;
    pl->warps = warps;
    pl->cap = cap;
    pl->grid = grid;
    pl->j_pub = std::max(1, std::min(k_eff, (3 * grid) / 4));
    pl->r_pub = (k_eff + pl->j_pub - 1) / pl->j_pub;
    pl->smem = scan_smem_bytes(qb, ix->pitch, warps, cap);
    return ANN_OK;
}

int query_scan(ann_index* ix, const float* d_queries, int b, int k_eff, int k_out, int64_t* d_out_ids, float* d_out_dist,
               int32_t* d_out_count, cudaStream_t st) {
    ScanPlan pl;
    int rc = plan_scan(ix, b, k_eff, &pl);
    if (rc) return rc;
    CUDA_TRY(ix->q_padded.ensure((size_t)b * ix->pitch));
    CUDA_TRY(ix->qstate.ensure((size_t)b));
    CUDA_TRY(ix->pub_keys.ensure((size_t)b * kPubStride));
    const int group = std::min(b, kScanGroup);
    CUDA_TRY(ix->pool.ensure((size_t)group * kScanPoolCap));
    CUDA_TRY(ix->special_rows.ensure((size_t)std::max(group, 1) * kSpecialCap));

    PrepParams pp{};
    pp.queries = d_queries;
    pp.b = b;
    pp.dim = ix->dim;
    pp.pitch = ix->pitch;
    pp.metric = ix->metric;
    pp.kp = ix->kp;
    pp.q_padded = ix->q_padded.p;
    pp.q_shadow = nullptr;
    pp.qstate = ix->qstate.p;
    pp.max_norm_bits = &ix->scalars->max_norm_bits;
    pp.max_resid_bits = &ix->scalars->max_resid_bits;
    pp.path = 1;
    pp.accum_f32 = ix->accum_f32 ? 1 : 0;
    pp.pub_keys = ix->pub_keys.p;
    pp.pub_stride = kPubStride;
    pp.bad_queries = &ix->scalars->bad_queries;
    { TimedScope ts_(ix, st, ann_index::kLblPrep); CUDA_TRY(launch_prep_queries(pp, st)); }
    ix->launches++;

    for (int g0 = 0; g0 < b; g0 += group) {
        const int gn = std::min(group, b - g0);
        for (int q0 = 0; q0 < gn; q0 += pl.qb) {
            ScanParams sp{};
            sp.rows = ix->rows;
            sp.n_rows = ix->n;
            sp.pitch = ix->pitch;
            sp.metric = ix->metric;
            sp.inv_norm = ix->inv_norm;
            sp.queries = ix->q_padded.p + (size_t)(g0 + q0) * ix->pitch;
            sp.nq = std::min(pl.qb, gn - q0);
            sp.qstate = ix->qstate.p + g0 + q0;
            sp.pool = ix->pool.p + (size_t)q0 * kScanPoolCap;
            sp.pool_cap = kScanPoolCap;
            sp.special_rows = ix->special_rows.p + (size_t)q0 * kSpecialCap;
            sp.pub_keys = ix->pub_keys.p + (size_t)(g0 + q0) * kPubStride;
            sp.pub_stride = kPubStride;
            sp.k = k_eff;
            sp.r_pub = pl.r_pub;
            sp.j_pub = pl.j_pub;
            sp.warps = pl.warps;
            sp.cap = pl.cap;
            {
                TimedScope ts(ix, st);
                CUDA_TRY(launch_scan(sp, pl.qb, pl.grid, pl.smem, st));
            }
            ix->launches++;
        }
        SelectParams fp{};
        fp.qstate = ix->qstate.p + g0;
        fp.pool = ix->pool.p;
        fp.pool_cap = kScanPoolCap;
        fp.special_rows = ix->special_rows.p;
        fp.pub_keys = ix->pub_keys.p + (size_t)g0 * kPubStride;
        fp.pub_stride = kPubStride;
        fp.pub_count = pl.grid;
        fp.j_pub = pl.j_pub;
        fp.k = k_eff;
        fp.rows = ix->rows;
        fp.ids = ix->ids;
        fp.row_norm = ix->row_norm;
        fp.n_rows = ix->n;
        fp.dim = ix->dim;
        fp.pitch = ix->pitch;
        fp.metric = ix->metric;
        fp.l2_squared = ix->l2_squared ? 1 : 0;
        fp.accum_f32 = ix->accum_f32 ? 1 : 0;
        fp.queries = ix->q_padded.p + (size_t)g0 * ix->pitch;
        fp.q_pitch = ix->pitch;
        fp.out_ids = d_out_ids + (size_t)g0 * k_out;
        fp.out_dist = d_out_dist + (size_t)g0 * k_out;
        fp.out_count = d_out_count ? d_out_count + g0 : nullptr;
        fp.k_out = k_out;
        { TimedScope ts_(ix, st, ann_index::kLblFinalize); CUDA_TRY(launch_finalize(fp, gn, st)); }
        ix->launches++;
    }
    collect_flags_kernel<<<std::min(64, (b + 255) / 256), 256, 0, st>>>(ix->qstate.p, b, ix->scalars);
    CUDA_TRY(cudaGetLastError());
    ix->launches++;
    ix->last_path = 1;
    return ANN_OK;
}


// K3 flow: geometric row chunks, each scored by the tensor-core filter against the thresholds learnt from the rows
// before it; an approximate compaction between chunks; one exact finalize at the end.
// mode 0: the whole flow.  mode 1 / 2: the two halves of a sharded query -- 1 = prepare + seed launch + publish this
// shard's bounds into `seed_keys_out`; 2 = take the global threshold from every shard's bounds (`peers`), then chunks +
// finalize on the scratch mode 1 left in place.
// mode 3 / 4 split mode 2 once more around a SECOND cross-shard round: 3 = thresholds + chunks + last compaction, which
// publishes this shard's k best approximate bounds into `seed_keys_out`; 4 = exact finalize against the k-th best bound of
// all shards (`peers`), so that every shard rescores only its share of the global survivors.
// Sliced seeding (`slice` != nullptr): in mode 1 the shard seeds only the queries [q_begin, q_begin + q_count) but over
// `scale` times the rows, and publishes ONE bound per query (the k-th best group maximum as an upper bound) into every
// peer's [b] array; in mode 3 the thresholds of all b queries come from the array the slice owners filled (`bounds`).
// Same GEMM work per shard as seeding every query over a short prefix, the same rows seen per query as the union of all
// shards' prefixes -- but k times less to exchange and no k-th-of-the-union selection on the consumer side.
struct SliceSeed {
    int q_begin = 0, q_count = 0, scale = 1;
    const uint32_t* bounds = nullptr;
};

int query_gemm(ann_index* ix, QueryState* qs_base, const float* d_queries, int b, int k_eff, int k_out, int64_t* d_out_ids,
               float* d_out_dist, int32_t* d_out_count, cudaStream_t st, int mode = 0, uint32_t* seed_keys_out = nullptr,
               const PeerSeedKeys* peers = nullptr, int world = 1, uint32_t* const* push_dst = nullptr, int n_push = 0,
               const SliceSeed* slice = nullptr) {
    const int b_pad = (b + 127) / 128 * 128;
    CUDA_TRY(ix->q_padded.ensure((size_t)b * ix->pitch));
    const int qkp = (ix->kp + 63) / 64 * 64;
    CUDA_TRY(ix->q_shadow.ensure((size_t)b_pad * qkp));
    CUDA_TRY(ix->pool.ensure((size_t)b * kGemmPoolCap));
    CUDA_TRY(ix->special_rows.ensure((size_t)kSpecialCap));

    PrepParams pp{};
    pp.queries = d_queries;
    pp.b = b;
    pp.dim = ix->dim;
    pp.pitch = ix->pitch;
    pp.metric = ix->metric;
    pp.kp = ix->kp;
    pp.q_padded = ix->q_padded.p;
    pp.q_shadow = ix->q_shadow.p;
    pp.qkp = qkp;
    pp.qstate = qs_base;
    pp.max_norm_bits = &ix->scalars->max_norm_bits;
    pp.max_resid_bits = &ix->scalars->max_resid_bits;
    pp.path = 2;
    pp.accum_f32 = ix->accum_f32 ? 1 : 0;
    pp.pub_keys = nullptr;
    pp.pub_stride = 0;
    pp.bad_queries = &ix->scalars->bad_queries;
    if (mode < 2) {
        { TimedScope ts_(ix, st, ann_index::kLblPrep); CUDA_TRY(launch_prep_queries(pp, st)); }
        ix->launches++;
    }

    SelectParams fp{};
    if (ix->gemm_small_select) {   // pools on this path hold a few hundred entries; overflow still falls back exactly
        fp.sort_cap = 2048;
        fp.exact_cap = 1024;
    }
    fp.qstate = qs_base;
    fp.pool = ix->pool.p;
    fp.pool_cap = kGemmPoolCap;
    fp.special_rows = ix->special_rows.p;
    fp.global_special_rows = ix->scalars->special_list;
    fp.global_special_count = (int)std::min<unsigned long long>(ix->n_special, (unsigned long long)kSpecialCap);
    fp.pub_keys = nullptr;
    fp.k = k_eff;
    fp.rows = ix->rows;
    fp.ids = ix->ids;
    fp.row_norm = ix->row_norm;
    fp.n_rows = ix->n;
    fp.dim = ix->dim;
    fp.pitch = ix->pitch;
    fp.metric = ix->metric;
    fp.l2_squared = ix->l2_squared ? 1 : 0;
    fp.accum_f32 = ix->accum_f32 ? 1 : 0;
    fp.queries = ix->q_padded.p;
    fp.q_pitch = ix->pitch;
    fp.out_ids = d_out_ids;
    fp.out_dist = d_out_dist;
    fp.out_count = d_out_count;
    fp.k_out = k_out;

    if (mode == 4) {
        if (peers && world > 1) {
            fp.peer_keys = *peers;
            fp.peer_world = world;
        }
        { TimedScope ts_(ix, st, ann_index::kLblFinalize); CUDA_TRY(launch_finalize(fp, b, st)); }
        ix->launches++;
        collect_flags_kernel<<<std::min(64, (b + 255) / 256), 256, 0, st>>>(qs_base, b, ix->scalars);
        CUDA_TRY(cudaGetLastError());
        ix->launches++;
        ix->last_path = 2;
        return ANN_OK;
    }
    // where modes 1 / 3 publish: one local array, or (push) this shard's block inside every peer's receive buffer
    auto set_publish = [&](SelectParams& sp) {
        sp.seed_keys_out = seed_keys_out;
        sp.n_push = 0;
        for (int i = 0; i < n_push && i < kMaxPeers; ++i) sp.push_keys[sp.n_push++] = push_dst[i];
    };
    auto publish_no_bound = [&]() -> int {
        const size_t bytes = (size_t)b * k_out * sizeof(uint32_t);
        if (n_push > 0) {
            for (int i = 0; i < n_push; ++i) CUDA_TRY(cudaMemsetAsync(push_dst[i], 0xFF, bytes, st));
        } else if (seed_keys_out) {
            CUDA_TRY(cudaMemsetAsync(seed_keys_out, 0xFF, bytes, st));
        }
        return ANN_OK;
    };
    const int kHitBudget = ix->gemm_hit_budget;
    auto gemm_launch = [&](long long begin, long long end, int seed_mode, int q_first = 0, int q_n = -1) -> int {
        if (q_n < 0) q_n = b;
        GemmLaunch g{};
        g.q_shadow = ix->q_shadow.p + (size_t)q_first * qkp;   // a query sub-range: the operand rows behind it stay addressable
        g.qkp = qkp;
        g.shadow = ix->shadow;
        g.n_rows_total = ix->n;
        g.row_begin = begin;
        g.row_end = end;
        g.b = q_n;
        g.b_pad = b_pad - q_first;
        g.kp = ix->kp;
        g.cta_group = ix->gemm_cta_group;
        g.seed_mode = seed_mode;
        // a chunk is sized to add about kHitBudget candidates per query, so its hit density is budget / rows; above
        // ~1.5e-4 hits per score the epilogue is the bottleneck and wants 16 warps (same-process A/B, tools/ab_options.py:
        // 2.82 -> 2.52 ms per 4096-query batch over 1.25M rows; neutral at 10M rows where sparse chunks dominate)
        // With a single resident query tile (b <= 256) the launch is HBM bound and 8 warps measured 8 % faster.
        g.epi_warps = seed_mode ? 8   // the seed group is the column range of one of 8 epilogue warps (gemm_seed_group_rows)
                      : ix->gemm_epi_warps ? ix->gemm_epi_warps
                                         : ((b > 256 && (double)(end - begin) * 1.5e-4 < (double)kHitBudget) ? 16 : 8);
        g.mma_warps = ix->gemm_mma_warps;
        g.nb_stages = gemm_row_stages(ix->kp, ix->smem_optin);
        g.sm_count = ix->sm_count;
        g.qstate = qs_base + q_first;
        g.pool = ix->pool.p + (size_t)q_first * kGemmPoolCap;
        g.pool_cap = kGemmPoolCap;
        cudaEvent_t dbg0 = nullptr, dbg1 = nullptr;
        const bool dbg = getenv("B200ANN_DEBUG") != nullptr;
        if (dbg) {
            cudaEventCreate(&dbg0);
            cudaEventCreate(&dbg1);
            cudaEventRecord(dbg0, st);
        }
        {
            TimedScope ts(ix, st);
            CUDA_TRY(launch_gemm_filter(g, st));
        }
        ix->launches++;
        if (dbg) {
            std::vector<QueryState> h(b);
            cudaEventRecord(dbg1, st);
            CUDA_TRY(cudaStreamSynchronize(st));
            float dms = 0.f;
            cudaEventElapsedTime(&dms, dbg0, dbg1);
            cudaEventDestroy(dbg0);
            cudaEventDestroy(dbg1);
            CUDA_TRY(cudaMemcpy(h.data(), qs_base, sizeof(QueryState) * b, cudaMemcpyDeviceToHost));
            uint32_t mx = 0, mn = ~0u;
            double sum = 0;
            for (auto& x : h) {
                mx = std::max(mx, x.pool_count);
                mn = std::min(mn, x.pool_count);
                sum += x.pool_count;
            }
            fprintf(stderr, "[b200ann] gemm%s [%lld,%lld): %.3f ms %.1f TFLOP/s; pool_count min %u mean %.1f max %u tau[0]=%g\n",
                    seed_mode ? " seed" : "", begin, end, dms, 2.0 * (double)(end - begin) * ix->dim * q_n / (dms * 1e-3) / 1e12, mn,
                    sum / b, mx, float_from_order_key(h[0].tau_key));
        }
        return ANN_OK;
    };

    // Chunk schedule.  Thresholds only tighten between launches, so the work is cut into geometrically growing row
    // ranges; each launch filters against the k-th best seen so far (+ margin) and a compaction follows it.
    //   seed   : over the first S rows only the best score of every 32-row group is kept (fixed slots, no atomics); the
    //            k-th best group maximum is a valid threshold, as tight as having seen ~S rows.  Needs S/32 >= 4k groups.
    //   chunks : then [0, c1), [c1, c2), ... sized so that each adds roughly kHitBudget candidates per query.
    // Seeding is an extra pass over its rows (the first chunk scores them again), so it is kept short: 65536 rows (2048
    // group maxima, >= 8 groups per neighbour up to k = 256) measured 1.6 % faster than 131072 on the 10M x 200 batch.
    // One pool entry per group of G rows (128 with CTA pairs): 65536 rows give 512 group maxima, 5 per neighbour at k = 100
    // (the k-th best of them sits where the 111th best row does); larger k seeds over 4 groups per neighbour.
    const long long G = gemm_seed_group_rows(ix->gemm_cta_group);
    const long long seed_scale = (mode == 1 && slice) ? std::max(1, slice->scale) : 1;
    long long seed_rows = std::min<long long>(ix->n / 256 * 256, std::min<long long>((long long)kGemmPoolCap * G, std::max<long long>(65536, 4 * G * k_eff) * seed_scale));
    if (ix->gemm_seed_rows > 0) seed_rows = std::min<long long>(seed_rows, ix->gemm_seed_rows * seed_scale / 256 * 256);
    bool use_seed = seed_rows >= 2 * G * k_eff && seed_rows >= 4096;
    bool sliced = mode == 1 && slice != nullptr;
    if (mode >= 2) {   // what phase 1 decided and did
        use_seed = ix->sess.seeded;
        seed_rows = ix->sess.seed_rows;
        sliced = ix->sess.sliced;
    }
    // a chunk `growth` times the rows seen so far adds about (growth - 1) * 1.9 * k candidates per query
    double growth = std::floor(std::min<double>(8.0, std::max<double>(2.0, 1.0 + kHitBudget / (1.9 * std::max(1, k_eff)))));
    if (ix->gemm_growth_pct > 0) growth = ix->gemm_growth_pct / 100.0;
    long long begin = 0, end;
    ix->last_gemm_chunks = 0;
    if (use_seed) {
        double seen = (double)seed_rows;   // rows the threshold has been learnt from
        if (mode < 2 && sliced) {
            if (slice->q_count > 0) {
                int rc2 = gemm_launch(0, seed_rows, 1, slice->q_begin, slice->q_count);
                if (rc2) return rc2;
                SelectParams sp = fp;
                sp.qstate = qs_base + slice->q_begin;
                sp.pool = ix->pool.p + (size_t)slice->q_begin * kGemmPoolCap;
                sp.seed_count = (int)(seed_rows / G);
                sp.sort_cap = std::max(sp.sort_cap, sp.seed_count);
                sp.pub_single = 1;
                sp.seed_keys_out = seed_keys_out ? seed_keys_out + slice->q_begin : nullptr;
                sp.n_push = 0;
                for (int i = 0; i < n_push && i < kMaxPeers; ++i) sp.push_keys[sp.n_push++] = push_dst[i] + slice->q_begin;
                { TimedScope ts_(ix, st, ann_index::kLblCompact); CUDA_TRY(launch_compact_pool(sp, slice->q_count, st)); }
                ix->launches++;
            }
        } else if (mode < 2) {
            int rc2 = gemm_launch(0, seed_rows, 1);
            if (rc2) return rc2;
            SelectParams sp = fp;
            sp.seed_count = (int)(seed_rows / G);
            sp.sort_cap = std::max(sp.sort_cap, sp.seed_count);   // every seed entry is loaded
            if (mode == 1) set_publish(sp);
            { TimedScope ts_(ix, st, ann_index::kLblCompact); CUDA_TRY(launch_compact_pool(sp, b, st)); }
            ix->launches++;
        }
        if (mode == 1) {
            ix->sess.seeded = true;
            ix->sess.seed_rows = seed_rows;
            ix->sess.sliced = sliced;
            return ANN_OK;
        }
        if (mode >= 2 && sliced) {
            // every query's bound was learnt by its slice owner from seed_rows rows of THAT shard: as tight as this shard's
            // own seed over as many rows, and nothing to select
            if (slice && slice->bounds) {
                { TimedScope ts_(ix, st, ann_index::kLblSeedMerge); CUDA_TRY(launch_apply_bounds(slice->bounds, qs_base, b, st)); }
                ix->launches++;
            }
        } else if (mode >= 2 && peers && world > 1) {
            // K5c: the k-th best of the union of all shards' published bounds replaces this shard's own seed threshold:
            // as tight as one seed over world * seed_rows rows, for the price of one seed launch per shard
            { TimedScope ts_(ix, st, ann_index::kLblSeedMerge); CUDA_TRY(launch_seed_merge(*peers, world, qs_base, b, k_eff, st)); }
            ix->launches++;
            seen *= world;
        }
        // a threshold learnt from S rows lets through about 2.2 * k / S of the rows (group loss 1.13 x margin ~1.9)
        end = std::min<long long>(ix->n, std::max<long long>(seed_rows, (long long)((double)kHitBudget * seen / (2.2 * k_eff))));
        end = (end + 255) / 256 * 256;
        if (mode >= 2 && ix->n - end < end / 4) end = ix->n;   // no sliver of a last chunk behind a shard-sized first one
        if (end > ix->n) end = ix->n;
    } else if (mode == 1) {
        // too few rows to seed from: nothing to publish, phase 2 runs the unseeded schedule
        if (sliced) {
            const size_t bytes = (size_t)slice->q_count * sizeof(uint32_t);
            for (int i = 0; i < n_push && bytes; ++i) CUDA_TRY(cudaMemsetAsync(push_dst[i] + slice->q_begin, 0xFF, bytes, st));
            if (n_push == 0 && seed_keys_out && bytes) CUDA_TRY(cudaMemsetAsync(seed_keys_out + slice->q_begin, 0xFF, bytes, st));
        } else {
            int rc2 = publish_no_bound();
            if (rc2) return rc2;
        }
        ix->sess.seeded = false;
        ix->sess.seed_rows = 0;
        ix->sess.sliced = sliced;
        return ANN_OK;
    } else {
        // (mode 3 after an unseeded phase 1: bounds other slice owners delivered still help -- a shard too small to seed
        // is too small for them to matter, so they are simply not applied)
        // no seed: the first chunk is scored with tau = +inf (every row is a candidate), so it must fit the pool
        end = std::min<long long>(ix->n, kGemmPoolCap / 2);
    }
    for (;;) {
        int rc2 = gemm_launch(begin, end, 0);
        if (rc2) return rc2;
        ix->last_gemm_chunks++;
        if (end >= ix->n) break;
        { TimedScope ts_(ix, st, ann_index::kLblCompact); CUDA_TRY(launch_compact_pool(fp, b, st)); }
        ix->launches++;
        begin = end;
        end = std::min<long long>(ix->n, ((long long)((double)end * growth) + 255) / 256 * 256);
    }
    if (mode == 3) {   // last compaction + publish; the exact finalize follows the second cross-shard round (mode 4)
        SelectParams sp = fp;
        set_publish(sp);
        { TimedScope ts_(ix, st, ann_index::kLblCompact); CUDA_TRY(launch_compact_pool(sp, b, st)); }
        ix->launches++;
        return ANN_OK;
    }
    { TimedScope ts_(ix, st, ann_index::kLblFinalize); CUDA_TRY(launch_finalize(fp, b, st)); }
    ix->launches++;
    collect_flags_kernel<<<std::min(64, (b + 255) / 256), 256, 0, st>>>(qs_base, b, ix->scalars);
    CUDA_TRY(cudaGetLastError());
    ix->launches++;
    ix->last_path = 2;
    return ANN_OK;
}

bool gemm_eligible(const ann_index* ix, int b, int k_eff) {
    return ix->shadow != nullptr && ix->n_special <= (unsigned long long)kSpecialCap && !ix->gemm_blocked_by_update && k_eff <= 256 &&
           ix->n >= 1024 && b >= 1 &&
           gemm_row_stages(ix->kp, ix->smem_optin) > 0;
}

// ANN_FLAG_COSINE_UNIT_ROWS: the query is normalised too (MetricUtil.norm, C8) and the index answers 1 - q^.a^ as InnerProduct
int unit_queries(ann_index* ix, const float** d_queries, int b, cudaStream_t st) {
    if (!ix->unit_rows || b == 0 || !*d_queries) return ANN_OK;
    CUDA_TRY(ix->q_unit.ensure((size_t)b * ix->dim));
    CUDA_TRY(launch_normalize_rows_device(const_cast<float*>(*d_queries), 0, b, ix->dim, ix->dim, ix->q_unit.p, st));
    ix->launches++;
    *d_queries = ix->q_unit.p;
    return ANN_OK;
}

int query_core(ann_index* ix, const float* d_queries, int b, int k, int64_t* d_out_ids, float* d_out_dist,
               int32_t* d_out_count, cudaStream_t st) {
    if (b == 0) return ANN_OK;
    const long long n = ix->n;
    if (k == 0 || n == 0) {
        if (k > 0 || d_out_count) {
            CUDA_TRY(launch_fill_empty(d_out_ids, d_out_dist, d_out_count, b, k, st));
            ix->launches++;
        }
        return ANN_OK;
    }
    const int k_eff = (int)std::min<long long>(k, n);
    if (k_eff > 16384) return fail(ANN_ERR_INVALID_ARGUMENT, "min(k, size) > 16384 is not supported");
    if (k_eff > kMaxK || ix->path_opt == 3) {
        // Beyond the bounded selectors' range (or on request, path = 3): the exact fallback for every query -- exact distance
        // for every row + radix select on (distance, id).  Always correct, ~n*dim*4 + 144*n bytes of HBM traffic per query.
        CUDA_TRY(ix->fb_scratch.ensure(fallback_scratch_bytes(n, k_eff)));
        for (int q = 0; q < b; ++q) {
            FallbackParams fp{};
            fp.rows = ix->rows;
            fp.ids = ix->ids;
            fp.n_rows = n;
            fp.pitch = ix->pitch;
            fp.dim = ix->dim;
            fp.metric = ix->metric;
            fp.l2_squared = ix->l2_squared ? 1 : 0;
            fp.accum_f32 = ix->accum_f32 ? 1 : 0;
            fp.query = d_queries + (size_t)q * ix->dim;
            fp.scratch = ix->fb_scratch.p;
            fp.k = k_eff;
            fp.k_out = k;
            fp.out_ids = d_out_ids + (size_t)q * k;
            fp.out_dist = d_out_dist + (size_t)q * k;
            fp.out_count = d_out_count ? d_out_count + q : nullptr;
            {
                long long fl = 0;
                CUDA_TRY(launch_exact_fallback(fp, st, &fl));
                ix->launches += fl;
            }
        }
        ix->last_path = 3;
        return ANN_OK;
    }
    int path = ix->path_opt;
    if (path == 2 && !gemm_eligible(ix, b, k_eff))
        return fail(ANN_ERR_INVALID_ARGUMENT,
                    "path=2 (tensor-core filter) needs the bf16 shadow, at most 256 non-finite/zero-norm rows, size >= 1024, k <= 256 and dim <= ~640");
    // profiles/r01_crossover.txt: with the resident-query pipeline the tensor-core path is at least as fast as the scan for
    // every batch >= 2 at every shard size measured (100K .. 10M rows).  Single queries stay on the HBM-streaming scan by
    // default (the fp32 matrix itself is scanned, nothing approximate is involved); gemm_min_batch = 1 routes them to the
    // tensor-core path too, which reads the half-size bf16 shadow (0.85 vs 1.22 ms at 10M x 200).
    if (path == 0) path = (gemm_eligible(ix, b, k_eff) && b >= ix->gemm_min_batch) ? 2 : 1;
    if (path == 2) {
        // bound the per-call scratch (pool = 32 KB per query): very large batches go through in slices
        constexpr int kMaxGemmBatch = 16384;
        CUDA_TRY(ix->qstate.ensure((size_t)b));   // per-query state of the WHOLE batch stays addressable for the fallback
        for (int q0 = 0; q0 < b; q0 += kMaxGemmBatch) {
            const int nb = std::min(kMaxGemmBatch, b - q0);
            int rc = query_gemm(ix, ix->qstate.p + q0, d_queries + (size_t)q0 * ix->dim, nb, k_eff, k, d_out_ids + (size_t)q0 * k,
                                d_out_dist + (size_t)q0 * k, d_out_count ? d_out_count + q0 : nullptr, st);
            if (rc) return rc;
        }
        return ANN_OK;
    }
    return query_scan(ix, d_queries, b, k_eff, k, d_out_ids, d_out_dist, d_out_count, st);
}


// Queries the bounded selector flagged (QueryState.flags != 0) are answered again, exactly, by exact_fallback.cu.
// Needs the per-query flags on the host, so it synchronises `st`; used by the host-buffer entry point.
int resolve_flagged(ann_index* ix, const float* d_queries, int b, int k, int64_t* d_out_ids, float* d_out_dist,
                    int32_t* d_out_count, cudaStream_t st, const DeviceScalars* snapshot = nullptr, bool* ran = nullptr) {
    if (ran) *ran = false;
    if (b == 0 || k == 0 || ix->n == 0) return ANN_OK;
    DeviceScalars hs{};
    if (snapshot) {
        hs = *snapshot;
    } else {
        CUDA_TRY(cudaMemcpyAsync(&hs, ix->scalars, kScalarsHeader, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
    }
    if (!hs.error_flags) return ANN_OK;
    if (ran) *ran = true;
    std::vector<QueryState> h((size_t)b);
    CUDA_TRY(cudaMemcpyAsync(h.data(), ix->qstate.p, sizeof(QueryState) * (size_t)b, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    const int k_eff = (int)std::min<long long>(k, ix->n);
    CUDA_TRY(ix->fb_scratch.ensure(fallback_scratch_bytes(ix->n, k_eff)));
    for (int q = 0; q < b; ++q) {
        if (!h[q].flags) continue;
        FallbackParams fp{};
        fp.rows = ix->rows;
        fp.ids = ix->ids;
        fp.n_rows = ix->n;
        fp.pitch = ix->pitch;
        fp.dim = ix->dim;
        fp.metric = ix->metric;
        fp.l2_squared = ix->l2_squared ? 1 : 0;
        fp.accum_f32 = ix->accum_f32 ? 1 : 0;
        fp.query = d_queries + (size_t)q * ix->dim;
        fp.scratch = ix->fb_scratch.p;
        fp.k = k_eff;
        fp.k_out = k;
        fp.out_ids = d_out_ids + (size_t)q * k;
        fp.out_dist = d_out_dist + (size_t)q * k;
        fp.out_count = d_out_count ? d_out_count + q : nullptr;
        {
                long long fl = 0;
                CUDA_TRY(launch_exact_fallback(fp, st, &fl));
                ix->launches += fl;
            }
        ix->exact_fallback_queries++;
    }
    CUDA_TRY(cudaMemsetAsync(&ix->scalars->error_flags, 0, sizeof(uint32_t), st));
    return ANN_OK;
}

// read and clear the sticky device error word; call after a synchronisation point
int check_device_flags(ann_index* ix, cudaStream_t st) {
    DeviceScalars h{};
    CUDA_TRY(cudaMemcpyAsync(&h, ix->scalars, kScalarsHeader, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    ix->n_special = h.n_special;
    if (h.error_flags) {
        CUDA_TRY(cudaMemsetAsync(&ix->scalars->error_flags, 0, sizeof(uint32_t), st));
        CUDA_TRY(cudaStreamSynchronize(st));
        char msg[256];
        snprintf(msg, sizeof(msg),
                 "exact selector overflow (flags=0x%x): too many rows tie within the error margin of rank k, or too many "
                 "non-finite scores",
                 h.error_flags);
        return fail(ANN_ERR_CANDIDATE_OVERFLOW, msg);
    }
    return ANN_OK;
}

}  // namespace

// ================================================================== C ABI =====================================

extern "C" {

int ann_version(void) { return 100; }

const char* ann_last_error(void) { return g_last_error.c_str(); }

int ann_create(const ann_config* cfg, ann_index** out) {
    if (!cfg || !out) return fail(ANN_ERR_NULL_POINTER, "ann_create: cfg/out is NULL");
    *out = nullptr;
    if (cfg->metric < 0 || cfg->metric > 2) return fail(ANN_ERR_INVALID_ARGUMENT, "ann_create: unknown metric ordinal");
    if (cfg->dim < 1 || cfg->dim > kMaxDim) return fail(ANN_ERR_INVALID_ARGUMENT, "ann_create: dim must be in 1..1024");
    if (cfg->capacity_hint < 0) return fail(ANN_ERR_INVALID_ARGUMENT, "ann_create: negative capacity_hint");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        (void)cudaGetLastError();
        return fail(ANN_ERR_NO_DEVICE, "ann_create: no CUDA device (this engine has no CPU fallback)");
    }
    if (cfg->device < 0 || cfg->device >= ndev) return fail(ANN_ERR_INVALID_ARGUMENT, "ann_create: bad device ordinal");
    cudaDeviceProp prop{};
    CUDA_TRY(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10) {
        char msg[160];
        snprintf(msg, sizeof(msg), "ann_create: device %d is sm_%d%d; this library is built for sm_100a only", cfg->device,
                 prop.major, prop.minor);
        return fail(ANN_ERR_NO_DEVICE, msg);
    }
    ann_index* ix = new (std::nothrow) ann_index();
    if (!ix) return fail(ANN_ERR_OUT_OF_MEMORY, "ann_create: host allocation failed");
    ix->cfg = *cfg;
    ix->dim = cfg->dim;
    ix->metric = cfg->metric;
    ix->pitch = (cfg->dim + 3) / 4 * 4;
    ix->l2_squared = (cfg->flags & ANN_FLAG_L2_SQUARED) != 0;
    ix->accum_f32 = (cfg->flags & ANN_FLAG_ACCUM_F32) != 0;
    if (cfg->flags & ANN_FLAG_COSINE_UNIT_ROWS) {
        // Cosine the way the reference's HNSW / Faiss backends do it (DistanceFunctionGenerator.scala:11-30, Hnsw.scala:149-155,
        // QueryableIndexAdapter.scala:43-50): normalise once at append, then every scan is a plain inner product -- no norm
        // array, no per-row inv_norm read.  Internally the index IS an InnerProduct index over unit rows.
        if (cfg->metric != kMetricCosine) {
            delete ix;
            return fail(ANN_ERR_INVALID_ARGUMENT, "ann_create: ANN_FLAG_COSINE_UNIT_ROWS needs metric = Cosine");
        }
        ix->unit_rows = true;
        ix->metric = kMetricIP;
    }
    ix->use_shadow = (cfg->flags & ANN_FLAG_NO_SHADOW) == 0;
    ix->kp = (cfg->dim + (ix->metric == kMetricL2 ? 3 : 1) + 7) / 8 * 8;   // + augmented columns (common.cuh, append_kernels.cu)
    ix->device = cfg->device;
    ix->sm_count = prop.multiProcessorCount;
    ix->smem_optin = prop.sharedMemPerBlockOptin;
    auto cleanup = [&](int rc) {
        ann_destroy(ix);
        return rc;
    };
    if (cudaSetDevice(ix->device) != cudaSuccess) return cleanup(fail(ANN_ERR_CUDA, "cudaSetDevice failed"));
    if (cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ix->append_stream, cudaStreamNonBlocking) != cudaSuccess)
        return cleanup(fail(ANN_ERR_CUDA, "cudaStreamCreate failed"));
    if (cudaMalloc(&ix->scalars, sizeof(DeviceScalars)) != cudaSuccess)
        return cleanup(fail(ANN_ERR_OUT_OF_MEMORY, "cudaMalloc failed"));
    if (cudaHostAlloc(&ix->h_scalars, sizeof(DeviceScalars), cudaHostAllocDefault) != cudaSuccess)
        return cleanup(fail(ANN_ERR_OUT_OF_MEMORY, "cudaHostAlloc failed"));
    memset(ix->h_scalars, 0, sizeof(DeviceScalars));
    if (cudaHostAlloc(&ix->h_res, kSmallResultBytes, cudaHostAllocDefault) != cudaSuccess) {
        (void)cudaGetLastError();
        ix->h_res = nullptr;   // not fatal: results then go straight to the caller's buffers, as before
    }
    if (cudaMemsetAsync(ix->scalars, 0, sizeof(DeviceScalars), ix->stream) != cudaSuccess)
        return cleanup(fail(ANN_ERR_CUDA, "cudaMemset failed"));
    // Address ranges for the whole life of the index: as many rows as could ever fit the device (192 GB of fp32 rows, local
    // row numbers are 32 bit), halved until the reservation succeeds.  Physical memory follows the rows (grow()).
    if (!vm_api().ok) return cleanup(fail(ANN_ERR_CUDA, "ann_create: the driver's virtual memory management API is unavailable"));
    {
        const size_t row_bytes = (size_t)ix->pitch * sizeof(float);
        long long mr = (long long)std::min<unsigned long long>(0x7FFFFF80ull, (192ull << 30) / row_bytes) / 128 * 128;
        const long long floor_rows = std::max<long long>((cfg->capacity_hint + 127) / 128 * 128, 1 << 20);
        for (;;) {
            mr = std::max(mr, floor_rows);
            bool ok = ix->vm_rows.reserve((size_t)mr * row_bytes, ix->device) == cudaSuccess &&
                      ix->vm_ids.reserve((size_t)mr * sizeof(int64_t), ix->device) == cudaSuccess &&
                      ix->vm_norm.reserve((size_t)mr * sizeof(float), ix->device) == cudaSuccess &&
                      (ix->metric != kMetricCosine || ix->vm_inv.reserve((size_t)mr * sizeof(float), ix->device) == cudaSuccess) &&
                      (!ix->use_shadow || ix->vm_shadow.reserve((size_t)mr * ix->kp * sizeof(__nv_bfloat16), ix->device) == cudaSuccess);
            if (ok) break;
            ix->vm_rows.release();
            ix->vm_ids.release();
            ix->vm_norm.release();
            ix->vm_inv.release();
            ix->vm_shadow.release();
            if (mr <= floor_rows) return cleanup(fail(ANN_ERR_OUT_OF_MEMORY, "ann_create: reserving the address range for capacity_hint rows failed (more than the device can ever hold)"));
            mr /= 2;
        }
        ix->max_rows = mr;
        ix->rows = static_cast<float*>(ix->vm_rows.ptr());
        ix->ids = static_cast<int64_t*>(ix->vm_ids.ptr());
        ix->row_norm = static_cast<float*>(ix->vm_norm.ptr());
        ix->inv_norm = ix->metric == kMetricCosine ? static_cast<float*>(ix->vm_inv.ptr()) : nullptr;
        ix->shadow = ix->use_shadow ? static_cast<__nv_bfloat16*>(ix->vm_shadow.ptr()) : nullptr;
    }
    if (cfg->capacity_hint > 0) {
        int rc = grow(ix, cfg->capacity_hint, ix->stream);
        if (rc) return cleanup(rc);
    }
    if (cudaStreamSynchronize(ix->stream) != cudaSuccess) return cleanup(fail(ANN_ERR_CUDA, "stream sync failed"));
    *out = ix;
    return ANN_OK;
}

void ann_destroy(ann_index* ix) {
    if (!ix) return;
    cudaSetDevice(ix->device);
    cudaDeviceSynchronize();   // nothing may still read the ranges that are about to be unmapped
    ix->vm_rows.release();
    ix->vm_ids.release();
    ix->vm_norm.release();
    ix->vm_inv.release();
    ix->vm_shadow.release();
    cudaFree(ix->scalars);
    if (ix->h_scalars) cudaFreeHost(ix->h_scalars);
    if (ix->h_res) cudaFreeHost(ix->h_res);
    if (ix->co.pin_q) cudaFreeHost(ix->co.pin_q);
    if (ix->co.pin_res) cudaFreeHost(ix->co.pin_res);
    if (ix->append_stream) cudaStreamDestroy(ix->append_stream);
    ix->q_in.release();
    ix->q_unit.release();
    ix->q_padded.release();
    ix->out_dist.release();
    ix->q_shadow.release();
    ix->qstate.release();
    ix->pool.release();
    ix->special_rows.release();
    ix->pub_keys.release();
    ix->out_ids.release();
    ix->out_count.release();
    ix->fb_scratch.release();
    ix->upd_rows.release();
    ix->upd_slots.release();
    for (auto& pr : ix->ev_pairs) {
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
    }
    if (ix->stream) cudaStreamDestroy(ix->stream);
    (void)cudaGetLastError();
    delete ix;
}

int ann_size(const ann_index* ix, int64_t* n) {
    if (!ix || !n) return fail(ANN_ERR_NULL_POINTER, "ann_size: NULL argument");
    // lock-free like linkedQueue.size (BruteForceIndex.scala:34-36): a size() poll must not wait behind a running query
    // batch, which holds `mu` for its whole duration; pairs with the release store in append_and_publish
    *n = __atomic_load_n(&ix->n, __ATOMIC_ACQUIRE);
    return ANN_OK;
}

int ann_update_batch(ann_index* ix, const int64_t* slots, const float* rows, int64_t n) {
    if (!ix) return fail(ANN_ERR_NULL_POINTER, "ann_update_batch: index is NULL");
    if (n < 0) return fail(ANN_ERR_INVALID_ARGUMENT, "ann_update_batch: n < 0");
    if (n == 0) return ANN_OK;
    if (!slots || !rows) return fail(ANN_ERR_NULL_POINTER, "ann_update_batch: NULL buffer");
    std::lock_guard<std::mutex> alk(ix->append_mu);   // rewrites published rows: excludes appends AND queries
    std::lock_guard<std::mutex> lk(ix->mu);
    for (int64_t i = 0; i < n; ++i)
        if (slots[i] < 0 || slots[i] >= ix->n) return fail(ANN_ERR_INVALID_ARGUMENT, "ann_update_batch: slot outside [0, size)");
    int rc = set_device(ix);
    if (rc) return rc;
    cudaStream_t st = ix->stream;
    if (ix->sess.open) ix->sess.clobbered = true;   // published bounds may have been witnessed by rows that no longer exist
    CUDA_TRY(ix->upd_rows.ensure((size_t)n * ix->dim));
    CUDA_TRY(ix->upd_slots.ensure((size_t)n));
    CUDA_TRY(cudaMemcpyAsync(ix->upd_rows.p, rows, (size_t)n * ix->dim * sizeof(float), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(ix->upd_slots.p, slots, (size_t)n * sizeof(long long), cudaMemcpyHostToDevice, st));
    if (ix->unit_rows) {
        CUDA_TRY(launch_normalize_rows_device(ix->upd_rows.p, 0, n, ix->dim, ix->dim, nullptr, st));
        ix->launches++;
    }
    AppendParams ap{};
    ap.rows = ix->rows;
    ap.row0 = 0;
    ap.slots = ix->upd_slots.p;
    ap.staged = ix->upd_rows.p;
    ap.n_new = n;
    ap.dim = ix->dim;
    ap.pitch = ix->pitch;
    ap.metric = ix->metric;
    ap.inv_norm = ix->inv_norm;
    ap.row_norm = ix->row_norm;
    ap.shadow = ix->shadow;
    ap.kp = ix->kp;
    ap.max_norm_bits = &ix->scalars->max_norm_bits;   // only ever grows: the error bounds stay valid (looser)
    ap.max_resid_bits = &ix->scalars->max_resid_bits; // likewise
    ap.n_special = &ix->scalars->n_special;           // conservative: a repaired row does not lower the census
    ap.special_list = ix->scalars->special_list;
    if (ix->n_special > 0) ix->gemm_blocked_by_update = true;   // a repaired special row would be listed AND be a candidate
    CUDA_TRY(launch_append(ap, st));
    ix->launches++;
    CUDA_TRY(cudaStreamSynchronize(st));
    return check_device_flags(ix, st);
}

int ann_read_rows(ann_index* ix, int64_t start, int64_t n, int64_t* out_ids, float* out_rows) {
    if (!ix) return fail(ANN_ERR_NULL_POINTER, "ann_read_rows: index is NULL");
    std::lock_guard<std::mutex> lk(ix->mu);
    if (start < 0 || n < 0 || start + n > ix->n) return fail(ANN_ERR_INVALID_ARGUMENT, "ann_read_rows: range outside [0, size)");
    if (n == 0) return ANN_OK;
    int rc = set_device(ix);
    if (rc) return rc;
    if (out_ids) CUDA_TRY(cudaMemcpyAsync(out_ids, ix->ids + start, (size_t)n * sizeof(int64_t), cudaMemcpyDeviceToHost, ix->stream));
    if (out_rows)
        CUDA_TRY(cudaMemcpy2DAsync(out_rows, (size_t)ix->dim * sizeof(float), ix->rows + (size_t)start * ix->pitch,
                                   (size_t)ix->pitch * sizeof(float), (size_t)ix->dim * sizeof(float), (size_t)n,
                                   cudaMemcpyDeviceToHost, ix->stream));
    CUDA_TRY(cudaStreamSynchronize(ix->stream));
    return ANN_OK;
}

namespace {
// Shared tail of both append entry points.  Caller holds append_mu.  One synchronisation: the copy of the caller's rows,
// K1 and the read-back of the special-row census complete together; then the new row count is published under `mu`.
int append_and_publish(ann_index* ix, const int64_t* ids, const float* rows, int64_t n, cudaStream_t st) {
    int rc = append_device_core(ix, ids, rows, n, st);
    if (rc) return rc;
    // the special-row census decides which query kernels are legal, so it travels back with the same synchronisation
    CUDA_TRY(cudaMemcpyAsync(ix->h_scalars, ix->scalars, kScalarsHeader, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    std::lock_guard<std::mutex> lk(ix->mu);
    if (ix->sess.open) ix->sess.clobbered = true;   // new rows may raise the error bounds a pending sharded query was prepared with
    __atomic_store_n(&ix->n, ix->n + n, __ATOMIC_RELEASE);   // ann_size reads it without `mu`
    ix->n_special = ix->h_scalars->n_special;
    return ANN_OK;
}
}  // namespace

int ann_append_batch(ann_index* ix, const int64_t* ids, const float* rows, int64_t n) {
    if (!ix) return fail(ANN_ERR_NULL_POINTER, "ann_append_batch: index is NULL");
    if (n < 0) return fail(ANN_ERR_INVALID_ARGUMENT, "ann_append_batch: n < 0");
    if (n == 0) return ANN_OK;
    if (!rows) return fail(ANN_ERR_NULL_POINTER, "ann_append_batch: rows is NULL");
    std::lock_guard<std::mutex> alk(ix->append_mu);
    int rc = set_device(ix);
    if (rc) return rc;
    // host pointers go straight into place (cudaMemcpyDefault): no staging copy of the batch
    return append_and_publish(ix, ids, rows, n, ix->append_stream);
}

int ann_append_batch_device(ann_index* ix, const int64_t* d_ids, const float* d_rows, int64_t n, void* stream) {
    if (!ix) return fail(ANN_ERR_NULL_POINTER, "ann_append_batch_device: index is NULL");
    if (n < 0) return fail(ANN_ERR_INVALID_ARGUMENT, "ann_append_batch_device: n < 0");
    if (n == 0) return ANN_OK;
    if (!d_rows) return fail(ANN_ERR_NULL_POINTER, "ann_append_batch_device: rows is NULL");
    std::lock_guard<std::mutex> alk(ix->append_mu);
    int rc = set_device(ix);
    if (rc) return rc;
    // NULL = the legacy default stream, as everywhere in CUDA
    return append_and_publish(ix, d_ids, d_rows, n, (cudaStream_t)stream);
}

int ann_query_batch_device(ann_index* ix, const float* d_queries, int32_t b, int32_t dim, int32_t k, int64_t* d_out_ids,
                           float* d_out_dist, int32_t* d_out_count, void* stream) {
    if (!ix) return fail(ANN_ERR_NULL_POINTER, "ann_query_batch_device: index is NULL");
    if (b < 0) return fail(ANN_ERR_INVALID_ARGUMENT, "ann_query_batch_device: b < 0");
    if (k < 0) return fail(ANN_ERR_NEGATIVE_K, "ann_query_batch_device: k < 0");
    if (dim != ix->dim) return fail(ANN_ERR_DIMENSION_MISMATCH, "ann_query_batch_device: query dimension != index dimension");
    if (b == 0) return ANN_OK;
    if (!d_queries || (k > 0 && (!d_out_ids || !d_out_dist)))
        return fail(ANN_ERR_NULL_POINTER, "ann_query_batch_device: NULL buffer");
    std::lock_guard<std::mutex> lk(ix->mu);
    int rc = set_device(ix);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;  // NULL = the legacy default stream
    if (ix->sess.open) ix->sess.clobbered = true;   // the per-query scratch of a pending two-phase query is overwritten
    rc = unit_queries(ix, &d_queries, b, st);
    if (rc) return rc;
    rc = query_core(ix, d_queries, b, k, d_out_ids, d_out_dist, d_out_count, st);
    if (rc == ANN_OK && ix->device_fallback) rc = resolve_flagged(ix, d_queries, b, k, d_out_ids, d_out_dist, d_out_count, st);
    return rc;
}

// ---- two-phase sharded query: seed (publish this shard's bounds) -> [caller: cross-shard barrier] -> finish ----
namespace {
int seed_impl(ann_index* ix, const float* d_queries, int32_t b, int32_t dim, int32_t k, uint32_t* d_seed_keys,
              uint32_t* const* push_dst, int32_t n_push, void* stream, const SliceSeed* slice);
}
int ann_query_seed_device(ann_index* ix, const float* d_queries, int32_t b, int32_t dim, int32_t k, uint32_t* d_seed_keys,
                          void* stream) {
    return seed_impl(ix, d_queries, b, dim, k, d_seed_keys, nullptr, 0, stream, nullptr);
}
int ann_query_seed_slice_push_device(ann_index* ix, const float* d_queries, int32_t b, int32_t dim, int32_t k, int32_t q_begin,
                                     int32_t q_count, int32_t n_slices, uint32_t* const* dst, int32_t n_dst, void* stream) {
    if (n_dst < 1 || n_dst > kMaxPeers || !dst) return fail(ANN_ERR_INVALID_ARGUMENT, "ann_query_seed_slice_push_device: n_dst must be in [1, 16]");
    for (int i = 0; i < n_dst; ++i)
        if (!dst[i]) return fail(ANN_ERR_NULL_POINTER, "ann_query_seed_slice_push_device: NULL destination");
    if (q_begin < 0 || q_count < 0 || (long long)q_begin + q_count > b)
        return fail(ANN_ERR_INVALID_ARGUMENT, "ann_query_seed_slice_push_device: query slice outside the batch");
    if (n_slices < 1 || n_slices > kMaxPeers) return fail(ANN_ERR_INVALID_ARGUMENT, "ann_query_seed_slice_push_device: n_slices must be in [1, 16]");
    SliceSeed sl;
    sl.q_begin = q_begin;
    sl.q_count = q_count;
    sl.scale = n_slices;
    return seed_impl(ix, d_queries, b, dim, k, nullptr, dst, n_dst, stream, &sl);
}
int ann_query_seed_push_device(ann_index* ix, const float* d_queries, int32_t b, int32_t dim, int32_t k, uint32_t* const* dst,
                               int32_t n_dst, void* stream) {
    if (n_dst < 1 || n_dst > kMaxPeers || !dst) return fail(ANN_ERR_INVALID_ARGUMENT, "ann_query_seed_push_device: n_dst must be in [1, 16]");
    for (int i = 0; i < n_dst; ++i)
        if (!dst[i]) return fail(ANN_ERR_NULL_POINTER, "ann_query_seed_push_device: NULL destination");
    return seed_impl(ix, d_queries, b, dim, k, nullptr, dst, n_dst, stream, nullptr);
}
namespace {
int seed_impl(ann_index* ix, const float* d_queries, int32_t b, int32_t dim, int32_t k, uint32_t* d_seed_keys,
              uint32_t* const* push_dst, int32_t n_push, void* stream, const SliceSeed* slice) {
    if (!ix) return fail(ANN_ERR_NULL_POINTER, "ann_query_seed_device: index is NULL");
    if (b < 0) return fail(ANN_ERR_INVALID_ARGUMENT, "ann_query_seed_device: b < 0");
    if (k < 0) return fail(ANN_ERR_NEGATIVE_K, "ann_query_seed_device: k < 0");
    if (dim != ix->dim) return fail(ANN_ERR_DIMENSION_MISMATCH, "ann_query_seed_device: query dimension != index dimension");
    if (b > 0 && (!d_queries || (k > 0 && !d_seed_keys && n_push == 0))) return fail(ANN_ERR_NULL_POINTER, "ann_query_seed_device: NULL buffer");
    std::lock_guard<std::mutex> lk(ix->mu);
    int rc = set_device(ix);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    ix->sess = ann_index::SeedSession{};
    ix->sess.open = true;
    ix->sess.b = b;
    ix->sess.k = k;
    if (b == 0 || k == 0) return ANN_OK;
    const int k_eff = (int)std::min<long long>(k, ix->n);
    constexpr int kMaxGemmBatch = 16384;
    const bool gemm = ix->n > 0 && k_eff <= kMaxK && ix->path_opt != 3 && ix->path_opt != 1 && b <= kMaxGemmBatch &&
                      gemm_eligible(ix, b, k_eff) && (ix->path_opt == 2 || b >= ix->gemm_min_batch);
    ix->sess.sliced = slice != nullptr;
    if (!gemm) {   // scan / exact paths keep their own thresholds: nothing to publish, the finish call runs the whole query
        if (slice) {   // one "no bound" per query of the slice
            for (int i = 0; i < n_push && slice->q_count > 0; ++i)
                CUDA_TRY(cudaMemsetAsync(push_dst[i] + slice->q_begin, 0xFF, (size_t)slice->q_count * sizeof(uint32_t), st));
        } else if (n_push > 0) {
            for (int i = 0; i < n_push; ++i) CUDA_TRY(cudaMemsetAsync(push_dst[i], 0xFF, (size_t)b * k * sizeof(uint32_t), st));
        } else {
            CUDA_TRY(cudaMemsetAsync(d_seed_keys, 0xFF, (size_t)b * k * sizeof(uint32_t), st));
        }
        return ANN_OK;
    }
    ix->sess.gemm = true;
    CUDA_TRY(ix->qstate.ensure((size_t)b));
    rc = unit_queries(ix, &d_queries, b, st);
    if (rc) return rc;
    rc = query_gemm(ix, ix->qstate.p, d_queries, b, k_eff, k, nullptr, nullptr, nullptr, st, 1, d_seed_keys, nullptr, 1, push_dst, n_push, slice);
    if (rc) ix->sess = ann_index::SeedSession{};
    return rc;
}
}  // namespace

int ann_query_finish_device(ann_index* ix, const float* d_queries, int32_t b, int32_t dim, int32_t k,
                            const uint32_t* const* peer_seed_keys, int32_t world, int64_t* d_out_ids, float* d_out_dist,
                            int32_t* d_out_count, void* stream) {
    if (!ix) return fail(ANN_ERR_NULL_POINTER, "ann_query_finish_device: index is NULL");
    if (b < 0) return fail(ANN_ERR_INVALID_ARGUMENT, "ann_query_finish_device: b < 0");
    if (k < 0) return fail(ANN_ERR_NEGATIVE_K, "ann_query_finish_device: k < 0");
    if (dim != ix->dim) return fail(ANN_ERR_DIMENSION_MISMATCH, "ann_query_finish_device: query dimension != index dimension");
    if (world < 0 || world > kMaxPeers) return fail(ANN_ERR_INVALID_ARGUMENT, "ann_query_finish_device: world must be in [0, 16]");
    if (b > 0 && (!d_queries || (k > 0 && (!d_out_ids || !d_out_dist))))
        return fail(ANN_ERR_NULL_POINTER, "ann_query_finish_device: NULL buffer");
    std::lock_guard<std::mutex> lk(ix->mu);
    int rc = set_device(ix);
    if (rc) return rc;
    const ann_index::SeedSession sess = ix->sess;
    if (!sess.open || sess.b != b || sess.k != k || sess.filtered)
        return fail(ANN_ERR_INVALID_ARGUMENT, "ann_query_finish_device: no pending ann_query_seed_device call with this (b, k)");
    if (sess.sliced) {
        ix->sess.open = false;
        return fail(ANN_ERR_INVALID_ARGUMENT,
                    "ann_query_finish_device: the pending seed call was sliced (one bound per query); ann_query_filter_bounds_push_device must follow it");
    }
    if (sess.clobbered) {
        ix->sess.open = false;
        return fail(ANN_ERR_INVALID_ARGUMENT,
                    "ann_query_finish_device: another query, append or update ran on this index between the seed and finish calls");
    }
    if (b == 0) {
        ix->sess.open = false;
        return ANN_OK;
    }
    cudaStream_t st = (cudaStream_t)stream;
    rc = unit_queries(ix, &d_queries, b, st);
    if (rc) return rc;
    if (!sess.gemm) {
        ix->sess.open = false;
        rc = query_core(ix, d_queries, b, k, d_out_ids, d_out_dist, d_out_count, st);
    } else {
        PeerSeedKeys pk{};
        int w = 0;
        if (peer_seed_keys)
            for (int s2 = 0; s2 < world; ++s2) {
                if (!peer_seed_keys[s2]) {
                    ix->sess.open = false;
                    return fail(ANN_ERR_NULL_POINTER, "ann_query_finish_device: NULL seed-key pointer");
                }
                pk.keys[w++] = peer_seed_keys[s2];
            }
        const int k_eff = (int)std::min<long long>(k, ix->n);
        rc = query_gemm(ix, ix->qstate.p, d_queries, b, k_eff, k, d_out_ids, d_out_dist, d_out_count, st, 2, nullptr, &pk, w);
        ix->sess.open = false;
    }
    if (rc == ANN_OK && ix->device_fallback) rc = resolve_flagged(ix, d_queries, b, k, d_out_ids, d_out_dist, d_out_count, st);
    return rc;
}

// ---- three-phase sharded query: seed -> [barrier] -> filter (publishes the k best bounds) -> [barrier] -> rescore ----
namespace {
int check_session(ann_index* ix, const char* who, int32_t b, int32_t dim, int32_t k, int32_t world, bool want_filtered) {
    char msg[256];
    if (b < 0) { snprintf(msg, sizeof(msg), "%s: b < 0", who); return fail(ANN_ERR_INVALID_ARGUMENT, msg); }
    if (k < 0) { snprintf(msg, sizeof(msg), "%s: k < 0", who); return fail(ANN_ERR_NEGATIVE_K, msg); }
    if (dim != ix->dim) { snprintf(msg, sizeof(msg), "%s: query dimension != index dimension", who); return fail(ANN_ERR_DIMENSION_MISMATCH, msg); }
    if (world < 0 || world > kMaxPeers) { snprintf(msg, sizeof(msg), "%s: world must be in [0, 16]", who); return fail(ANN_ERR_INVALID_ARGUMENT, msg); }
    const ann_index::SeedSession& sess = ix->sess;
    if (!sess.open || sess.b != b || sess.k != k || sess.filtered != want_filtered) {
        snprintf(msg, sizeof(msg), "%s: out of sequence (expected ann_query_seed_device -> ann_query_filter_device -> "
                 "ann_query_rescore_device with the same (b, k))", who);
        return fail(ANN_ERR_INVALID_ARGUMENT, msg);
    }
    if (sess.clobbered) {
        ix->sess.open = false;
        snprintf(msg, sizeof(msg), "%s: another query, append or update ran on this index since ann_query_seed_device", who);
        return fail(ANN_ERR_INVALID_ARGUMENT, msg);
    }
    return ANN_OK;
}
}  // namespace

namespace {
int filter_impl(ann_index* ix, const float* d_queries, int32_t b, int32_t dim, int32_t k, const uint32_t* const* peer_seed_keys,
                int32_t world, uint32_t* d_kth_keys, uint32_t* const* push_dst, int32_t n_push, void* stream, const uint32_t* d_bounds);
}
int ann_query_filter_device(ann_index* ix, const float* d_queries, int32_t b, int32_t dim, int32_t k,
                            const uint32_t* const* peer_seed_keys, int32_t world, uint32_t* d_kth_keys, void* stream) {
    return filter_impl(ix, d_queries, b, dim, k, peer_seed_keys, world, d_kth_keys, nullptr, 0, stream, nullptr);
}
int ann_query_filter_bounds_push_device(ann_index* ix, const float* d_queries, int32_t b, int32_t dim, int32_t k,
                                        const uint32_t* d_bounds, int32_t world, uint32_t* const* kth_dst, int32_t n_dst, void* stream) {
    if (n_dst < 1 || n_dst > kMaxPeers || !kth_dst) return fail(ANN_ERR_INVALID_ARGUMENT, "ann_query_filter_bounds_push_device: n_dst must be in [1, 16]");
    for (int i = 0; i < n_dst; ++i)
        if (!kth_dst[i]) return fail(ANN_ERR_NULL_POINTER, "ann_query_filter_bounds_push_device: NULL destination");
    if (b > 0 && k > 0 && !d_bounds) return fail(ANN_ERR_NULL_POINTER, "ann_query_filter_bounds_push_device: NULL bound array");
    return filter_impl(ix, d_queries, b, dim, k, nullptr, world, nullptr, kth_dst, n_dst, stream, d_bounds);
}
int ann_query_filter_push_device(ann_index* ix, const float* d_queries, int32_t b, int32_t dim, int32_t k,
                                 const uint32_t* const* peer_seed_keys, int32_t world, uint32_t* const* kth_dst, int32_t n_dst,
                                 void* stream) {
    if (n_dst < 1 || n_dst > kMaxPeers || !kth_dst) return fail(ANN_ERR_INVALID_ARGUMENT, "ann_query_filter_push_device: n_dst must be in [1, 16]");
    for (int i = 0; i < n_dst; ++i)
        if (!kth_dst[i]) return fail(ANN_ERR_NULL_POINTER, "ann_query_filter_push_device: NULL destination");
    return filter_impl(ix, d_queries, b, dim, k, peer_seed_keys, world, nullptr, kth_dst, n_dst, stream, nullptr);
}
namespace {
int filter_impl(ann_index* ix, const float* d_queries, int32_t b, int32_t dim, int32_t k, const uint32_t* const* peer_seed_keys,
                int32_t world, uint32_t* d_kth_keys, uint32_t* const* push_dst, int32_t n_push, void* stream, const uint32_t* d_bounds) {
    if (!ix) return fail(ANN_ERR_NULL_POINTER, "ann_query_filter_device: index is NULL");
    std::lock_guard<std::mutex> lk(ix->mu);
    int rc = check_session(ix, "ann_query_filter_device", b, dim, k, world, false);
    if (rc) return rc;
    if (b > 0 && (!d_queries || (k > 0 && !d_kth_keys && n_push == 0))) {
        ix->sess.open = false;
        return fail(ANN_ERR_NULL_POINTER, "ann_query_filter_device: NULL buffer");
    }
    rc = set_device(ix);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    ix->sess.filtered = true;
    if (b == 0 || k == 0) return ANN_OK;
    if (!ix->sess.gemm) {   // scan / exact paths: nothing to publish, the rescore call answers the whole local query
        if (n_push > 0) {
            for (int i = 0; i < n_push; ++i) CUDA_TRY(cudaMemsetAsync(push_dst[i], 0xFF, (size_t)b * k * sizeof(uint32_t), st));
        } else {
            CUDA_TRY(cudaMemsetAsync(d_kth_keys, 0xFF, (size_t)b * k * sizeof(uint32_t), st));
        }
        return ANN_OK;
    }
    PeerSeedKeys pk{};
    int w = 0;
    if (peer_seed_keys)
        for (int s2 = 0; s2 < world; ++s2) {
            if (!peer_seed_keys[s2]) {
                ix->sess.open = false;
                return fail(ANN_ERR_NULL_POINTER, "ann_query_filter_device: NULL seed-key pointer");
            }
            pk.keys[w++] = peer_seed_keys[s2];
        }
    const int k_eff = (int)std::min<long long>(k, ix->n);
    if ((d_bounds != nullptr) != ix->sess.sliced) {
        ix->sess.open = false;
        return fail(ANN_ERR_INVALID_ARGUMENT, "ann_query_filter_device: the seed call and the filter call must both be sliced or both be not");
    }
    SliceSeed sl;
    sl.bounds = d_bounds;
    rc = query_gemm(ix, ix->qstate.p, d_queries, b, k_eff, k, nullptr, nullptr, nullptr, st, 3, d_kth_keys, &pk, w, push_dst, n_push,
                    d_bounds ? &sl : nullptr);
    if (rc) ix->sess.open = false;
    return rc;
}
}  // namespace

int ann_query_rescore_device(ann_index* ix, const float* d_queries, int32_t b, int32_t dim, int32_t k,
                             const uint32_t* const* peer_kth_keys, int32_t world, int64_t* d_out_ids, float* d_out_dist,
                             int32_t* d_out_count, void* stream) {
    if (!ix) return fail(ANN_ERR_NULL_POINTER, "ann_query_rescore_device: index is NULL");
    std::lock_guard<std::mutex> lk(ix->mu);
    int rc = check_session(ix, "ann_query_rescore_device", b, dim, k, world, true);
    if (rc) return rc;
    const bool gemm = ix->sess.gemm;
    ix->sess.open = false;
    if (b == 0) return ANN_OK;
    if (!d_queries || (k > 0 && (!d_out_ids || !d_out_dist))) return fail(ANN_ERR_NULL_POINTER, "ann_query_rescore_device: NULL buffer");
    rc = set_device(ix);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    rc = unit_queries(ix, &d_queries, b, st);
    if (rc) return rc;
    if (!gemm) {
        rc = query_core(ix, d_queries, b, k, d_out_ids, d_out_dist, d_out_count, st);
    } else {
        PeerSeedKeys pk{};
        int w = 0;
        if (peer_kth_keys)
            for (int s2 = 0; s2 < world; ++s2) {
                if (!peer_kth_keys[s2]) return fail(ANN_ERR_NULL_POINTER, "ann_query_rescore_device: NULL key pointer");
                pk.keys[w++] = peer_kth_keys[s2];
            }
        const int k_eff = (int)std::min<long long>(k, ix->n);
        // the published arrays have pitch k; with k_eff < k (shard smaller than k) every row of the shard is a candidate anyway
        rc = query_gemm(ix, ix->qstate.p, d_queries, b, k_eff, k, d_out_ids, d_out_dist, d_out_count, st, 4, nullptr,
                        k_eff == k ? &pk : nullptr, k_eff == k ? w : 0);
    }
    if (rc == ANN_OK && ix->device_fallback) rc = resolve_flagged(ix, d_queries, b, k, d_out_ids, d_out_dist, d_out_count, st);
    return rc;
}

int ann_exchange_merge_slice_device(int32_t device, const void* const* peer_local, int32_t world, int32_t b, int32_t k,
                                    int32_t q_begin, int32_t q_count, int64_t* d_out_ids, float* d_out_dist, int32_t* d_out_count,
                                    void* stream) {
    if (world < 1 || world > kMaxPeers) return fail(ANN_ERR_INVALID_ARGUMENT, "ann_exchange_merge_slice_device: world must be in [1, 16]");
    if (b < 0 || q_begin < 0 || q_count < 0 || (long long)q_begin + q_count > b)
        return fail(ANN_ERR_INVALID_ARGUMENT, "ann_exchange_merge_slice_device: query range outside the batch");
    if (k < 0) return fail(ANN_ERR_NEGATIVE_K, "ann_exchange_merge_slice_device: k < 0");
    if (!peer_local) return fail(ANN_ERR_NULL_POINTER, "ann_exchange_merge_slice_device: NULL pointer table");
    if (b == 0 || k == 0 || q_count == 0) return ANN_OK;
    if (!d_out_ids || !d_out_dist) return fail(ANN_ERR_NULL_POINTER, "ann_exchange_merge_slice_device: NULL output buffer");
    if ((long long)(world + 1) * k * 12 > 200 * 1024) return fail(ANN_ERR_INVALID_ARGUMENT, "ann_exchange_merge_slice_device: (world + 1) * k too large");
    PeerBlocks pb{};
    for (int s = 0; s < world; ++s) {
        if (!peer_local[s]) return fail(ANN_ERR_NULL_POINTER, "ann_exchange_merge_slice_device: NULL block pointer");
        pb.local[s] = static_cast<const unsigned char*>(peer_local[s]);
    }
    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(launch_exchange_merge(pb, world, b, k, q_begin, q_count, (cudaStream_t)stream, d_out_ids, d_out_dist, d_out_count));
    return ANN_OK;
}

namespace {

// The host-buffer query: H2D, query path, D2H, exact fallback for flagged queries.  Takes `mu`.
int query_host(ann_index* ix, const float* queries, int32_t b, int32_t k, int64_t* out_ids, float* out_dist, int32_t* out_count,
               bool dest_pinned = false) {
    std::lock_guard<std::mutex> lk(ix->mu);
    int rc = set_device(ix);
    if (rc) return rc;
    cudaStream_t st = ix->stream;
    if (ix->sess.open) ix->sess.clobbered = true;
    const size_t nk = (size_t)b * (size_t)std::max(k, 1);
    CUDA_TRY(ix->q_in.ensure((size_t)b * ix->dim));
    CUDA_TRY(ix->out_ids.ensure(nk));
    CUDA_TRY(ix->out_dist.ensure(nk));
    CUDA_TRY(ix->out_count.ensure((size_t)b));
    CUDA_TRY(cudaMemcpyAsync(ix->q_in.p, queries, (size_t)b * ix->dim * sizeof(float), cudaMemcpyHostToDevice, st));
    const float* d_q = ix->q_in.p;
    rc = unit_queries(ix, &d_q, b, st);
    if (rc) return rc;
    rc = query_core(ix, d_q, b, k, ix->out_ids.p, ix->out_dist.p, ix->out_count.p, st);
    if (rc) return rc;
    // one round trip in the common case: results and the sticky flag word come back together
    DeviceScalars hs{};
    const size_t ids_bytes = (size_t)b * k * sizeof(int64_t), dist_bytes = (size_t)b * k * sizeof(float), cnt_bytes = (size_t)b * sizeof(int32_t);
    const size_t hdr_off = (ids_bytes + dist_bytes + cnt_bytes + 15) / 16 * 16;
    const bool landing = !dest_pinned && ix->h_res != nullptr && hdr_off + kScalarsHeader <= kSmallResultBytes;   // see ann_index::h_res
    auto copy_out = [&]() -> int {
        unsigned char* const h_ids = landing ? ix->h_res : reinterpret_cast<unsigned char*>(out_ids);
        unsigned char* const h_dist = landing ? ix->h_res + ids_bytes : reinterpret_cast<unsigned char*>(out_dist);
        unsigned char* const h_cnt = landing ? ix->h_res + ids_bytes + dist_bytes : reinterpret_cast<unsigned char*>(out_count);
        void* const h_hdr = landing ? static_cast<void*>(ix->h_res + hdr_off) : static_cast<void*>(&hs);
        if (k > 0) {
            CUDA_TRY(cudaMemcpyAsync(h_ids, ix->out_ids.p, ids_bytes, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaMemcpyAsync(h_dist, ix->out_dist.p, dist_bytes, cudaMemcpyDeviceToHost, st));
        }
        if (out_count) CUDA_TRY(cudaMemcpyAsync(h_cnt, ix->out_count.p, cnt_bytes, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(h_hdr, ix->scalars, kScalarsHeader, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        if (landing) {
            if (k > 0) {
                memcpy(out_ids, h_ids, ids_bytes);
                memcpy(out_dist, h_dist, dist_bytes);
            }
            if (out_count) memcpy(out_count, h_cnt, cnt_bytes);
            memcpy(&hs, h_hdr, kScalarsHeader);
        }
        return ANN_OK;
    };
    rc = copy_out();
    if (rc) return rc;
    if (hs.error_flags) {   // rare: some queries were flagged by the bounded selector -> exact fallback, then copy again
        bool ran = false;
        rc = resolve_flagged(ix, d_q, b, k, ix->out_ids.p, ix->out_dist.p, ix->out_count.p, st, &hs, &ran);
        if (rc) return rc;
        rc = copy_out();
        if (rc) return rc;
        if (hs.error_flags) return check_device_flags(ix, st);
    }
    return ANN_OK;
}

}  // namespace

// One waiting host call of the micro-batcher.
struct ann_index::PendingQuery {
    const float* q;
    int b, k;
    int64_t* ids;
    float* dist;
    int32_t* cnt;
    int rc = ANN_OK;
    std::string err;
    bool done = false, promoted = false;
};

namespace {

// Online callers issue ONE vector per call from many threads (QueryIndexThriftController.scala:39-90 runs each RPC's query on
// a FuturePool thread, UnsafeQueryIndexServer.scala:26-32,64-68), and a device call costs about a millisecond of corpus
// streaming whatever its batch size.  So concurrent small calls on one handle are combined: the first caller to arrive while
// no merged call is running becomes the leader, takes every waiting call with its k (up to max_batch queries), packs the
// queries into pinned staging, runs ONE device batch (tensor-core path from 2 queries up), scatters the rows back and wakes
// the others; whatever arrived meanwhile forms the next merged call, led by the first of them.  No timer: a lone caller
// runs at once, and under load the batch size adapts to the arrival rate.  Results are identical to separate calls (every
// query is answered exactly, independently of its batch).
int query_coalesced(ann_index* ix, const float* queries, int32_t b, int32_t k, int64_t* out_ids, float* out_dist, int32_t* out_count) {
    ann_index::Coalescer& co = ix->co;
    ann_index::PendingQuery me;
    me.q = queries;
    me.b = b;
    me.k = k;
    me.ids = out_ids;
    me.dist = out_dist;
    me.cnt = out_count;
    std::vector<ann_index::PendingQuery*> batch;
    {
        std::unique_lock<std::mutex> lk(co.mu);
        co.pending.push_back(&me);
        co.n_pending.store((int)co.pending.size(), std::memory_order_release);
        if (co.leader_active) {
            co.cv.wait(lk, [&] { return me.done || me.promoted; });
            if (me.done) {
                if (me.rc) return fail(me.rc, me.err);
                return ANN_OK;
            }
        }
        co.leader_active = true;
        // The batch that just finished released its callers a moment ago (closed-loop clients re-issue within microseconds).
        // Cutting the next batch at once would leave them behind and two half-sized groups would ping-pong forever (measured:
        // 32 + 32 instead of 64 calls per device batch), so whoever leads right after a multi-call batch waits until that many
        // calls are waiting again -- at most linger_us.  A lone caller never waits.
        if (co.linger_us > 0 && co.last_batch_calls > 1) {
            const auto now = std::chrono::steady_clock::now();
            const auto until = std::max(now, co.last_done) + std::chrono::microseconds(co.linger_us);
            if (now - co.last_done < std::chrono::microseconds(4 * co.linger_us)) {
                const int want = std::min(co.last_batch_calls, co.max_batch);
                lk.unlock();   // spin (timed waits have ~50 us of timer slack, longer than the whole linger); arrivals need the lock
                while (co.n_pending.load(std::memory_order_acquire) < want && std::chrono::steady_clock::now() < until) {
                }
                lk.lock();
            }
        }
        // my batch: me, then every waiting call with my k in arrival order while the merged batch stays within max_batch
        co.pending.erase(std::find(co.pending.begin(), co.pending.end(), &me));
        batch.push_back(&me);
        int taken = b;
        for (auto it = co.pending.begin(); it != co.pending.end();) {
            if ((*it)->k == k && taken + (*it)->b <= co.max_batch) {
                taken += (*it)->b;
                batch.push_back(*it);
                it = co.pending.erase(it);
            } else {
                ++it;
            }
        }
        co.n_pending.store((int)co.pending.size(), std::memory_order_release);
    }
    int total = 0;
    for (auto* r : batch) total += r->b;
    int rc = ANN_OK;
    std::string err;
    if (batch.size() == 1) {
        rc = query_host(ix, queries, b, k, out_ids, out_dist, out_count);
        if (rc) err = g_last_error;
    } else {
        // pinned staging: [queries total*dim] and [ids total*k | dist total*k | count total]
        const size_t qb = (size_t)total * ix->dim * sizeof(float), rb = (size_t)total * k * 12 + (size_t)total * 4;
        cudaSetDevice(ix->device);
        if (qb > co.cap_q) {
            if (co.pin_q) cudaFreeHost(co.pin_q);
            co.pin_q = nullptr;
            co.cap_q = 0;
            if (cudaHostAlloc(&co.pin_q, std::max(qb, (size_t)1 << 20), cudaHostAllocDefault) == cudaSuccess) co.cap_q = std::max(qb, (size_t)1 << 20);
        }
        if (rb > co.cap_res) {
            if (co.pin_res) cudaFreeHost(co.pin_res);
            co.pin_res = nullptr;
            co.cap_res = 0;
            if (cudaHostAlloc(&co.pin_res, std::max(rb, (size_t)1 << 20), cudaHostAllocDefault) == cudaSuccess) co.cap_res = std::max(rb, (size_t)1 << 20);
        }
        if (co.cap_q < qb || co.cap_res < rb) {
            (void)cudaGetLastError();
            rc = ANN_ERR_OUT_OF_MEMORY;
            err = "micro-batcher: pinned staging allocation failed";
        } else {
            size_t off = 0;
            for (auto* r : batch) {
                memcpy(co.pin_q + off, r->q, (size_t)r->b * ix->dim * sizeof(float));
                off += (size_t)r->b * ix->dim;
            }
            int64_t* s_ids = reinterpret_cast<int64_t*>(co.pin_res);
            float* s_dist = reinterpret_cast<float*>(co.pin_res + (size_t)total * k * 8);
            int32_t* s_cnt = reinterpret_cast<int32_t*>(co.pin_res + (size_t)total * k * 12);
            rc = query_host(ix, co.pin_q, total, k, s_ids, s_dist, s_cnt, /*dest_pinned=*/true);
            if (rc) err = g_last_error;
            else {
                size_t row = 0;
                for (auto* r : batch) {
                    memcpy(r->ids, s_ids + row * k, (size_t)r->b * k * sizeof(int64_t));
                    memcpy(r->dist, s_dist + row * k, (size_t)r->b * k * sizeof(float));
                    if (r->cnt) memcpy(r->cnt, s_cnt + row, (size_t)r->b * sizeof(int32_t));
                    row += (size_t)r->b;
                }
            }
        }
    }
    {
        std::lock_guard<std::mutex> lk(co.mu);
        co.batches++;
        co.merged_calls += (long long)batch.size();
        co.last_batch_calls = (int)batch.size();
        co.last_done = std::chrono::steady_clock::now();
        for (auto* r : batch) {
            if (r == &me) continue;
            r->rc = rc;
            r->err = err;
            r->done = true;
        }
        if (!co.pending.empty()) co.pending.front()->promoted = true;   // hand the lead to the first call that arrived meanwhile
        else co.leader_active = false;
    }
    co.cv.notify_all();   // one broadcast (63 separate wake-ups cost the callers 100-300 us of their next batch's linger)
    if (rc) return fail(rc, err);
    return ANN_OK;
}

}  // namespace

int ann_query_batch(ann_index* ix, const float* queries, int32_t b, int32_t dim, int32_t k, int64_t* out_ids, float* out_dist,
                    int32_t* out_count) {
    if (!ix) return fail(ANN_ERR_NULL_POINTER, "ann_query_batch: index is NULL");
    if (b < 0) return fail(ANN_ERR_INVALID_ARGUMENT, "ann_query_batch: b < 0");
    if (k < 0) return fail(ANN_ERR_NEGATIVE_K, "ann_query_batch: k < 0");
    if (dim != ix->dim) return fail(ANN_ERR_DIMENSION_MISMATCH, "ann_query_batch: query dimension != index dimension");
    if (b == 0) return ANN_OK;
    if (!queries || (k > 0 && (!out_ids || !out_dist))) return fail(ANN_ERR_NULL_POINTER, "ann_query_batch: NULL buffer");
    if (k > 0 && ix->co.max_batch > 0 && b <= ix->co.small_b && b <= ix->co.max_batch)
        return query_coalesced(ix, queries, b, k, out_ids, out_dist, out_count);
    return query_host(ix, queries, b, k, out_ids, out_dist, out_count);
}

int ann_merge_topk_device(int32_t device, const int64_t* d_ids, const float* d_dist, const int32_t* d_count, int32_t shards,
                          int32_t b, int32_t k, int64_t* d_out_ids, float* d_out_dist, int32_t* d_out_count, void* stream) {
    if (shards < 1 || b < 0) return fail(ANN_ERR_INVALID_ARGUMENT, "ann_merge_topk_device: shards < 1 or b < 0");
    if (k < 0) return fail(ANN_ERR_NEGATIVE_K, "ann_merge_topk_device: k < 0");
    if (b == 0 || k == 0) return ANN_OK;
    if (!d_ids || !d_dist || !d_count || !d_out_ids || !d_out_dist)
        return fail(ANN_ERR_NULL_POINTER, "ann_merge_topk_device: NULL buffer");
    if ((long long)shards * k > 16384) return fail(ANN_ERR_INVALID_ARGUMENT, "ann_merge_topk_device: shards*k > 16384");
    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(launch_merge(d_ids, d_dist, d_count, shards, b, k, d_out_ids, d_out_dist, d_out_count, (cudaStream_t)stream));
    return ANN_OK;
}

int ann_exchange_merge_device(int32_t device, const void* const* peer_local, void* const* peer_final, int32_t world, int32_t b,
                              int32_t k, int32_t q_begin, int32_t q_count, void* stream) {
    if (world < 1 || world > kMaxPeers) return fail(ANN_ERR_INVALID_ARGUMENT, "ann_exchange_merge_device: world must be in [1, 16]");
    if (b < 0 || q_begin < 0 || q_count < 0 || (long long)q_begin + q_count > b)
        return fail(ANN_ERR_INVALID_ARGUMENT, "ann_exchange_merge_device: query range outside the batch");
    if (k < 0) return fail(ANN_ERR_NEGATIVE_K, "ann_exchange_merge_device: k < 0");
    if (!peer_local || !peer_final) return fail(ANN_ERR_NULL_POINTER, "ann_exchange_merge_device: NULL pointer table");
    if (b == 0 || k == 0 || q_count == 0) return ANN_OK;
    if ((long long)(world + 1) * k * 12 > 200 * 1024) return fail(ANN_ERR_INVALID_ARGUMENT, "ann_exchange_merge_device: (world + 1) * k too large");
    PeerBlocks pb{};
    for (int s = 0; s < world; ++s) {
        if (!peer_local[s] || !peer_final[s]) return fail(ANN_ERR_NULL_POINTER, "ann_exchange_merge_device: NULL block pointer");
        pb.local[s] = static_cast<const unsigned char*>(peer_local[s]);
        pb.final_[s] = static_cast<unsigned char*>(peer_final[s]);
    }
    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(launch_exchange_merge(pb, world, b, k, q_begin, q_count, (cudaStream_t)stream));
    return ANN_OK;
}

int ann_peer_push_device(int32_t device, const void* d_src, void* const* peer_dst, int32_t n_dst, size_t nbytes, void* stream) {
    if (n_dst < 0 || n_dst > kMaxPeers) return fail(ANN_ERR_INVALID_ARGUMENT, "ann_peer_push_device: n_dst must be in [0, 16]");
    if (nbytes == 0 || n_dst == 0) return ANN_OK;
    if (!d_src || !peer_dst) return fail(ANN_ERR_NULL_POINTER, "ann_peer_push_device: NULL pointer");
    for (int i = 0; i < n_dst; ++i)
        if (!peer_dst[i]) return fail(ANN_ERR_NULL_POINTER, "ann_peer_push_device: NULL destination");
    if ((nbytes & 15) || ((uintptr_t)d_src & 15)) return fail(ANN_ERR_INVALID_ARGUMENT, "ann_peer_push_device: 16-byte alignment required");
    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(launch_peer_push(d_src, peer_dst, n_dst, nbytes, (cudaStream_t)stream));
    return ANN_OK;
}

size_t ann_result_block_bytes(int32_t b, int32_t k) { return (b < 0 || k < 0) ? 0 : result_block_bytes(b, k); }

int ann_set_option(ann_index* ix, const char* name, int64_t value) {
    if (!ix || !name) return fail(ANN_ERR_NULL_POINTER, "ann_set_option: NULL argument");
    std::lock_guard<std::mutex> lk(ix->mu);
    if (!strcmp(name, "path")) {
        if (value < 0 || value > 3) return fail(ANN_ERR_INVALID_ARGUMENT, "path must be 0 (auto), 1 (scan), 2 (gemm) or 3 (exact fallback)");
        ix->path_opt = (int)value;
        return ANN_OK;
    }
    if (!strcmp(name, "timing")) {
        ix->timing = value != 0;
        ix->timing_level = (int)value;
        for (double& m : ix->ms_by_label) m = 0.0;
        ix->kernel_ms_total = 0.0;
        ix->kernel_launches_timed = 0;
        ix->ev_used = 0;
        return ANN_OK;
    }
    if (!strcmp(name, "device_fallback")) {
        ix->device_fallback = value != 0;
        return ANN_OK;
    }
    if (!strcmp(name, "gemm_hit_budget")) {
        if (value < 100 || value > kGemmPoolCap / 2) return fail(ANN_ERR_INVALID_ARGUMENT, "gemm_hit_budget must be in [100, 2048]");
        ix->gemm_hit_budget = (int)value;
        return ANN_OK;
    }
    if (!strcmp(name, "gemm_small_select")) {
        ix->gemm_small_select = value ? 1 : 0;
        return ANN_OK;
    }
    if (!strcmp(name, "gemm_growth_pct")) {
        if (value != 0 && (value < 125 || value > 800)) return fail(ANN_ERR_INVALID_ARGUMENT, "gemm_growth_pct must be 0 (auto) or in [125, 800]");
        ix->gemm_growth_pct = (int)value;
        return ANN_OK;
    }
    if (!strcmp(name, "gemm_seed_rows")) {
        if (value < 0) return fail(ANN_ERR_INVALID_ARGUMENT, "gemm_seed_rows must be >= 0");
        ix->gemm_seed_rows = value;
        return ANN_OK;
    }
    if (!strcmp(name, "gemm_epi_warps")) {
        if (value != 0 && value != 8 && value != 16) return fail(ANN_ERR_INVALID_ARGUMENT, "gemm_epi_warps must be 0 (auto), 8 or 16");
        ix->gemm_epi_warps = (int)value;
        return ANN_OK;
    }
    if (!strcmp(name, "gemm_mma_warps")) {
        if (value != 1 && value != 2) return fail(ANN_ERR_INVALID_ARGUMENT, "gemm_mma_warps must be 1 or 2");
        ix->gemm_mma_warps = (int)value;
        return ANN_OK;
    }
    if (!strcmp(name, "gemm_cta_group")) {
        if (value != 1 && value != 2) return fail(ANN_ERR_INVALID_ARGUMENT, "gemm_cta_group must be 1 or 2");
        ix->gemm_cta_group = (int)value;
        return ANN_OK;
    }
    if (!strcmp(name, "coalesce_max_batch")) {   // 0 = every host call runs on its own
        if (value < 0 || value > 16384) return fail(ANN_ERR_INVALID_ARGUMENT, "coalesce_max_batch must be in [0, 16384]");
        std::lock_guard<std::mutex> clk(ix->co.mu);
        ix->co.max_batch = (int)value;
        return ANN_OK;
    }
    if (!strcmp(name, "coalesce_linger_us")) {
        if (value < 0 || value > 100000) return fail(ANN_ERR_INVALID_ARGUMENT, "coalesce_linger_us must be in [0, 100000]");
        std::lock_guard<std::mutex> clk(ix->co.mu);
        ix->co.linger_us = (int)value;
        return ANN_OK;
    }
    if (!strcmp(name, "coalesce_small_b")) {
        if (value < 1) return fail(ANN_ERR_INVALID_ARGUMENT, "coalesce_small_b must be >= 1");
        std::lock_guard<std::mutex> clk(ix->co.mu);
        ix->co.small_b = (int)value;
        return ANN_OK;
    }
    if (!strcmp(name, "gemm_min_batch")) {
        if (value < 1) return fail(ANN_ERR_INVALID_ARGUMENT, "gemm_min_batch must be >= 1");
        ix->gemm_min_batch = (int)value;
        return ANN_OK;
    }
    return fail(ANN_ERR_UNKNOWN_OPTION, std::string("unknown option: ") + name);
}

int ann_get_stat(const ann_index* ix, const char* name, int64_t* value) {
    if (!ix || !name || !value) return fail(ANN_ERR_NULL_POINTER, "ann_get_stat: NULL argument");
    if (!strcmp(name, "pending_error")) {
        // synchronise the device and surface (then clear) the sticky selector-overflow word
        ann_index* m = const_cast<ann_index*>(ix);
        std::lock_guard<std::mutex> lk(m->mu);
        CUDA_TRY(cudaSetDevice(m->device));
        CUDA_TRY(cudaDeviceSynchronize());
        *value = 0;
        return check_device_flags(m, m->stream);
    }
    if (!strncmp(name, "us_", 3)) {   // per-kernel-class totals since "timing" = 2 was set; synchronises
        static const char* const names[] = {"us_main", "us_prep", "us_compact", "us_seed_merge", "us_finalize", "us_other"};
        for (int l = 0; l < ann_index::kLblCount; ++l)
            if (!strcmp(name, names[l])) {
                ann_index* m = const_cast<ann_index*>(ix);
                std::lock_guard<std::mutex> lk(m->mu);
                CUDA_TRY(cudaSetDevice(m->device));
                CUDA_TRY(cudaDeviceSynchronize());
                harvest_timing(m);
                *value = (int64_t)(m->ms_by_label[l] * 1000.0);
                return ANN_OK;
            }
    }
    if (!strcmp(name, "kernel_us") || !strcmp(name, "kernel_launches_timed")) {
        // dominant-kernel time accumulated since timing was switched on; synchronises the device
        ann_index* m = const_cast<ann_index*>(ix);
        std::lock_guard<std::mutex> lk(m->mu);
        CUDA_TRY(cudaSetDevice(m->device));
        CUDA_TRY(cudaDeviceSynchronize());
        harvest_timing(m);
        *value = !strcmp(name, "kernel_us") ? (int64_t)(m->kernel_ms_total * 1000.0) : m->kernel_launches_timed;
        return ANN_OK;
    }
    if (!strcmp(name, "launches")) *value = ix->launches.load();
    else if (!strcmp(name, "last_path")) *value = ix->last_path;
    else if (!strcmp(name, "last_gemm_chunks")) *value = ix->last_gemm_chunks;
    else if (!strcmp(name, "exact_fallback_queries")) *value = ix->exact_fallback_queries;
    else if (!strcmp(name, "n_special")) *value = (int64_t)ix->n_special;
    else if (!strcmp(name, "row_bytes")) *value = (int64_t)ix->n * ix->pitch * 4;
    else if (!strcmp(name, "shadow_bytes")) *value = ix->shadow ? (int64_t)ix->n * ix->kp * 2 : 0;
    else if (!strcmp(name, "capacity")) *value = ix->cap;
    else if (!strcmp(name, "dim")) *value = ix->dim;
    else if (!strcmp(name, "coalesced_batches")) *value = ix->co.batches;
    else if (!strcmp(name, "coalesced_calls")) *value = ix->co.merged_calls;
    else if (!strcmp(name, "mapped_bytes")) *value = (int64_t)(ix->vm_rows.mapped + ix->vm_ids.mapped + ix->vm_norm.mapped + ix->vm_inv.mapped + ix->vm_shadow.mapped);
    else if (!strcmp(name, "max_rows")) *value = ix->max_rows;
    else if (!strcmp(name, "sm_count")) *value = ix->sm_count;
    else return fail(ANN_ERR_UNKNOWN_OPTION, std::string("unknown stat: ") + name);
    return ANN_OK;
}

}  // extern "C"
