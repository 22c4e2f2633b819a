// select_kernels.cu -- candidate-pool compaction, exact fp64 rescoring (K4), final (distance, id) ordering,
// and the cross-shard k-merge (K5).
//
//   finalize  : replaces the heap drain `priorityQueue.dequeueAll.reverse` (BruteForceIndex.scala:87-89) and the
//               per-pair `metric.distance(row, query)` (Metric.scala:89-94, 120-125, 153-158) for the few hundred
//               rows that survive the approximate scan.  The arithmetic is the oracle's, operation for operation:
//               fp64 accumulation in index order with explicit round-to-nearest mul/add (no contraction), one
//               rounding to fp32, then `1 - x` in fp32 for Cosine / InnerProduct.
//   merge     : replaces ComposedQueryable.queryWithDistance's flatten + sort + take (ShardApi.scala:77-85),
//               ordered by (Float.compare(distance), id) so that R shards give the single-shard answer bit for bit.
#include "common.cuh"
#include "kernels.h"

namespace b200ann {

namespace {

constexpr int kSelThreads = 512;
constexpr int kSortCap = 4096;    // approximate-stage sort capacity per query
constexpr int kExactCap = 2048;   // survivors + specials rescored exactly (power of two)

__device__ void bitonic_sort_entries(entry_t* a, int n2) {
    for (int k = 2; k <= n2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n2; i += blockDim.x) {
                int ixj = i ^ j;
                if (ixj > i) {
                    entry_t x = a[i], y = a[ixj];
                    bool up = ((i & k) == 0);
                    if ((x > y) == up) {
                        a[i] = y;
                        a[ixj] = x;
                    }
                }
            }
            __syncthreads();
        }
    }
}

__device__ __forceinline__ bool pair_greater(uint32_t ka, long long ia, uint32_t kb, long long ib) {
    return ka > kb || (ka == kb && ia > ib);
}

__device__ void bitonic_sort_pairs(uint32_t* key, long long* id, int n2) {
    for (int k = 2; k <= n2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n2; i += blockDim.x) {
                int ixj = i ^ j;
                if (ixj > i) {
                    uint32_t ka = key[i], kb = key[ixj];
                    long long ia = id[i], ib = id[ixj];
                    bool up = ((i & k) == 0);
                    if (pair_greater(ka, ia, kb, ib) == up) {
                        key[i] = kb;
                        key[ixj] = ka;
                        id[i] = ib;
                        id[ixj] = ia;
                    }
                }
            }
            __syncthreads();
        }
    }
}

// smallest key x with |{v <= x}| >= J over `count` published keys; whole CTA, result broadcast through smem
__device__ uint32_t cta_kth_smallest_key(const uint32_t* v, int count, int J, uint32_t* scratch) {
    if (J > count) return 0xFFFFFFFFu;
    __shared__ int cnt_s;
    uint32_t x = 0;
    for (int bit = 31; bit >= 0; --bit) {
        uint32_t trial = x | ((1u << bit) - 1u);
        if (threadIdx.x == 0) cnt_s = 0;
        __syncthreads();
        int c = 0;
        for (int i = threadIdx.x; i < count; i += blockDim.x) c += (v[i] <= trial) ? 1 : 0;
        if (c) atomicAdd(&cnt_s, c);
        __syncthreads();
        if (cnt_s < J) x |= (1u << bit);
        __syncthreads();
    }
    (void)scratch;
    return x;
}

// exact distance, operation for operation the oracle's distance_f64 (oracle/oracle.c)
__device__ float exact_distance(int metric, const float* __restrict__ a, const float* __restrict__ b, int d, int l2_squared) {
    if (metric == kMetricL2) {
        double acc = 0.0;
        for (int i = 0; i < d; ++i) {
            double diff = __dsub_rn((double)a[i], (double)b[i]);
            double sq = __dmul_rn(diff, diff);
            acc = __dadd_rn(acc, sq);
        }
        return __double2float_rn(l2_squared ? acc : __dsqrt_rn(acc));
    }
    double dot = 0.0;
    for (int i = 0; i < d; ++i) dot = __dadd_rn(dot, __dmul_rn((double)a[i], (double)b[i]));
    if (metric == kMetricIP) return __fsub_rn(1.0f, __double2float_rn(dot));
    double na = 0.0, nb = 0.0;
    for (int i = 0; i < d; ++i) na = __dadd_rn(na, __dmul_rn((double)a[i], (double)a[i]));
    for (int i = 0; i < d; ++i) nb = __dadd_rn(nb, __dmul_rn((double)b[i], (double)b[i]));
    double cs = __ddiv_rn(dot, __dmul_rn(__dsqrt_rn(na), __dsqrt_rn(nb)));
    return __fsub_rn(1.0f, __double2float_rn(cs));
}

// Shared front end of compaction and finalize: filter the pool by the current threshold, sort, find the
// survivors (everything within the margin of the k-th best).  Returns the survivor count in *n_surv and the
// tightened threshold in *tau_out; entries are left sorted in `buf`.  Returns false on overflow.
__device__ bool select_survivors(const SelectParams& p, int q, entry_t* buf, int* n_surv, float* tau_out) {
    __shared__ int n_s;
    __shared__ float tau_s;
    __shared__ int ok_s;
    QueryState* qs = p.qstate + q;
    const float eps_abs = qs->eps_abs, eps_rel = qs->eps_rel;
    const int pool_n = min((int)qs->pool_count, p.pool_cap);
    float tau = INFINITY;
    uint32_t tk = qs->tau_key;
    if (tk < 0xFF800000u) tau = float_from_order_key(tk);
    if (p.pub_keys) {
        uint32_t gk = cta_kth_smallest_key(p.pub_keys + (size_t)q * p.pub_stride, p.pub_count, p.j_pub, nullptr);
        if (gk < 0xFF800000u) tau = fminf(tau, widen(float_from_order_key(gk), eps_abs, eps_rel));
    }
    if (threadIdx.x == 0) {
        n_s = 0;
        ok_s = 1;
    }
    __syncthreads();
    const entry_t* pool = p.pool + (size_t)q * p.pool_cap;
    for (int i = threadIdx.x; i < pool_n; i += blockDim.x) {
        entry_t e = pool[i];
        if (entry_g(e) <= tau) {
            int slot = atomicAdd(&n_s, 1);
            if (slot < kSortCap) buf[slot] = e;
        }
    }
    __syncthreads();
    int n = n_s;
    if (n > kSortCap) {
        if (threadIdx.x == 0) atomicOr(&qs->flags, kFlagSurvivorOverflow);
        return false;
    }
    int n2 = 2;
    while (n2 < n) n2 <<= 1;
    for (int i = n + threadIdx.x; i < n2; i += blockDim.x) buf[i] = kEntryPad;
    __syncthreads();
    bitonic_sort_entries(buf, n2);
    if (threadIdx.x == 0) {
        float t = tau;
        if (n >= p.k && p.k > 0) t = fminf(t, widen(entry_g(buf[p.k - 1]), eps_abs, eps_rel));
        tau_s = t;
    }
    __syncthreads();
    tau = tau_s;
    // survivors = sorted prefix with g <= tau : binary search by one thread
    if (threadIdx.x == 0) {
        int lo = 0, hi = n;
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (entry_g(buf[mid]) <= tau) lo = mid + 1;
            else hi = mid;
        }
        n_s = lo;
    }
    __syncthreads();
    *n_surv = n_s;
    *tau_out = tau;
    return ok_s != 0;
}

}  // namespace

// ---- approx-only compaction between GEMM chunks: pool <- survivors, tau <- k-th best + margin --------------
__global__ void __launch_bounds__(kSelThreads) compact_pool_kernel(SelectParams p) {
    extern __shared__ __align__(16) unsigned char sm[];
    entry_t* buf = reinterpret_cast<entry_t*>(sm);
    const int q = blockIdx.x;
    QueryState* qs = p.qstate + q;
    if (qs->pool_count > (uint32_t)p.pool_cap) {
        if (threadIdx.x == 0) atomicOr(&qs->flags, kFlagPoolOverflow);
        return;
    }
    int n_surv;
    float tau;
    if (!select_survivors(p, q, buf, &n_surv, &tau)) return;
    entry_t* pool = p.pool + (size_t)q * p.pool_cap;
    for (int i = threadIdx.x; i < n_surv; i += blockDim.x) pool[i] = buf[i];
    if (threadIdx.x == 0) {
        qs->pool_count = n_surv;
        qs->tau_key = float_order_key(tau);
    }
}

// ---- finalize: survivors + specials -> exact distances -> (distance, id) order -> outputs -------------------
__global__ void __launch_bounds__(kSelThreads) finalize_kernel(SelectParams p) {
    extern __shared__ __align__(16) unsigned char sm[];
    entry_t* buf = reinterpret_cast<entry_t*>(sm);                       // kSortCap entries
    long long* cid = reinterpret_cast<long long*>(sm + kSortCap * 8);      // kExactCap ids
    uint32_t* ckey = reinterpret_cast<uint32_t*>(sm + kSortCap * 8 + kExactCap * 8);  // kExactCap keys
    const int q = blockIdx.x;
    QueryState* qs = p.qstate + q;
    int64_t* oid = p.out_ids + (size_t)q * p.k_out;
    float* od = p.out_dist + (size_t)q * p.k_out;

    auto fail_fill = [&]() {
        for (int j = threadIdx.x; j < p.k_out; j += blockDim.x) {
            oid[j] = -1;
            od[j] = INFINITY;
        }
        if (threadIdx.x == 0 && p.out_count) p.out_count[q] = 0;
    };

    if (qs->pool_count > (uint32_t)p.pool_cap) {
        if (threadIdx.x == 0) atomicOr(&qs->flags, kFlagPoolOverflow);
        fail_fill();
        return;
    }
    int n_surv;
    float tau;
    if (!select_survivors(p, q, buf, &n_surv, &tau)) {
        fail_fill();
        return;
    }
    const int n_spec = min((int)qs->special_count, kSpecialCap);
    const int n_cand = n_surv + n_spec;
    if (n_cand > kExactCap || (qs->flags & (kFlagPoolOverflow | kFlagSpecialOverflow))) {
        if (threadIdx.x == 0 && n_cand > kExactCap) atomicOr(&qs->flags, kFlagSurvivorOverflow);
        fail_fill();
        return;
    }
    const float* qv = p.queries + (size_t)q * p.q_pitch;
    for (int i = threadIdx.x; i < n_cand; i += blockDim.x) {
        uint32_t row = i < n_surv ? entry_row(buf[i]) : p.special_rows[(size_t)q * kSpecialCap + (i - n_surv)];
        float dist = exact_distance(p.metric, p.rows + (size_t)row * p.pitch, qv, p.dim, p.l2_squared);
        ckey[i] = float_order_key(dist);
        cid[i] = p.ids[row];
    }
    int n2 = 2;
    while (n2 < n_cand) n2 <<= 1;
    for (int i = n_cand + threadIdx.x; i < n2; i += blockDim.x) {
        ckey[i] = 0xFFFFFFFFu;
        cid[i] = 0x7FFFFFFFFFFFFFFFll;
    }
    __syncthreads();
    bitonic_sort_pairs(ckey, cid, n2);
    const int cnt = min(p.k, n_cand);
    for (int j = threadIdx.x; j < p.k_out; j += blockDim.x) {
        if (j < cnt) {
            oid[j] = cid[j];
            od[j] = float_from_order_key(ckey[j]);
        } else {
            oid[j] = -1;
            od[j] = INFINITY;
        }
    }
    if (threadIdx.x == 0 && p.out_count) p.out_count[q] = cnt;
}

__global__ void fill_empty_kernel(int64_t* out_ids, float* out_dist, int32_t* out_count, int b, int k_out) {
    long long n = (long long)b * k_out;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        out_ids[i] = -1;
        out_dist[i] = INFINITY;
    }
    if (out_count)
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < b; i += gridDim.x * blockDim.x) out_count[i] = 0;
}

// ---- K5: merge `shards` sorted lists of k per query ----------------------------------------------------------
__global__ void __launch_bounds__(kSelThreads) merge_topk_kernel(const int64_t* ids, const float* dist, const int32_t* count,
                                                                 int shards, int b, int k, int n2, int64_t* out_ids,
                                                                 float* out_dist, int32_t* out_count) {
    extern __shared__ __align__(16) unsigned char sm[];
    long long* cid = reinterpret_cast<long long*>(sm);
    uint32_t* ckey = reinterpret_cast<uint32_t*>(sm + (size_t)n2 * 8);
    const int q = blockIdx.x;
    __shared__ int total_s;
    if (threadIdx.x == 0) {
        int t = 0;
        for (int s = 0; s < shards; ++s) t += min(max(count[(size_t)s * b + q], 0), k);
        total_s = t;
    }
    for (int i = threadIdx.x; i < n2; i += blockDim.x) {
        int s = i / k, j = i - s * k;
        bool valid = s < shards && j < min(max(count[(size_t)s * b + q], 0), k);
        if (valid) {
            size_t off = ((size_t)s * b + q) * k + j;
            ckey[i] = float_order_key(dist[off]);
            cid[i] = ids[off];
        } else {
            ckey[i] = 0xFFFFFFFFu;
            cid[i] = 0x7FFFFFFFFFFFFFFFll;
        }
    }
    __syncthreads();
    // NaN distances share the pad key; order them before pads by id (pads carry INT64_MAX).
    bitonic_sort_pairs(ckey, cid, n2);
    const int cnt = min(k, total_s);
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
        size_t off = (size_t)q * k + j;
        if (j < cnt) {
            out_ids[off] = cid[j];
            out_dist[off] = float_from_order_key(ckey[j]);
        } else {
            out_ids[off] = -1;
            out_dist[off] = INFINITY;
        }
    }
    if (threadIdx.x == 0 && out_count) out_count[q] = cnt;
}

cudaError_t launch_compact_pool(const SelectParams& p, int b, cudaStream_t stream) {
    size_t smem = (size_t)kSortCap * 8;
    cudaError_t e = cudaFuncSetAttribute(compact_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    compact_pool_kernel<<<b, kSelThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_finalize(const SelectParams& p, int b, cudaStream_t stream) {
    size_t smem = (size_t)kSortCap * 8 + (size_t)kExactCap * 12;
    cudaError_t e = cudaFuncSetAttribute(finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    finalize_kernel<<<b, kSelThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_fill_empty(int64_t* out_ids, float* out_dist, int32_t* out_count, int b, int k_out, cudaStream_t stream) {
    long long n = (long long)b * k_out;
    int grid = (int)((n + 255) / 256);
    if (grid < 1) grid = 1;
    if (grid > 1024) grid = 1024;
    fill_empty_kernel<<<grid, 256, 0, stream>>>(out_ids, out_dist, out_count, b, k_out);
    return cudaGetLastError();
}

cudaError_t launch_merge(const int64_t* ids, const float* dist, const int32_t* count, int shards, int b, int k,
                         int64_t* out_ids, float* out_dist, int32_t* out_count, cudaStream_t stream) {
    int n2 = 2;
    while (n2 < shards * k) n2 <<= 1;
    size_t smem = (size_t)n2 * 12;
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(merge_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    merge_topk_kernel<<<b, kSelThreads, smem, stream>>>(ids, dist, count, shards, b, k, n2, out_ids, out_dist, out_count);
    return cudaGetLastError();
}

}  // namespace b200ann
