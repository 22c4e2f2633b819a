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
#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace b200ann {

namespace {

constexpr int kSelThreads = 256;
constexpr int kCompactThreads = 128;   // compaction: ~600 entries per query, barrier-bound.  Smaller CTAs put more queries in flight
                                       // per SM: 0.239 -> 0.190 ms per 4096-query batch (5 launches); 64 threads are back at 0.237,
                                       // and 128 threads for finalize are slower (0.313 -> 0.337: fewer candidates rescored per pass)
constexpr int kSortCap = 4096;    // default approximate-stage capacity per query (SelectParams::sort_cap overrides)
constexpr int kRankSortMax = 384; // up to this many exact candidates are ordered by rank counting, more by the bitonic network
constexpr int kExactCap = 2048;   // default survivors + specials rescored exactly (power of two; SelectParams::exact_cap)
__host__ __device__ inline int sort_cap_of(const SelectParams& p) { return p.sort_cap > 0 ? p.sort_cap : kSortCap; }
__host__ __device__ inline int exact_cap_of(const SelectParams& p) { return p.exact_cap > 0 ? p.exact_cap : kExactCap; }

__device__ __forceinline__ bool pair_greater(uint32_t ka, long long ia, uint32_t kb, long long ib) {
    return ka > kb || (ka == kb && ia > ib);
}

// One side of the margin, rounded up: g + eps.  A witness row's exact badness is at most its approximate g + eps; a
// candidate's exact badness is at least its approximate g - eps.  widen() is both sides at once.
__device__ __forceinline__ float widen_half_up(float g, float eps_abs, float eps_rel) {
    return __fadd_ru(g, __fadd_ru(eps_abs, __fmul_ru(eps_rel, fabsf(g))));
}

// Slot allocation for a warp: the lanes with `take` get consecutive slots from ONE atomic on `counter` (all 32 lanes must
// call it).  The plain one-atomic-per-thread form serialises a few hundred shared-memory atomics on a single address per CTA.
__device__ __forceinline__ int warp_slot(bool take, int* counter) {
    const unsigned m = __ballot_sync(0xFFFFFFFFu, take);
    if (m == 0) return 0;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    return base + __popc(m & ((1u << lane) - 1u));
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

// k-th smallest 32-bit key among the n entries of `buf` (k >= 1, n >= k): 4 MSB-first passes of 8 bits, each a
// 256-bin shared-memory histogram over the entries still matching the prefix and a one-warp scan that picks the bin.
__device__ uint32_t kth_key_radix(const entry_t* buf, int n, int k, uint32_t* hist) {
    __shared__ uint32_t prefix_s, krem_s;
    if (threadIdx.x == 0) {
        prefix_s = 0;
        krem_s = (uint32_t)k;
    }
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        const uint32_t prefix = prefix_s;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            // plain shared-memory atomics: aggregating equal bins with __match_any_sync first was measured SLOWER (compaction
            // 42 -> 60 us per 4096-query launch), the match itself costs more than the conflicts it removes
            const uint32_t key = (uint32_t)(buf[i] >> 32);
            if (pass == 0 || (key >> (shift + 8)) == (prefix >> (shift + 8))) atomicAdd(&hist[(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            const int lane = threadIdx.x;
            uint32_t c[8], sum = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                c[j] = hist[lane * 8 + j];
                sum += c[j];
            }
            uint32_t incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= o) incl += t;
            }
            const uint32_t excl = incl - sum, krem = krem_s;
            // exactly one lane has excl < krem <= incl
            if (excl < krem && krem <= incl) {
                uint32_t cum = excl;
                int bin = lane * 8 + 7;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (cum + c[j] >= krem) {
                        bin = lane * 8 + j;
                        break;
                    }
                    cum += c[j];
                }
                prefix_s = prefix | ((uint32_t)bin << shift);
                krem_s = krem - cum;
            }
        }
        __syncthreads();
    }
    return prefix_s;
}

// Shared front end of compaction and finalize: load the pool entries that pass the current threshold into `buf`
// (unsorted), find the k-th best approximate badness with a radix select, tighten the threshold to k-th + margin.
// Everything with g <= *tau_out may still belong to the exact top-k.  Returns false on overflow.
__device__ bool load_and_threshold(const SelectParams& p, int q, entry_t* buf, uint32_t* hist, int* n_out, float* tau_out,
                                   uint32_t* kth_key_out = nullptr) {
    __shared__ int n_s;
    const int sort_cap = sort_cap_of(p);
    QueryState* qs = p.qstate + q;
    const float eps_abs = qs->eps_abs, eps_rel = qs->eps_rel;
    const int pool_n = p.seed_count > 0 ? min(p.seed_count, p.pool_cap) : min((int)qs->pool_count, p.pool_cap);
    float tau = INFINITY;
    uint32_t tk = qs->tau_key;
    if (tk < 0xFF800000u) tau = float_from_order_key(tk);
    if (p.pub_keys) {
        uint32_t gk = cta_kth_smallest_key(p.pub_keys + (size_t)q * p.pub_stride, p.pub_count, p.j_pub, nullptr);
        if (gk < 0xFF800000u) tau = fminf(tau, widen(float_from_order_key(gk), eps_abs, eps_rel));
    }
    if (p.peer_world > 0 && p.k > 0) {
        // k-th smallest of the union of every shard's published bounds (P2P loads, k*4 bytes per shard): at least k distinct
        // rows of the whole index have exact badness <= U, so nothing above U + eps (this shard's margin) can be in the answer
        const int total = p.peer_world * p.k;     // <= sort_cap (checked by the launcher)
        for (int i = threadIdx.x; i < total; i += blockDim.x) {
            const int s = i / p.k, j = i - s * p.k;
            buf[i] = (entry_t)__ldcg(p.peer_keys.keys[s] + (size_t)q * p.k + j) << 32;
        }
        __syncthreads();
        const uint32_t u = kth_key_radix(buf, total, p.k, hist);
        if (u < 0xFF800000u) tau = fminf(tau, widen_half_up(float_from_order_key(u), eps_abs, eps_rel));
        __syncthreads();
    }
    if (threadIdx.x == 0) n_s = 0;
    __syncthreads();
    const entry_t* pool = p.pool + (size_t)q * p.pool_cap;
    for (int i0 = 0; i0 < pool_n; i0 += blockDim.x) {     // whole warps iterate together: one shared-memory atomic per warp
        const int i = i0 + threadIdx.x;
        entry_t e = 0;
        bool keep = false;
        if (i < pool_n) {
            e = pool[i];
            const float ge = entry_g(e);
            keep = ge <= tau && ge < kSpecialG;   // special rows scored -1e38 by the tensor-core filter are handled separately
        }
        const int slot = warp_slot(keep, &n_s);
        if (keep && slot < sort_cap) buf[slot] = e;
    }
    __syncthreads();
    const int n = n_s;
    if (n > sort_cap) {
        if (threadIdx.x == 0) atomicOr(&qs->flags, kFlagSurvivorOverflow);
        return false;
    }
    if (kth_key_out) *kth_key_out = 0xFFFFFFFFu;
    // With the shards' round-2 bounds in hand (this shard's own k best among them) the cut above is already at least as tight
    // as this shard's own k-th best + margin: the second selection would be four more histogram passes for nothing.
    if (n >= p.k && p.k > 0 && !(p.peer_world > 0 && !kth_key_out)) {
        const uint32_t kk = kth_key_radix(buf, n, p.k, hist);
        tau = fminf(tau, widen(float_from_order_key(kk), eps_abs, eps_rel));
        if (kth_key_out) *kth_key_out = kk;
    }
    *n_out = n;
    *tau_out = tau;
    return true;
}

}  // namespace

// ---- approx-only compaction between GEMM chunks: pool <- survivors, tau <- k-th best + margin --------------
__global__ void __launch_bounds__(kCompactThreads, 8) compact_pool_kernel(SelectParams p) {
    extern __shared__ __align__(16) unsigned char sm[];
    entry_t* buf = reinterpret_cast<entry_t*>(sm);
    uint32_t* pubk = reinterpret_cast<uint32_t*>(sm + (size_t)sort_cap_of(p) * 8);   // [k] keys to publish (only when publishing)
    __shared__ uint32_t hist[256];
    __shared__ int n_keep, n_pub;
    const int q = blockIdx.x;
    QueryState* qs = p.qstate + q;
    const bool publishing = p.seed_keys_out != nullptr || p.n_push > 0;
    // the staged keys go to the local array, or (push mode) into this shard's block of EVERY peer's receive buffer
    auto publish = [&]() {
        __syncthreads();
        if (p.n_push > 0) {
            for (int d = 0; d < p.n_push; ++d) {
                uint32_t* out = p.push_keys[d] + (size_t)q * p.k;
                for (int j = threadIdx.x; j < p.k; j += blockDim.x) out[j] = pubk[j];
            }
        } else {
            uint32_t* out = p.seed_keys_out + (size_t)q * p.k;
            for (int j = threadIdx.x; j < p.k; j += blockDim.x) out[j] = pubk[j];
        }
    };
    auto publish_one = [&](uint32_t key) {   // sliced seeding: one bound per query, to every peer's [b] array
        if (threadIdx.x < max(p.n_push, 1)) {
            uint32_t* out = p.n_push > 0 ? p.push_keys[threadIdx.x] : p.seed_keys_out;
            out[q] = key;
        }
    };
    auto publish_nothing = [&]() {   // a flagged query publishes "no bound" (never stale keys of an earlier batch)
        if (!publishing) return;
        if (p.pub_single) {
            publish_one(0xFFFFFFFFu);
            return;
        }
        for (int j = threadIdx.x; j < p.k; j += blockDim.x) pubk[j] = 0xFFFFFFFFu;
        publish();
    };
    if (p.seed_count == 0 && qs->pool_count > (uint32_t)p.pool_cap) {
        if (threadIdx.x == 0) atomicOr(&qs->flags, kFlagPoolOverflow);
        publish_nothing();
        return;
    }
    int n;
    float tau;
    uint32_t kk;
    if (!load_and_threshold(p, q, buf, hist, &n, &tau, &kk)) {
        publish_nothing();
        return;
    }
    // The k best keys of `buf`, each widened to an upper bound on the exact badness of the row (or 32-row group: distinct
    // groups => distinct rows) that produced it.  Order does not matter to the consumer; entries equal to the k-th key all
    // publish the same value, so the tail is simply filled with it.  Fewer than k entries: every one is a witness, the rest
    // is "no bound".
    auto stage_best = [&]() {
        const float ea = qs->eps_abs, er = qs->eps_rel;
        if (threadIdx.x == 0) n_pub = 0;
        __syncthreads();
        if (kk != 0xFFFFFFFFu) {
            for (int i0 = 0; i0 < n; i0 += blockDim.x) {
                const int i = i0 + threadIdx.x;
                const uint32_t key = i < n ? (uint32_t)(buf[i] >> 32) : 0xFFFFFFFFu;
                const bool take = i < n && key < kk;                 // fewer than k entries are strictly below the k-th key
                const int slot = warp_slot(take, &n_pub);
                if (take) pubk[slot] = float_order_key(widen_half_up(float_from_order_key(key), ea, er));
            }
            __syncthreads();
            const uint32_t fill = float_order_key(widen_half_up(float_from_order_key(kk), ea, er));
            for (int j = n_pub + threadIdx.x; j < p.k; j += blockDim.x) pubk[j] = fill;
        } else {
            for (int i = threadIdx.x; i < min(n, p.k); i += blockDim.x) pubk[i] = float_order_key(widen_half_up(entry_g(buf[i]), ea, er));
            for (int j = n + threadIdx.x; j < p.k; j += blockDim.x) pubk[j] = 0xFFFFFFFFu;
        }
    };
    if (p.seed_count > 0) {   // seed entries carry no row: keep only the threshold (K5c publish: the k best group maxima)
        if (publishing && p.pub_single) {
            // the k-th best group maximum, widened to an upper bound on the exact badness of the k-th best of k real rows
            publish_one(kk != 0xFFFFFFFFu ? float_order_key(widen_half_up(float_from_order_key(kk), qs->eps_abs, qs->eps_rel)) : 0xFFFFFFFFu);
        } else if (publishing) {
            if (kk != 0xFFFFFFFFu) stage_best();
            else
                for (int j = threadIdx.x; j < p.k; j += blockDim.x) pubk[j] = 0xFFFFFFFFu;   // fewer than k groups: no bound
            publish();
        }
        if (threadIdx.x == 0) {
            qs->pool_count = 0;
            qs->tau_key = float_order_key(tau);
        }
        return;
    }
    if (threadIdx.x == 0) n_keep = 0;
    __syncthreads();
    entry_t* pool = p.pool + (size_t)q * p.pool_cap;
    for (int i0 = 0; i0 < n; i0 += blockDim.x) {
        const int i = i0 + threadIdx.x;
        const entry_t e = i < n ? buf[i] : 0;
        const bool keep = i < n && entry_g(e) <= tau;
        const int slot = warp_slot(keep, &n_keep);
        if (keep) pool[slot] = e;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        qs->pool_count = n_keep;
        qs->tau_key = float_order_key(tau);
    }
    if (publishing) {
        // Second cross-shard round (sharded query, after the last chunk): publish this shard's k best approximate keys.  The
        // k-th smallest over all shards' arrays bounds the GLOBAL k-th best, which lets every shard rescore only its share of
        // the ~2k global survivors instead of its own ~2k (the local k-th of 1/R of the rows is far looser than the global one).
        stage_best();
        publish();
    }
}

// ---- finalize: survivors + specials -> exact distances -> (distance, id) order -> outputs -------------------
// (4 CTAs per SM on purpose: 5 / 6 per SM -- 48 / 40 registers -- measured 0.311 -> 0.326 / 0.351 ms per 4096-query batch; the
// kernel waits on scattered 800-byte row reads, 3.5 TB/s of DRAM traffic, and more CTAs only add contention)
template <int METRIC, bool ACC32>
__global__ void __launch_bounds__(kSelThreads, 4) finalize_kernel(SelectParams p) {
    extern __shared__ __align__(16) unsigned char sm[];
    const int sort_cap = sort_cap_of(p), exact_cap = exact_cap_of(p);
    entry_t* buf = reinterpret_cast<entry_t*>(sm);                       // sort_cap entries
    long long* cid = reinterpret_cast<long long*>(sm + (size_t)sort_cap * 8);      // exact_cap ids
    uint32_t* ckey = reinterpret_cast<uint32_t*>(sm + (size_t)sort_cap * 8 + (size_t)exact_cap * 8);  // exact_cap keys
    __shared__ uint32_t hist[256];
    __shared__ int n_cand_s;
    const int q = blockIdx.x;
    QueryState* qs = p.qstate + q;
    int64_t* oid = p.out_ids + (size_t)q * p.k_out;
    float* od = p.out_dist + (size_t)q * p.k_out;

    auto fail_fill = [&]() {
        for (int j = threadIdx.x; j < p.k_out; j += blockDim.x) {
            oid[j] = -1;
            od[j] = INFINITY;
        }
        if (threadIdx.x == 0 && p.out_count) p.out_count[q] = -1;   // flagged: the row is INVALID (not an empty list)
    };

    if (qs->pool_count > (uint32_t)p.pool_cap) {
        if (threadIdx.x == 0) atomicOr(&qs->flags, kFlagPoolOverflow);
        fail_fill();
        return;
    }
    int n;
    float tau;
    if (!load_and_threshold(p, q, buf, hist, &n, &tau)) {
        fail_fill();
        return;
    }
    if (qs->flags & (kFlagPoolOverflow | kFlagSpecialOverflow)) {
        fail_fill();
        return;
    }
    // Exact rescoring of every survivor (g <= tau) and every special row; unordered, the final sort orders them.
    // The candidate rows are scattered over the whole matrix, so every chain is a sequence of dependent DRAM misses.
    const uint32_t* spec_rows = p.global_special_rows ? p.global_special_rows : p.special_rows + (size_t)q * kSpecialCap;
    const int n_spec = p.global_special_rows ? min(p.global_special_count, kSpecialCap) : min((int)qs->special_count, kSpecialCap);
    if (threadIdx.x == 0) n_cand_s = 0;
    __syncthreads();
    uint32_t* crow = ckey;   // row indices first, replaced by the distance keys
    for (int i0 = 0; i0 < n + n_spec; i0 += blockDim.x) {
        const int i = i0 + threadIdx.x;
        uint32_t row = 0;
        bool keep = false;
        if (i < n) {
            const entry_t e = buf[i];
            keep = entry_g(e) <= tau;
            row = entry_row(e);
        } else if (i < n + n_spec) {
            keep = true;
            row = spec_rows[i - n];
        }
        const int slot = warp_slot(keep, &n_cand_s);
        if (keep && slot < exact_cap) crow[slot] = row;
    }
    __syncthreads();
    const int n_cand = n_cand_s;
    if (n_cand > exact_cap) {
        if (threadIdx.x == 0) atomicOr(&qs->flags, kFlagSurvivorOverflow);
        fail_fill();
        return;
    }
    {
        // The candidate rows are scattered over the whole matrix and every rescoring thread walks its row front to back, so
        // without help each 64-byte step of that walk is a fresh DRAM miss (13 dependent misses per 800-byte row).  Ask for
        // every line of every candidate row (and its id) up front: all the misses of the CTA are then in flight together,
        // and the walks below run out of L2.
        const int row_bytes = p.dim * 4;
        const int lines = (row_bytes + 127) / 128 + 1;          // rows are not 128-byte aligned: one more line may be touched
        for (int i = threadIdx.x; i < n_cand * lines; i += blockDim.x) {
            const int c = i / lines, j = i - c * lines;
            const char* base = reinterpret_cast<const char*>(p.rows + (size_t)crow[c] * p.pitch);
            const char* at = base + min(j * 128, row_bytes - 4);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(at));
            if (j == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.ids + crow[c]));
        }
        // the query as doubles (and its squared norm) once per CTA, in the shared memory the approximate entries occupied
        double* q64 = reinterpret_cast<double*>(buf);                                  // <= 8 KB (dim <= 1024)
        uint32_t* okey = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(buf) + 8192);   // exact_cap keys (sort_cap * 8 >= 8192 + exact_cap * 4)
        __shared__ double nb_s;
        __shared__ float nbf_s;
        float* q32 = reinterpret_cast<float*>(buf);      // ANN_FLAG_ACCUM_F32: the query stays fp32 (same region)
        const float* qv = p.queries + (size_t)q * p.q_pitch;
        for (int i = threadIdx.x; i < p.dim; i += blockDim.x) {
            if (ACC32) q32[i] = qv[i];
            else q64[i] = (double)qv[i];
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            if (ACC32) nbf_s = (METRIC == kMetricCosine) ? exact_query_norm2_f32(q32, p.dim) : 0.f;
            else nb_s = (METRIC == kMetricCosine) ? exact_query_norm2(q64, p.dim) : 0.0;
        }
        __syncthreads();
        const ExactQuery eq{q64, ACC32 ? 0.0 : nb_s};
        const float nbf = ACC32 ? nbf_s : 0.f;
        if (ACC32) {
            // fp32 convention: the value IS a sequentially rounded fp32 chain, nothing shorter can stand in for it
            for (int c = threadIdx.x; c < n_cand; c += blockDim.x) {
                const uint32_t row = crow[c];
                const float* rp = p.rows + (size_t)row * p.pitch;
                cid[c] = p.ids[row];
                okey[c] = float_order_key(exact_distance_rows_f32<true, METRIC>(METRIC, rp, q32, nbf, p.dim, p.l2_squared));
            }
        } else {
            // fp64 convention, 8 lanes per candidate.  The oracle's value is f = fl32(S_seq) with S_seq the sequentially
            // rounded fp64 sum -- a 200-long dependent chain per row.  ANY summation order gives a sum within
            // gamma_(d-1) * sum|t_i| of the real sum (the terms t_i themselves are computed identically: products of floats
            // are exact in fp64, (a-b)^2 is rounded the same way elementwise), so a lane-parallel sum S_par pins S_seq into
            // [S_par - E, S_par + E], E = 2 d 2^-53 sum|t_i|.  fl32 (and sqrt, and the quotient of Cosine) are monotone, so
            // when both ends of the interval round to the SAME float, that float is fl32(S_seq) -- bit for bit, with a
            // 25-long chain and coalesced 128-byte row reads.  Otherwise (about 1 candidate in 10^5) the candidate is
            // rescored by the sequential chain below.  Exactness is never a matter of probability.
            uint32_t* redo = okey + exact_cap;            // candidates whose interval straddles a rounding boundary
            __shared__ int n_redo_s;
            if (threadIdx.x == 0) n_redo_s = 0;
            for (int i = p.dim + threadIdx.x; i < p.pitch; i += blockDim.x) q64[i] = 0.0;   // pad columns (rows hold zeros there)
            __syncthreads();
            const int sub = threadIdx.x & 7, n4 = p.pitch >> 2;
            const double du = 2.1 * (double)(p.dim + 4) * 1.1102230246251565e-16;             // 2.1 * (d + 4) * 2^-53
            const double qn = (double)qs->qnorm;                                               // |b| rounded up
            for (int c0 = 0; c0 < n_cand; c0 += blockDim.x >> 3) {
                const int c = c0 + (threadIdx.x >> 3);
                const bool live = c < n_cand;
                const uint32_t row = live ? crow[c] : 0u;
                const float4* a4 = reinterpret_cast<const float4*>(p.rows + (size_t)row * p.pitch);
                double sx = 0.0, sy = 0.0, sz = 0.0, sw = 0.0, ax = 0.0, ay = 0.0;
                if (live) {
                    for (int j = sub; j < n4; j += 8) {
                        const float4 v = __ldg(a4 + j);
                        const double* bq = q64 + (j << 2);
                        if (METRIC == kMetricL2) {
                            const double d0 = __dsub_rn((double)v.x, bq[0]), d1 = __dsub_rn((double)v.y, bq[1]);
                            const double d2 = __dsub_rn((double)v.z, bq[2]), d3 = __dsub_rn((double)v.w, bq[3]);
                            sx = __dadd_rn(sx, __dmul_rn(d0, d0));
                            sy = __dadd_rn(sy, __dmul_rn(d1, d1));
                            sz = __dadd_rn(sz, __dmul_rn(d2, d2));
                            sw = __dadd_rn(sw, __dmul_rn(d3, d3));
                        } else {
                            sx = __fma_rn((double)v.x, bq[0], sx);
                            sy = __fma_rn((double)v.y, bq[1], sy);
                            sz = __fma_rn((double)v.z, bq[2], sz);
                            sw = __fma_rn((double)v.w, bq[3], sw);
                            if (METRIC == kMetricCosine) {
                                ax = __fma_rn((double)v.x, (double)v.x, ax);
                                ay = __fma_rn((double)v.y, (double)v.y, ay);
                                ax = __fma_rn((double)v.z, (double)v.z, ax);
                                ay = __fma_rn((double)v.w, (double)v.w, ay);
                            }
                        }
                    }
                }
                double S = (sx + sy) + (sz + sw), A = ax + ay;
#pragma unroll
                for (int o = 1; o < 8; o <<= 1) {
                    S += __shfl_xor_sync(0xFFFFFFFFu, S, o);
                    if (METRIC == kMetricCosine) A += __shfl_xor_sync(0xFFFFFFFFu, A, o);
                }
                if (live && sub == 0) {
                    const double an = (double)p.row_norm[row];                                  // |a| rounded up
                    float f_lo, f_hi;
                    if (METRIC == kMetricL2) {
                        const double E = du * S;
                        const double lo = fmax(__dsub_rd(S, E), 0.0), hi = __dadd_ru(S, E);
                        f_lo = __double2float_rn(p.l2_squared ? lo : __dsqrt_rn(lo));
                        f_hi = __double2float_rn(p.l2_squared ? hi : __dsqrt_rn(hi));
                    } else if (METRIC == kMetricIP) {
                        const double E = du * an * qn;
                        f_lo = __double2float_rn(__dsub_rd(S, E));
                        f_hi = __double2float_rn(__dadd_ru(S, E));
                    } else {
                        const double E = du * an * qn, Ea = du * an * an;
                        const double d_lo = __dsub_rd(S, E), d_hi = __dadd_ru(S, E);
                        const double a_lo = fmax(__dsub_rd(A, Ea), 0.0), a_hi = __dadd_ru(A, Ea);
                        const double sqb = __dsqrt_rn(nb_s);
                        const double den_lo = __dmul_rn(__dsqrt_rn(a_lo), sqb), den_hi = __dmul_rn(__dsqrt_rn(a_hi), sqb);
                        // the quotient is monotone in each argument: its extremes over the box are at the four corners
                        const float c0f = __double2float_rn(__ddiv_rn(d_lo, den_lo)), c1f = __double2float_rn(__ddiv_rn(d_lo, den_hi));
                        const float c2f = __double2float_rn(__ddiv_rn(d_hi, den_lo)), c3f = __double2float_rn(__ddiv_rn(d_hi, den_hi));
                        f_lo = fminf(fminf(c0f, c1f), fminf(c2f, c3f));
                        f_hi = fmaxf(fmaxf(c0f, c1f), fmaxf(c2f, c3f));
                        if (!(c0f == c0f && c1f == c1f && c2f == c2f && c3f == c3f)) f_lo = NAN;   // fminf / fmaxf drop NaN
                    }
                    // certain only if both ends are the same finite float (NaN, +-inf and overflowing bounds go sequential)
                    if (f_lo == f_hi && fabsf(f_lo) <= 3.4028234e38f && (METRIC != kMetricL2 || S == S)) {
                        const float dist = METRIC == kMetricL2 ? f_lo : __fsub_rn(1.0f, f_lo);
                        okey[c] = float_order_key(dist);
                    } else {
                        redo[atomicAdd(&n_redo_s, 1)] = (uint32_t)c;
                    }
                    cid[c] = p.ids[row];
                }
            }
            __syncthreads();
            for (int i = threadIdx.x; i < n_redo_s; i += blockDim.x) {
                const int c = (int)redo[i];
                const float* rp = p.rows + (size_t)crow[c] * p.pitch;
                okey[c] = float_order_key(exact_distance_rows<true, METRIC>(METRIC, rp, eq, p.dim, p.l2_squared));
            }
        }
        __syncthreads();
        if (n_cand <= kRankSortMax) {
            // Few candidates (the normal case: ~2k): order them by rank counting instead of a bitonic network -- the output
            // position of a candidate is the number of candidates that precede it under (distance key, id, slot).  Eight lanes
            // share one candidate (each counts every 8th rival, three shuffles add up): broadcast shared-memory reads, no barriers.
            const int cnt = min(p.k, n_cand);
            const int sub = threadIdx.x & 7;
            for (int c0 = 0; c0 < n_cand; c0 += blockDim.x >> 3) {
                const int c = c0 + (threadIdx.x >> 3);
                const bool live = c < n_cand;
                const uint32_t key = live ? okey[c] : 0u;
                const long long id = live ? cid[c] : 0;
                int rank = 0, equal = 0;
                if (live)
                    for (int j = sub; j < n_cand; j += 8) {
                        const uint32_t kj = okey[j];
                        rank += kj < key ? 1 : 0;
                        equal += kj == key ? 1 : 0;
                    }
#pragma unroll
                for (int o = 1; o < 8; o <<= 1) {
                    rank += __shfl_xor_sync(0xFFFFFFFFu, rank, o);
                    equal += __shfl_xor_sync(0xFFFFFFFFu, equal, o);
                }
                int tie = 0;
                if (live && equal > 1)                      // exact distance ties (itself included): order those by (id, slot)
                    for (int j = sub; j < n_cand; j += 8) {
                        if (okey[j] != key) continue;
                        const long long ij = cid[j];
                        tie += (ij < id || (ij == id && j < c)) ? 1 : 0;
                    }
#pragma unroll
                for (int o = 1; o < 8; o <<= 1) tie += __shfl_xor_sync(0xFFFFFFFFu, tie, o);
                rank += tie;
                if (live && sub == 0 && rank < cnt) {
                    oid[rank] = id;
                    od[rank] = float_from_order_key(key);
                }
            }
            for (int j = cnt + threadIdx.x; j < p.k_out; j += blockDim.x) {
                oid[j] = -1;
                od[j] = INFINITY;
            }
            if (threadIdx.x == 0 && p.out_count) p.out_count[q] = cnt;
            return;
        }
        for (int i = threadIdx.x; i < n_cand; i += blockDim.x) ckey[i] = okey[i];
        __syncthreads();
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
        bool flagged = false;   // a shard could not answer this query with its bounded selector (count = -1): the merged row
        for (int s = 0; s < shards; ++s) {   // is invalid too and says so, instead of silently missing that shard's rows
            const int c = count[(size_t)s * b + q];
            flagged = flagged || c < 0;
            t += min(max(c, 0), k);
        }
        total_s = flagged ? -1 : t;
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
    const int cnt = total_s < 0 ? -1 : min(k, total_s);
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

// ---- K5b: shard merge fused with its exchange over NVLink peer memory ------------------------------------------
// Every rank keeps its local top-k of the whole batch in a "result block" (ids | distances | counts, see kernels.h) that
// its peers have mapped.  Rank r owns the queries [q_begin, q_begin + q_count): one CTA per query PULLS that query's k
// entries from each of the `world` blocks (P2P loads), merges them, and PUSHES the merged row into every rank's final
// block (P2P stores).  So the all-gather and the merge are one kernel, each rank moves and merges only 1/world of the
// batch, and no rank ever holds the world * b * k gathered lists.
// The lists arrive sorted by (distance key, id), so the merge is rank counting instead of a sort: the final position of
// entry j of list s is j + the number of entries of every other list that precede it (two binary searches' worth of
// shared-memory probes), ties between lists broken by shard index.  No barriers after the load.
// Two delivery modes: push (slice_ids == nullptr) writes the merged rows of this rank's slice into EVERY rank's final
// block; pull-only (slice_ids != nullptr) writes them to plain [q_count][k] arrays on this rank -- the answer stays
// partitioned across the ranks, nothing is pushed and no barrier is needed after the kernel.
// A row some shard flagged (count < 0: its bounded selector could not answer it) is delivered as count = -1.
__global__ void __launch_bounds__(kSelThreads) exchange_merge_kernel(PeerBlocks pb, int world, int b, int k, int q_begin,
                                                                     int64_t* slice_ids, float* slice_dist, int32_t* slice_count) {
    extern __shared__ __align__(16) unsigned char sm[];
    long long* lid = reinterpret_cast<long long*>(sm);                         // [world][k]
    long long* oid = lid + (size_t)world * k;                                  // [k] merged ids
    uint32_t* lkey = reinterpret_cast<uint32_t*>(oid + k);                     // [world][k]
    uint32_t* okey = lkey + (size_t)world * k;                                 // [k] merged keys
    __shared__ int cnt_s[kMaxPeers];
    __shared__ int flagged_s;
    const int q = q_begin + blockIdx.x;
    const size_t dist_off = (size_t)b * k * 8, cnt_off = (size_t)b * k * 12;
    if (threadIdx.x == 0) flagged_s = 0;
    __syncthreads();
    if (threadIdx.x < world) {
        const int c = *reinterpret_cast<const int32_t*>(pb.local[threadIdx.x] + cnt_off + (size_t)q * 4);
        if (c < 0) flagged_s = 1;
        cnt_s[threadIdx.x] = min(max(c, 0), k);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < world * k; i += blockDim.x) {
        const int s = i / k, j = i - s * k;
        if (j < cnt_s[s]) {
            lid[i] = *reinterpret_cast<const long long*>(pb.local[s] + ((size_t)q * k + j) * 8);
            lkey[i] = float_order_key(*reinterpret_cast<const float*>(pb.local[s] + dist_off + ((size_t)q * k + j) * 4));
        }
    }
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
        oid[j] = -1;
        okey[j] = float_order_key(INFINITY);
    }
    __syncthreads();
    int total = 0;
    for (int s = 0; s < world; ++s) total += cnt_s[s];
    for (int i = threadIdx.x; i < world * k; i += blockDim.x) {
        const int s = i / k, j = i - s * k;
        if (j >= cnt_s[s]) continue;
        const uint32_t key = lkey[i];
        const long long id = lid[i];
        int rank = j;
        for (int t = 0; t < world && rank < k; ++t) {
            if (t == s) continue;
            // entries of list t that sort before (key, id, s): strictly smaller pairs, and equal pairs when t < s
            const uint32_t* tk = lkey + (size_t)t * k;
            const long long* ti = lid + (size_t)t * k;
            int lo = 0, hi = cnt_s[t];
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                const uint32_t mk = tk[mid];
                const long long mi = ti[mid];
                const bool before = mk < key || (mk == key && (mi < id || (mi == id && t < s)));
                if (before) lo = mid + 1;
                else hi = mid;
            }
            rank += lo;
        }
        if (rank < k) {
            oid[rank] = id;
            okey[rank] = key;
        }
    }
    __syncthreads();
    const int cnt = flagged_s ? -1 : min(k, total);
    if (slice_ids) {
        const size_t row = (size_t)blockIdx.x * k;
        for (int j = threadIdx.x; j < k; j += blockDim.x) {
            const bool ok = j < cnt;
            slice_ids[row + j] = ok ? oid[j] : -1;
            slice_dist[row + j] = ok ? float_from_order_key(okey[j]) : INFINITY;
        }
        if (threadIdx.x == 0 && slice_count) slice_count[blockIdx.x] = cnt;
        return;
    }
    for (int i = threadIdx.x; i < world * k; i += blockDim.x) {
        const int p = i / k, j = i - p * k;
        const bool ok = j < cnt;
        unsigned char* f = pb.final_[p];
        *reinterpret_cast<long long*>(f + ((size_t)q * k + j) * 8) = ok ? oid[j] : -1;
        *reinterpret_cast<float*>(f + dist_off + ((size_t)q * k + j) * 4) = ok ? float_from_order_key(okey[j]) : INFINITY;
    }
    if (threadIdx.x < world) *reinterpret_cast<int32_t*>(pb.final_[threadIdx.x] + cnt_off + (size_t)q * 4) = cnt;
    __threadfence_system();   // the peers read these rows after the next cross-rank barrier
}

// ---- K5c: global threshold from every shard's published seed bounds -------------------------------------------------
// One CTA per query: pull the k published bounds of each of the `world` shards (P2P loads, k*4 bytes each), take the
// k-th smallest U of the union -- at least k distinct rows of the whole index have exact badness <= U, so the global
// k-th best is <= U -- and tighten this shard's threshold to U + eps (this shard's own error margin for its candidates).
__global__ void __launch_bounds__(kSelThreads) seed_merge_kernel(PeerSeedKeys pk, int world, QueryState* qstate, int k) {
    extern __shared__ __align__(16) unsigned char sm[];
    entry_t* keys = reinterpret_cast<entry_t*>(sm);   // [world][k] keys in the entries' high words (the radix select's layout)
    __shared__ uint32_t hist[256];
    const int q = blockIdx.x;
    const int total = world * k;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int s = i / k, j = i - s * k;
        keys[i] = (entry_t)__ldcg(pk.keys[s] + (size_t)q * k + j) << 32;
    }
    __syncthreads();
    // 4 histogram passes instead of a 32-step bisection (84 us -> ~25 us per 4096-query launch, profiles/r02_breakdown.json)
    const uint32_t u = kth_key_radix(keys, total, k, hist);   // pads are 0xFFFFFFFF: fewer than k bounds => no bound
    if (threadIdx.x == 0 && u < 0xFF800000u) {
        QueryState* qs = qstate + q;
        const float tau_g = widen_half_up(float_from_order_key(u), qs->eps_abs, qs->eps_rel);
        const uint32_t tg = float_order_key(tau_g);
        if (tg < qs->tau_key) qs->tau_key = tg;
    }
}

// ---- sliced seeding, consumer side: one delivered bound per query -> this shard's threshold -------------------------------
__global__ void __launch_bounds__(256) apply_bounds_kernel(const uint32_t* __restrict__ bounds, QueryState* qstate, int b) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= b) return;
    const uint32_t u = __ldcg(bounds + q);
    if (u >= 0xFF800000u) return;                       // the owner had no bound for this query
    QueryState* qs = qstate + q;
    const uint32_t tg = float_order_key(widen_half_up(float_from_order_key(u), qs->eps_abs, qs->eps_rel));
    if (tg < qs->tau_key) qs->tau_key = tg;
}

cudaError_t launch_apply_bounds(const uint32_t* bounds, QueryState* qstate, int b, cudaStream_t stream) {
    if (b <= 0) return cudaSuccess;
    apply_bounds_kernel<<<(b + 255) / 256, 256, 0, stream>>>(bounds, qstate, b);
    return cudaGetLastError();
}

// ---- replicate a slice to every peer (the all-gather of a partitioned query batch, by push) ----------------------------
__global__ void __launch_bounds__(256) peer_push_kernel(const uint4* __restrict__ src, PeerBlocks dst, int n_dst, size_t n16) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = src[i];
        for (int d = 0; d < n_dst; ++d) reinterpret_cast<uint4*>(dst.final_[d])[i] = v;
    }
}

cudaError_t launch_peer_push(const void* src, void* const* dst, int n_dst, size_t nbytes, cudaStream_t stream) {
    if (nbytes == 0 || n_dst <= 0) return cudaSuccess;
    if (n_dst > kMaxPeers || (nbytes & 15) || ((uintptr_t)src & 15)) return cudaErrorInvalidValue;
    PeerBlocks pb{};
    for (int d = 0; d < n_dst; ++d) {
        if (((uintptr_t)dst[d]) & 15) return cudaErrorInvalidValue;
        pb.final_[d] = static_cast<unsigned char*>(dst[d]);
    }
    const size_t n16 = nbytes / 16;
    const int grid = (int)std::min<size_t>((n16 + 255) / 256, 148 * 4);
    peer_push_kernel<<<grid, 256, 0, stream>>>(static_cast<const uint4*>(src), pb, n_dst, n16);
    return cudaGetLastError();
}

cudaError_t launch_seed_merge(const PeerSeedKeys& pk, int world, QueryState* qstate, int b, int k, cudaStream_t stream) {
    if (b <= 0 || k <= 0 || world <= 0) return cudaSuccess;
    const size_t smem = (size_t)world * k * 8;
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(seed_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    seed_merge_kernel<<<b, kSelThreads, smem, stream>>>(pk, world, qstate, k);
    return cudaGetLastError();
}

cudaError_t launch_compact_pool(const SelectParams& p, int b, cudaStream_t stream) {
    size_t smem = (size_t)sort_cap_of(p) * 8 + (size_t)std::max(p.k, 1) * 4;   // candidate entries + the keys staged for publishing
    cudaError_t e = cudaFuncSetAttribute(compact_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    compact_pool_kernel<<<b, kCompactThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_finalize(const SelectParams& p, int b, cudaStream_t stream) {
    size_t smem = (size_t)sort_cap_of(p) * 8 + (size_t)exact_cap_of(p) * 12;
    if ((size_t)sort_cap_of(p) * 8 < 8192 + (size_t)exact_cap_of(p) * 8) return cudaErrorInvalidValue;   // q64 | keys | redo list
    if (p.peer_world < 0 || p.peer_world > kMaxPeers || (long long)p.peer_world * p.k > sort_cap_of(p)) return cudaErrorInvalidValue;
    void (*fn)(SelectParams);
    if (p.accum_f32)
        fn = p.metric == kMetricL2 ? finalize_kernel<kMetricL2, true>
             : p.metric == kMetricCosine ? finalize_kernel<kMetricCosine, true> : finalize_kernel<kMetricIP, true>;
    else
        fn = p.metric == kMetricL2 ? finalize_kernel<kMetricL2, false>
             : p.metric == kMetricCosine ? finalize_kernel<kMetricCosine, false> : finalize_kernel<kMetricIP, false>;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    fn<<<b, kSelThreads, smem, stream>>>(p);
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

cudaError_t launch_exchange_merge(const PeerBlocks& pb, int world, int b, int k, int q_begin, int q_count, cudaStream_t stream,
                                  int64_t* slice_ids, float* slice_dist, int32_t* slice_count) {
    if (q_count <= 0) return cudaSuccess;
    const size_t smem = (size_t)(world + 1) * k * 12;
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(exchange_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    exchange_merge_kernel<<<q_count, kSelThreads, smem, stream>>>(pb, world, b, k, q_begin, slice_ids, slice_dist, slice_count);
    return cudaGetLastError();
}

}  // namespace b200ann
