// kernels.h -- host-visible launch interfaces of the CUDA kernels (internal to libb200ann.so).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "common.cuh"

namespace b200ann {

// ---------------------------------------------------------------- K1: append (norms, shadow, flags)
struct AppendParams {
    const float* rows;      // [n_new][pitch] already in place inside the index matrix
    long long row0;         // first new row (global index inside the shard)
    const long long* slots; // optional: row r lives at slots[r] instead of row0 + r (Updatable.update)
    const float* staged;    // with slots: the new embeddings [n_new][dim], scattered into `rows` by the kernel
    long long n_new;
    int dim, pitch, metric;
    float* inv_norm;        // [cap]  1/|a|            (Cosine only, else nullptr)
    float* row_norm;        // [cap]  |a| rounded up   (always)
    __nv_bfloat16* shadow;  // [cap][kp] or nullptr
    int kp;                 // shadow pitch in elements (multiple of 8)
    uint32_t* max_norm_bits;  // running max of |a| (float bits, non-negative => integer order)
    uint32_t* max_resid_bits; // running max of |shadow row - exact operand row|, the measured bf16 rounding error (float bits)
    unsigned long long* n_special;   // rows with non-finite entries or (Cosine) zero norm
    uint32_t* special_list;          // [kSpecialCap] indices of the first special rows (tensor-core path rescoring list)
};
cudaError_t launch_append(const AppendParams& p, cudaStream_t stream);

// MetricUtil.norm (Metric.scala:285-289, convention C8) on device rows: in place over rows [row0, row0+n) of a pitched matrix
// (out == nullptr) or from a dense [n][dim] source into `out`.  metric_kernels.cu.
cudaError_t launch_normalize_rows_device(float* rows, long long row0, long long n, int dim, int pitch, float* out, cudaStream_t stream);

// ---------------------------------------------------------------- query preparation
struct PrepParams {
    const float* queries;   // [b][dim]
    int b, dim, pitch, metric, kp;
    float* q_padded;        // [b][pitch]  zero padded fp32 copy (scan path)
    __nv_bfloat16* q_shadow;  // [b_pad][qkp] bf16 operand (gemm path) or nullptr; only the first kp columns are used
    int qkp;                  // query operand pitch in elements: kp rounded up to 64 (128-byte aligned rows => 4 sectors per box row)
    QueryState* qstate;     // [b]
    const uint32_t* max_norm_bits;
    const uint32_t* max_resid_bits;
    int path;               // 1 = scan (fp32 error model), 2 = gemm (bf16 error model)
    int accum_f32;          // the exact distances are fp32-accumulated: their own rounding error joins the margins
    uint32_t* pub_keys;     // [b][pub_stride] reset to 0xFFFFFFFF (scan path) or nullptr
    int pub_stride;
    unsigned long long* bad_queries;  // count of queries with non-finite entries / zero norm under Cosine
};
cudaError_t launch_prep_queries(const PrepParams& p, cudaStream_t stream);

// ---------------------------------------------------------------- K2: streaming scan
struct ScanParams {
    const float* rows;
    long long n_rows;
    int pitch, metric;
    const float* inv_norm;
    const float* queries;   // [nq][pitch] padded
    int nq;
    QueryState* qstate;     // [nq]
    entry_t* pool;          // [nq][pool_cap]
    int pool_cap;
    uint32_t* special_rows; // [nq][kSpecialCap]
    uint32_t* pub_keys;     // [nq][pub_stride]
    int pub_stride;
    int k, r_pub, j_pub;
    int warps, cap;
};
size_t scan_smem_bytes(int qb, int pitch, int warps, int cap);
cudaError_t launch_scan(const ScanParams& p, int qb, int grid, size_t smem, cudaStream_t stream);

// ---------------------------------------------------------------- selection: compaction + exact finalize
constexpr int kMaxPeers = 16;
// `keys[s]` is shard s's published [b][k] bound array mapped into this process (peer memory over NVLink)
struct PeerSeedKeys {
    const uint32_t* keys[kMaxPeers];
};
struct SelectParams {
    QueryState* qstate;     // [b]
    entry_t* pool;          // [b][pool_cap]
    int pool_cap;
    const uint32_t* special_rows;  // [b][kSpecialCap]  per-query lists written by the scan
    const uint32_t* global_special_rows;  // tensor-core path: the index-wide list of special rows, same for every query
    int global_special_count;
    const uint32_t* pub_keys;      // [b][pub_stride] or nullptr
    int pub_stride, pub_count, j_pub;
    int k;
    int seed_count;         // > 0: the pool holds exactly this many seed entries (group maxima); derive tau, discard them
    uint32_t* seed_keys_out;  // with seed_count: also publish, per query, the k best group maxima as upper bounds on exact
                              // badness (order keys of g + eps, [b][k], 0xFFFFFFFF padded) for the other shards (K5c);
                              // without seed_count (compaction after the last chunk): the k best pool entries, likewise
    uint32_t* push_keys[kMaxPeers];   // n_push > 0: publish by PUSH instead -- this shard's [b][k] block inside every peer's
    int n_push;                       // receive buffer (P2P stores are fire-and-forget; pulling 400-byte rows from 7 peers
                                      // per query was latency-bound: 61 us per 4096-query round at 8 GPUs)
    int pub_single;                   // seed compaction of a query SLICE (sliced seeding): publish ONE key per query -- the
                                      // k-th best group maximum as an upper bound -- at push_keys[d][q] / seed_keys_out[q]
    PeerSeedKeys peer_keys;   // finalize: every shard's bounds published after the last chunk (second cross-shard round)
    int peer_world;           // number of valid peer_keys entries; 0 = single shard, no global bound
    int sort_cap, exact_cap;  // shared-memory capacities (entries) of the approximate and exact stages; 0 = 4096 / 2048.
                              // exact_cap must be a power of two; exceeding either flags the query for the exact fallback
    // exact rescoring inputs
    const float* rows;
    const int64_t* ids;
    const float* row_norm;  // [n_rows] |a| rounded up (K1): bounds the summation error of the lane-parallel rescoring
    long long n_rows;
    int dim, pitch, metric, l2_squared;
    int accum_f32;          // ANN_FLAG_ACCUM_F32: exact distances accumulate sequentially in fp32 (oracle accum=1)
    const float* queries;   // [b][q_pitch] fp32
    int q_pitch;
    int64_t* out_ids;       // [b][k_out]
    float* out_dist;
    int32_t* out_count;
    int k_out;
};
// approx-only: sort the pool, keep everything within the margin of the k-th best, tighten tau
cudaError_t launch_compact_pool(const SelectParams& p, int b, cudaStream_t stream);
// compaction + exact fp64 rescoring + (distance, id) ordering + output
cudaError_t launch_finalize(const SelectParams& p, int b, cudaStream_t stream);
// sliced seeding, consumer side: tau_q <- min(tau_q, bound_q + this shard's margin) for the b bounds the slice owners delivered
cudaError_t launch_apply_bounds(const uint32_t* bounds, QueryState* qstate, int b, cudaStream_t stream);
// fill outputs for k == 0 / empty index
cudaError_t launch_fill_empty(int64_t* out_ids, float* out_dist, int32_t* out_count, int b, int k_out, cudaStream_t stream);

// ---------------------------------------------------------------- K5: shard merge
cudaError_t launch_merge(const int64_t* ids, const float* dist, const int32_t* count, int shards, int b, int k,
                         int64_t* out_ids, float* out_dist, int32_t* out_count, cudaStream_t stream);

// sets the thread-local message behind ann_last_error() and returns `code` (index.cu)
int report_error(int code, const char* msg);

// K5b: the same merge fused with the exchange.  A "result block" for (b, k) is one allocation laid out
//   [ids: b*k int64][dist: b*k float][count: b int32]        (result_block_bytes)
// `local[s]` is rank s's block of per-shard results and `final_[s]` its block for the merged answer, both mapped into this
// process (peer memory over NVLink).  This rank merges the queries [q_begin, q_begin + q_count) and writes them to every final block.
struct PeerBlocks {
    const unsigned char* local[kMaxPeers];
    unsigned char* final_[kMaxPeers];
};
inline size_t result_block_bytes(long long b, long long k) { return (size_t)(b * k * 12 + b * 4); }
// slice_ids != nullptr: pull-only delivery into plain [q_count][k] arrays of this rank (final_ is not touched)
cudaError_t launch_exchange_merge(const PeerBlocks& pb, int world, int b, int k, int q_begin, int q_count, cudaStream_t stream,
                                  int64_t* slice_ids = nullptr, float* slice_dist = nullptr, int32_t* slice_count = nullptr);

// copy `nbytes` (multiple of 16, 16-byte aligned) from local `src` to each of `dst[0..n_dst)` with P2P stores
cudaError_t launch_peer_push(const void* src, void* const* dst, int n_dst, size_t nbytes, cudaStream_t stream);

// K5c: threshold seeding shared between the shards of one index.  Every shard publishes, per query, k witnessed upper
// bounds on exact badness (SelectParams::seed_keys_out); the k-th smallest over the union of all shards' bounds is a bound
// on the GLOBAL k-th best, so a shard may discard everything above it (+ its own error margin) even where its own rows
// alone would not justify that.  `keys[s]` is shard s's [b][k] key array mapped into this process (peer memory).
cudaError_t launch_seed_merge(const PeerSeedKeys& pk, int world, QueryState* qstate, int b, int k, cudaStream_t stream);

// ---------------------------------------------------------------- exact fallback (degenerate ties, NaN queries)
struct FallbackParams {
    const float* rows;
    const int64_t* ids;
    long long n_rows;
    int pitch, dim, metric, l2_squared, accum_f32;
    const float* query;     // [dim] fp32 on the device
    void* scratch;          // fallback_scratch_bytes(n_rows, k)
    int k, k_out;           // k = min(requested k, n_rows)
    int64_t* out_ids;       // [k_out] slots of this query
    float* out_dist;
    int32_t* out_count;     // may be nullptr
};
size_t fallback_scratch_bytes(long long n, int k);
cudaError_t launch_exact_fallback(const FallbackParams& p, cudaStream_t stream, long long* launches);

// ---------------------------------------------------------------- K3: tcgen05 GEMM filter
struct GemmLaunch {
    const void* q_shadow;       // bf16 [b_pad][qkp]
    int qkp;
    const void* shadow;         // bf16 [n_rows_total][kp]
    long long n_rows_total;
    long long row_begin, row_end;   // chunk of corpus rows scored by this launch
    int b, b_pad, kp;
    int cta_group;              // 1 or 2 (tcgen05 cta_group)
    int seed_mode;              // 1: threshold seeding launch (group maxima at fixed pool slots)
    int epi_warps;              // 8 or 16 epilogue warps (hit-dense chunks want 16); anything else = 8
    int mma_warps;              // 1 or 2 MMA-issuing threads (2: warps 1 and 3 alternate query tiles); anything else = default
    int nb_stages;              // from gemm_row_stages()
    int sm_count;
    QueryState* qstate;
    entry_t* pool;
    int pool_cap;
};
cudaError_t launch_gemm_filter(const GemmLaunch& g, cudaStream_t stream);
int gemm_seed_group_rows(int cta_group);
int gemm_row_stages(int kp, size_t smem_optin);   // 1 when a shared-memory plan exists for this operand width, else 0 (use the scan)

}  // namespace b200ann
