/*
 * b200ann.h -- C ABI of the B200-native exact nearest-neighbour engine.
 *
 * This is the drop-in boundary for ONE path of sagspot/the-algorithm: the exact dense scan behind
 * com.twitter.ann.brute_force.BruteForceIndex and the ann.common Queryable / Appendable contracts.
 * Plain pointers and sizes only; no C++ or torch types.  A JVM binds it the way the reference binds
 * Faiss: an opaque native handle + caller-owned flat buffers
 * (ann/src/main/java/com/twitter/ann/faiss/swig/Index.java:12-37,95-101; swigfaissJNI.java:267-269).
 * The Scala-side binding a maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - every call returns ANN_OK (0) or a negative ann_status; nothing throws, aborts or exits.
 *     ann_last_error() gives the thread-local message of the last failure on the calling thread.
 *   - host buffers are caller-owned and only borrowed for the duration of the call; outputs are
 *     fully written before return.  `_device` variants take device pointers on the index's GPU and a
 *     cudaStream_t (passed as void*; NULL is the legacy default stream); they enqueue work on that stream, so
 *     they are ordered after whatever produced their inputs there, and return without synchronising.
 *   - results: nearest first, ascending (Float.compare(distance), id) -- the deterministic refinement
 *     of the reference's heap order (BruteForceIndex.scala:73-89).  Row q of the outputs has
 *     out_count[q] = min(max(k,0), size) valid entries; unused slots hold id = -1, distance = +inf.
 *   - there is no CPU fallback: every compute entry point fails with ANN_ERR_NO_DEVICE / ANN_ERR_CUDA
 *     when no sm_100 device is usable.
 *   - a handle may be used from several threads.  Appends copy and prepare their rows without blocking queries (the
 *     storage grows in place: virtual ranges reserved once, physical memory mapped behind the rows) and publish the new row
 *     count at the end; queries on one handle share its scratch and run one device batch at a time, and concurrent small
 *     host queries are combined into one device batch.  A query observes every append whose call returned before the
 *     query was issued; appends in flight may or may not be visible (as with the reference's ConcurrentLinkedQueue).
 */
#ifndef B200ANN_H_
#define B200ANN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define ANN_API __attribute__((visibility("default")))
#else
#define ANN_API
#endif

/* Metric ordinals = thrift DistanceMetric (ann/src/main/thrift/.../ann_common.thrift:16-19);
 * semantics = ann/src/main/scala/com/twitter/ann/common/Metric.scala:88-94 (L2), :119-125 (Cosine),
 * :150-158 (InnerProduct).  EditDistance (:187-261) is a string metric and is out of scope. */
typedef enum ann_metric { ANN_METRIC_L2 = 0, ANN_METRIC_COSINE = 1, ANN_METRIC_INNER_PRODUCT = 2 } ann_metric;

typedef enum ann_status {
    ANN_OK = 0,
    ANN_ERR_INVALID_ARGUMENT = -1,   /* bad metric / dim / config field                                  */
    ANN_ERR_NULL_POINTER = -2,       /* a required pointer was NULL                                      */
    ANN_ERR_DIMENSION_MISMATCH = -3, /* mirrors BadRequest VECTOR_DIMENSION_MISMATCH, ann_common.thrift:146-161 */
    ANN_ERR_NEGATIVE_K = -4,         /* k < 0 (k == 0 is a success with empty results, BruteForceIndex.scala:83) */
    ANN_ERR_NO_DEVICE = -5,          /* no CUDA device / not sm_100                                      */
    ANN_ERR_CUDA = -6,               /* a CUDA runtime or driver call failed (message has the detail)    */
    ANN_ERR_OUT_OF_MEMORY = -7,      /* host or device allocation failed                                 */
    ANN_ERR_CANDIDATE_OVERFLOW = -8, /* more near-ties around rank k than the exact selector can hold    */
    ANN_ERR_UNKNOWN_OPTION = -9
} ann_status;

/* flags for ann_config.flags */
#define ANN_FLAG_L2_SQUARED 0x1u /* return squared L2 (Faiss METRIC_L2 style, QueryableIndexAdapter.scala:174) */
#define ANN_FLAG_NO_SHADOW 0x2u  /* do not keep the bf16 shadow matrix: batched tensor-core path disabled */
/* Accumulator convention of the returned distances.  The reference's arithmetic (EmbeddingMath.Float) is unshipped;
 * MetricUtil.dot is typed Float with no cast (Metric.scala:264-269) while l2distance / cosineSimilarity return Double
 * (:271-283).  Default (flag clear): sums accumulate in fp64 in index order and round once to fp32 (oracle accum=0).
 * With this flag every sum accumulates sequentially in fp32 with individually rounded operations (oracle accum=1), so the
 * neighbour ids match a reference whose EmbeddingMath accumulates in Float.  Selection stays exact under either. */
#define ANN_FLAG_ACCUM_F32 0x4u
/* Cosine as normalised InnerProduct, the way the reference's HNSW / Faiss backends answer Cosine
 * (hnsw/DistanceFunctionGenerator.scala:11-30; Hnsw.scala:149-155; faiss/QueryableIndexAdapter.scala:43-50): every appended
 * row is stored as MetricUtil.norm(row) (Metric.scala:285-289) and every query is normalised the same way, so the distance
 * is 1 - q^.a^ with InnerProduct's arithmetic and a scan reads neither norms nor inv-norms.  Needs metric = Cosine.
 * ann_read_rows returns the stored (unit) rows.  Distances agree with the default Cosine arithmetic to ~1e-7 but are
 * not bit-identical to it; the matching oracle is InnerProduct over oracle-normalised inputs. */
#define ANN_FLAG_COSINE_UNIT_ROWS 0x8u

typedef struct ann_config {
    int32_t metric;        /* ann_metric                                                          */
    int32_t dim;           /* embedding dimension, 1..1024                                        */
    int64_t capacity_hint; /* rows to reserve up front (0 = grow on demand)                       */
    int32_t device;        /* CUDA device ordinal                                                 */
    uint32_t flags;        /* ANN_FLAG_*                                                          */
} ann_config;

typedef struct ann_index ann_index; /* opaque; owns device memory */

/* BruteForceIndex.apply(metric, futurePool, initialEmbeddings) -- BruteForceIndex.scala:29-37.
 * The FuturePool stays on the JVM side; initialEmbeddings become one ann_append_batch call. */
ANN_API int ann_create(const ann_config *cfg, ann_index **out);

/* Releases device memory.  NULL is ignored.  (swig Index.delete(), Index.java:24-37) */
ANN_API void ann_destroy(ann_index *ix);

/* Appendable.append (Api.scala:133-145; BruteForceIndex.scala:48-52), batched: n rows at once.
 * ids[n] (NULL => insertion index), rows[n*dim] row-major fp32.  Nothing is validated beyond the
 * buffer shape (duplicates and non-finite values are stored as given, like the reference). */
ANN_API int ann_append_batch(ann_index *ix, const int64_t *ids, const float *rows, int64_t n);
ANN_API int ann_append_batch_device(ann_index *ix, const int64_t *d_ids, const float *d_rows, int64_t n,
                                    void *stream);

/* Updatable.update (Api.scala:148-150; the HNSW backend implements it, hnsw/Hnsw.scala:149-182): overwrite the embeddings
 * stored at `slots` (insertion indices, 0 <= slot < size; ids keep their value) with rows[n*dim] and refresh their norms /
 * bf16 shadow rows.  The id -> slot map is the caller's (the host mirrors keep one). */
ANN_API int ann_update_batch(ann_index *ix, const int64_t *slots, const float *rows, int64_t n);

/* Reads rows [start, start+n) and their ids back into host buffers (out_ids[n], out_rows[n*dim], either may be NULL).
 * What SerializableBruteForceIndex.toDirectory iterates to persist the index (BruteForceIndex.scala:142-161). */
ANN_API int ann_read_rows(ann_index *ix, int64_t start, int64_t n, int64_t *out_ids, float *out_rows);

/* Number of rows visible to queries (linkedQueue size, BruteForceIndex.scala:34-36). */
ANN_API int ann_size(const ann_index *ix, int64_t *n);

/* Queryable.queryWithDistance (Api.scala:40-50; BruteForceIndex.scala:66-91) for b queries at once.
 * queries[b*dim]; out_ids[b*k], out_dist[b*k], out_count[b] (out_count may be NULL).
 * Queryable.query (BruteForceIndex.scala:56-64) is the same call with out_dist ignored by the caller.
 * dim is passed again so that a mismatching caller is refused instead of reading out of bounds. */
ANN_API int ann_query_batch(ann_index *ix, const float *queries, int32_t b, int32_t dim, int32_t k,
                            int64_t *out_ids, float *out_dist, int32_t *out_count);
ANN_API int ann_query_batch_device(ann_index *ix, const float *d_queries, int32_t b, int32_t dim, int32_t k,
                                   int64_t *d_out_ids, float *d_out_dist, int32_t *d_out_count, void *stream);

/* ComposedQueryable.queryWithDistance (ShardApi.scala:72-86): merge `shards` per-shard result sets
 * (laid out [shards][b][k], counts [shards][b]) into the global top-k per query, canonical order.
 * Device pointers: this is the kernel that runs after the NCCL all-gather of each rank's local top-k. */
ANN_API int ann_merge_topk_device(int32_t device, const int64_t *d_ids, const float *d_dist, const int32_t *d_count,
                                  int32_t shards, int32_t b, int32_t k, int64_t *d_out_ids, float *d_out_dist,
                                  int32_t *d_out_count, void *stream);

/* The same merge fused with its exchange (one kernel instead of all-gather + merge).  A "result block" for (b, k) is one
 * device allocation laid out [ids: b*k int64][dist: b*k float][count: b int32] (ann_result_block_bytes).  Every rank lets
 * ann_query_batch_device write its shard's results into its own local block, all ranks map each other's local and final
 * blocks (CUDA IPC / symmetric memory; peer_local[s], peer_final[s] are rank s's blocks as seen from this process), and after
 * a cross-rank barrier each rank calls this with its own slice [q_begin, q_begin + q_count) of the batch: the kernel pulls
 * the slice's rows from all `world` local blocks over NVLink, merges them in canonical order and pushes the merged rows into
 * all `world` final blocks.  A second cross-rank barrier makes every final block complete.  world <= 16. */
ANN_API int ann_exchange_merge_device(int32_t device, const void *const *peer_local, void *const *peer_final, int32_t world,
                                      int32_t b, int32_t k, int32_t q_begin, int32_t q_count, void *stream);
ANN_API size_t ann_result_block_bytes(int32_t b, int32_t k);

/* Replicate a device buffer into peers' memory with P2P stores: copies nbytes (a multiple of 16, 16-byte aligned pointers)
 * from d_src to every peer_dst[i].  This is the all-gather of a PARTITIONED query batch: each GPU copies only its b/R slice
 * from the host and pushes it into the same rows of every peer's batch buffer, so the queries cross PCIe once instead of R
 * times; one cross-GPU barrier then makes every copy of the batch complete. */
ANN_API int ann_peer_push_device(int32_t device, const void *d_src, void *const *peer_dst, int32_t n_dst, size_t nbytes,
                                 void *stream);

/* The shard-side half of ComposedQueryable.queryWithDistance's fan-out (ShardApi.scala:72-79) when the shards are GPUs of
 * one box: ann_query_batch_device split in two so that the shards can share what they learn before the expensive part.
 *   ann_query_seed_device   prepares the batch, scores a small prefix of this shard's rows and publishes, per query, k
 *                           witnessed upper bounds on the exact distance key into d_seed_keys[b*k] (uint32; 0xFFFFFFFF =
 *                           no bound: shard too small, or a batch this shard answers with its streaming scan);
 *   -- the caller makes every shard's published keys visible to every shard (cross-rank barrier; the key arrays live in
 *      peer-mapped memory, peer_seed_keys[s] is shard s's array as seen from this process) --
 *   ann_query_finish_device takes the k-th smallest bound over all `world` shards as the batch's global threshold (at
 *                           least k rows of the whole index are at least that near, so nothing farther can be in the
 *                           global top-k), scores the shard against it and writes this shard's candidates for the global
 *                           top-k: every row of the shard that can belong to it, with exact distances, in canonical order.
 *                           The list is NOT necessarily the shard's own exact top-k (rows beyond the global threshold may
 *                           be missing while rows just inside its error margin are present) and out_count[q] may be
 *                           smaller than k even when the shard holds more than k rows.
 * Merged with ann_exchange_merge_device / ann_merge_topk_device the lists give exactly the single-index answer.  With 8
 * shards a shard's threshold is as tight as if it had scored 8x the prefix, which removes most of its candidate handling.
 * The two calls must be issued in this order with the same (b, dim, k) and the same stream; any other query, append or
 * update on the handle in between makes the finish call fail with ANN_ERR_INVALID_ARGUMENT (the scratch is shared).
 * world = 0 or peer_seed_keys = NULL: no sharing (the finish call then completes an ordinary local query). */
ANN_API int ann_query_seed_device(ann_index *ix, const float *d_queries, int32_t b, int32_t dim, int32_t k,
                                  uint32_t *d_seed_keys, void *stream);
ANN_API int ann_query_finish_device(ann_index *ix, const float *d_queries, int32_t b, int32_t dim, int32_t k,
                                    const uint32_t *const *peer_seed_keys, int32_t world, int64_t *d_out_ids,
                                    float *d_out_dist, int32_t *d_out_count, void *stream);

/* The same sharded query with a SECOND cross-shard round, so that the exact rescoring also shrinks with the shard count.
 * A shard's own k-th best (of 1/R of the rows) is far looser than the global k-th, so with ann_query_finish_device every
 * shard still rescores ~2k survivors per query.  Split once more:
 *   ann_query_seed_device     as above;                                         -- barrier --
 *   ann_query_filter_device   global seed threshold, tensor-core chunks, last compaction; publishes this shard's k best
 *                             approximate scores, each widened to an upper bound on the exact distance key of its row,
 *                             into d_kth_keys[b*k] (0xFFFFFFFF = no bound);      -- barrier --
 *   ann_query_rescore_device  takes the k-th smallest of all shards' bounds as the global cut, rescores only the rows of
 *                             this shard under it (about 2k/R per query) exactly and writes them in canonical order.
 * Same session rules as above (same (b, dim, k) and stream, nothing else on the handle in between).  A query the bounded
 * selector cannot answer on some shard is reported with out_count[q] = -1 (the row is invalid, not empty); the merges
 * below propagate the -1, so every consumer sees it and can re-ask with "device_fallback" = 1. */
ANN_API int ann_query_filter_device(ann_index *ix, const float *d_queries, int32_t b, int32_t dim, int32_t k,
                                    const uint32_t *const *peer_seed_keys, int32_t world, uint32_t *d_kth_keys,
                                    void *stream);
ANN_API int ann_query_rescore_device(ann_index *ix, const float *d_queries, int32_t b, int32_t dim, int32_t k,
                                     const uint32_t *const *peer_kth_keys, int32_t world, int64_t *d_out_ids,
                                     float *d_out_dist, int32_t *d_out_count, void *stream);

/* The same two publishing phases with PUSH delivery: instead of one local array that every peer then reads over NVLink
 * (400-byte rows pulled from 7 peers per query are latency-bound: 61 us per 4096-query round at 8 GPUs), the keys are written
 * straight into this shard's [b*k] block of EVERY peer's receive buffer -- dst[i] / kth_dst[i] is that block inside peer i's
 * memory as mapped into this process (own copy included) -- and the consuming calls (ann_query_filter_*_device's
 * peer_seed_keys, ann_query_rescore_device's peer_kth_keys) are given pointers into the LOCAL receive buffer. */
ANN_API int ann_query_seed_push_device(ann_index *ix, const float *d_queries, int32_t b, int32_t dim, int32_t k,
                                       uint32_t *const *dst, int32_t n_dst, void *stream);
ANN_API int ann_query_filter_push_device(ann_index *ix, const float *d_queries, int32_t b, int32_t dim, int32_t k,
                                         const uint32_t *const *peer_seed_keys, int32_t world, uint32_t *const *kth_dst,
                                         int32_t n_dst, void *stream);

/* SLICED seeding (round 1 of the sharded query, replaces ann_query_seed_push_device + the seed half of
 * ann_query_filter_push_device).  Instead of every shard seeding ALL b queries over a short prefix, publishing k bounds per
 * query and every consumer selecting the k-th smallest of the R*k bounds, shard s seeds only ITS slice
 * [q_begin, q_begin + q_count) of the batch, over n_slices times the rows, and publishes ONE bound per query: the same
 * tensor-core work per shard, the same number of rows seen per query, k times less to exchange and no selection kernel.
 *   ann_query_seed_slice_push_device   dst[i] = the [b] uint32 bound array inside peer i's memory (own copy included); this
 *                                      shard writes entries [q_begin, q_begin + q_count) of each (0xFFFFFFFF = no bound);
 *                                                                                               -- barrier --
 *   ann_query_filter_bounds_push_device  d_bounds = the LOCAL [b] array every slice owner has written; thresholds of all b
 *                                      queries come from it (plus this shard's own error margin); then as
 *                                      ann_query_filter_push_device (chunks, last compaction, k best bounds pushed to kth_dst);
 *                                                                                               -- barrier --
 *   ann_query_rescore_device           unchanged.
 * The slices of the shards must tile [0, b).  A shard too small to seed publishes "no bound" for its slice; the queries of
 * that slice are then flagged on the other shards (count = -1) and re-answered by the callers' exact fallback. */
ANN_API int ann_query_seed_slice_push_device(ann_index *ix, const float *d_queries, int32_t b, int32_t dim, int32_t k,
                                             int32_t q_begin, int32_t q_count, int32_t n_slices, uint32_t *const *dst,
                                             int32_t n_dst, void *stream);
ANN_API int ann_query_filter_bounds_push_device(ann_index *ix, const float *d_queries, int32_t b, int32_t dim, int32_t k,
                                                const uint32_t *d_bounds, int32_t world, uint32_t *const *kth_dst,
                                                int32_t n_dst, void *stream);

/* Pull-only form of ann_exchange_merge_device: merge this rank's slice [q_begin, q_begin + q_count) of the batch from all
 * `world` local result blocks (P2P loads) into plain arrays on this device -- d_out_ids / d_out_dist [q_count*k],
 * d_out_count [q_count] (may be NULL).  The merged answer stays partitioned across the ranks: nothing is pushed, and no
 * barrier is needed after the kernel (only before it, so that every local block is complete). */
ANN_API int ann_exchange_merge_slice_device(int32_t device, const void *const *peer_local, int32_t world, int32_t b,
                                            int32_t k, int32_t q_begin, int32_t q_count, int64_t *d_out_ids,
                                            float *d_out_dist, int32_t *d_out_count, void *stream);

/* ---- one process, several GPUs: ShardedAppendable + ComposedQueryable as ONE handle (ShardApi.scala:34-48, 58-87) --------
 * The form a single-JVM host calls.  ann_sharded_create builds one shard per listed device (device_ids == NULL: devices
 * 0..n_devices-1; cfg->device is ignored, cfg->capacity_hint is the total over all shards; a device may be listed more
 * than once, which puts several shards on it) and enables peer access between them.  ann_sharded_append_batch cuts a batch into n_devices contiguous parts, part s to shard s, copied in parallel
 * (ids == NULL => insertion index of the composed index).  ann_sharded_query_batch answers b host queries into host
 * buffers: every device runs the three-phase sharded query above on its own stream from its own host thread, ordered
 * across devices by CUDA events, then merges and returns its 1/n_devices slice of the batch.  The result is bit for bit
 * the single-index answer (globally unique ids).  Flagged queries are re-answered with every shard's exact fallback. */
typedef struct ann_sharded_index ann_sharded_index;
ANN_API int ann_sharded_create(const ann_config *cfg, const int32_t *device_ids, int32_t n_devices,
                               ann_sharded_index **out);
ANN_API void ann_sharded_destroy(ann_sharded_index *sx);
ANN_API int ann_sharded_append_batch(ann_sharded_index *sx, const int64_t *ids, const float *rows, int64_t n);
ANN_API int ann_sharded_size(const ann_sharded_index *sx, int64_t *n);
ANN_API int ann_sharded_query_batch(ann_sharded_index *sx, const float *queries, int32_t b, int32_t dim, int32_t k,
                                    int64_t *out_ids, float *out_dist, int32_t *out_count);
/* Borrow shard `shard` (owned by the composed handle) and its row count. */
ANN_API int ann_sharded_shard(ann_sharded_index *sx, int32_t shard, ann_index **out, int64_t *rows);
/* "two_round" (1 = seed + k-best rounds, default; 0 = seed round only), "sliced_seeds" (two_round only; 1 = every shard
 * seeds its slice of the batch and delivers one bound per query, default; 0 = every shard seeds every query and delivers k
 * bounds per query); any other name is applied to every shard.
 * Stats: "shards", "peer_access", "fallback_batches", "queries", "dim"; any other name is summed over the shards. */
ANN_API int ann_sharded_set_option(ann_sharded_index *sx, const char *name, int64_t value);
ANN_API int ann_sharded_get_stat(const ann_sharded_index *sx, const char *name, int64_t *value);

/* ---- the reference's on-disk format (csrc/persist.cu) -------------------------------------------------------------------
 * SerializableBruteForceIndex.toDirectory (BruteForceIndex.scala:142-161) / BruteForceDeserialization.fromDirectory
 * (BruteForceDeserialization.scala:42-63): one file `BruteForceFileData` holding back-to-back TBinaryProtocol encodings of
 * PersistedEmbedding {1: binary id, 2: embedding.Embedding} (serialization.thrift:7-10; ThriftIteratorIO.scala:14-56) until
 * end of file; sharded indexes are `shard_<i>/` sub-directories (ShardedSerialization.scala:9-11,28-66).  A `_SUCCESS`
 * marker is added (IndexOutputFile.scala:29,58-62).
 *   id_format : how Injection[T, Array[Byte]] encodes the id (AnnInjections.scala:8-12) -- Long = 8 bytes big-endian, Int = 4;
 *               ANN_ID_AUTO accepts either when reading and writes Long.
 *   layout    : which member of the tensor union carries the floats when WRITING (the inner embedding.thrift is not in the
 *               open-source tree; the assumption is stated in csrc/persist.cu).  The reader accepts all three.
 * ann_load_directory: cfg gives metric / device / flags / capacity_hint; cfg->dim may be 0 (taken from the first record).
 * A truncated trailing record ends the stream silently, like the reference's END_OF_FILE handling. */
#define ANN_ID_AUTO 0
#define ANN_ID_INT64_BE 1
#define ANN_ID_INT32_BE 2
#define ANN_LAYOUT_FLOAT_TENSOR 0  /* Embedding{1: GeneralTensor{5: FloatTensor{1: list<double>}}}  (default) */
#define ANN_LAYOUT_DOUBLE_TENSOR 1 /* Embedding{1: GeneralTensor{6: DoubleTensor{1: list<double>}}}            */
#define ANN_LAYOUT_RAW_FLOAT 2     /* Embedding{1: GeneralTensor{1: RawTypedTensor{1: FLOAT, 2: LE float32 bytes}}} */
ANN_API int ann_save_directory(ann_index *ix, const char *directory, int32_t id_format, int32_t layout);
ANN_API int ann_load_directory(const ann_config *cfg, const char *directory, int32_t id_format, ann_index **out);
ANN_API int ann_sharded_save_directory(ann_sharded_index *sx, const char *directory, int32_t id_format, int32_t layout);
/* Loads `shard_<i>/` sub-directories (or a plain index directory) into a composed handle over the listed devices; rows are
 * re-dealt over the devices, so the directory may have been written with any number of shards. */
ANN_API int ann_sharded_load_directory(const ann_config *cfg, const char *directory, int32_t id_format,
                                       const int32_t *device_ids, int32_t n_devices, ann_sharded_index **out);
/* The record codec alone (host code, no device): encode one PersistedEmbedding into `out` (returns its size in bytes, also
 * when out == NULL or too small; -1 on bad arguments), decode the record at the head of `bytes` (*consumed = 0 at end of
 * stream; row may be NULL to learn *dim first). */
ANN_API int64_t ann_persisted_embedding_encode(int64_t id, int32_t id_format, const float *row, int32_t dim, int32_t layout,
                                               unsigned char *out, int64_t capacity);
ANN_API int ann_persisted_embedding_decode(const unsigned char *bytes, int64_t len, int32_t id_format, int64_t *id,
                                           float *row, int32_t row_capacity, int32_t *dim, int64_t *consumed);

/* KnnHelper.findNearestNeighbours (ann/src/main/scala/com/twitter/ann/scalding/offline/KnnHelper.scala:168-215, 248-347):
 * the exact k nearest corpus rows of every query, host buffers in and out -- the offline all-pairs job behind
 * KnnTruthSetGenerator.  The corpus is cut into tiles of corpus_tile_rows rows that fit the device (<= 0: as many as fit;
 * one tile = one index = one of the reference's "search groups"), queries stream through in tiles of query_tile (<= 0: 4096)
 * with the H2D copy of the next tile and the D2H copy of the previous one overlapped with the kernels of the current one, and
 * per-corpus-tile lists are merged on the device in canonical (distance, id) order, so the answer equals one index over the
 * whole corpus.  cfg gives metric / dim / device / flags.  out_ids / out_dist are [nq][k], out_count [nq] (may be NULL). */
ANN_API int ann_knn_join(const ann_config *cfg, const int64_t *corpus_ids, const float *corpus_rows, int64_t n,
                         const float *queries, int64_t nq, int32_t k, int64_t corpus_tile_rows, int32_t query_tile,
                         int64_t *out_ids, float *out_dist, int32_t *out_count);

/* The Metric trait itself (Metric.scala:76-86) for plain vector pairs, host buffers: out[i] = metric.distance(a_i, b_i) for
 * the n rows of a[n*dim] and b[n*dim] -- L2 (Metric.scala:89-94), Cosine (:120-125), InnerProduct (:153-158), computed on the
 * device with exactly the arithmetic the query path returns (DESIGN.md conventions C1-C4; flags: ANN_FLAG_L2_SQUARED).
 * `absoluteDistance` (:82-86) is the same number for these three metrics. */
ANN_API int ann_distance_pairs(int32_t metric, uint32_t flags, int32_t dim, const float *a, const float *b, int64_t n,
                               float *out, int32_t device);

/* MetricUtil.norm (Metric.scala:285-289): out = rows with every row scaled to unit L2 norm (convention C8: squared norm
 * accumulated in fp64 in index order, fp64 divide, one rounding to fp32; a zero row becomes NaN).  This is the step the
 * reference's HNSW / Faiss backends apply before treating Cosine as InnerProduct (DistanceFunctionGenerator.scala:11-15,
 * Hnsw.scala:149-155, QueryableIndexAdapter.scala:43-50); the brute-force index does NOT need it (norms are kept per row). */
ANN_API int ann_normalize_rows(int32_t dim, const float *rows, int64_t n, float *out, int32_t device);

/* Concurrent single-vector load against one handle, measured inside the library -- the C-ABI counterpart of the reference's
 * load generator (service/loadtest/AnnLoadTestWorker.scala:92-115; latency percentiles in microseconds and achieved rate as
 * LoadTestRecorder.scala:114-187 reports them).  `threads` host threads each issue `calls_per_thread` one-vector
 * ann_query_batch calls over queries[nq*dim] (what QueryIndexThriftController.scala:39-90 does per RPC); with expect_ids
 * ([nq*k], may be NULL) every answer is compared and mismatches counted.  Concurrent small calls on one handle are combined
 * into device batches by the library's micro-batcher (options "coalesce_max_batch", "coalesce_small_b"). */
typedef struct ann_load_stats {
    double qps, avg_us, p50_us, p90_us, p99_us, wall_seconds;
    int64_t calls, device_batches, mismatches;
} ann_load_stats;
ANN_API int ann_loadtest(ann_index *ix, const float *queries, int32_t nq, int32_t dim, int32_t k, int32_t threads,
                         int32_t calls_per_thread, const int64_t *expect_ids, ann_load_stats *out);

/* Tuning / introspection.
 * Options: "path" (0 auto, 1 streaming scan, 2 tensor-core GEMM filter, 3 exact fallback for every query), "gemm_min_batch", "gemm_cta_group" (1|2), "gemm_epi_warps" (0 auto, 8, 16),
 *          "gemm_mma_warps" (MMA-issuing threads per CTA pair: 2 = two warps alternate query tiles, default; 1 = a single issuer),
 *          "gemm_hit_budget" (candidates one chunk may add per query, default 500), "gemm_seed_rows" (0 = default 65536),
 *          "gemm_growth_pct" (chunk growth factor in percent, 0 = derived from the hit budget),
 *          "timing" (1 = bracket every scan / GEMM-filter launch with CUDA events on its stream),
 *          "coalesce_max_batch" (queries per merged device call of the host entry point's micro-batcher, default 2048; 0 = off),
 *          "coalesce_small_b" (host calls with at most this many queries are combined with concurrent ones, default 32),
 *          "coalesce_linger_us" (how long a caller that inherits the lead waits for the previous batch's callers to return, default 100),
 *          "device_fallback" (1 = ann_query_batch_device synchronises its stream and re-answers flagged queries with the
 *          exact fallback, like the host entry point always does; 0 = stay asynchronous and report them, default).
 * Stats:   "launches" (kernels launched so far), "last_path", "shadow_bytes", "row_bytes", "n_special", "capacity", "dim",
 *          "sm_count", "kernel_us" / "kernel_launches_timed" (accumulated since "timing" was set; synchronises),
 *          "coalesced_batches" / "coalesced_calls" (device batches the micro-batcher ran / host calls it served),
 *          "mapped_bytes" (physical memory mapped behind the index arrays), "max_rows" (rows the reserved address range holds),
 *          "pending_error" (synchronises the device and returns the sticky selector-overflow status, if any). */
ANN_API int ann_set_option(ann_index *ix, const char *name, int64_t value);
ANN_API int ann_get_stat(const ann_index *ix, const char *name, int64_t *value);

/* Thread-local message for the last non-zero status returned on this thread ("" if none). */
ANN_API const char *ann_last_error(void);

/* Library version as major*10000 + minor*100 + patch. */
ANN_API int ann_version(void);

#ifdef __cplusplus
}
#endif
#endif /* B200ANN_H_ */
