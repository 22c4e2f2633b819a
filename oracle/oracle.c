/*
 * oracle.c -- CPU restatement of the reference's exact-kNN path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg may
 * load this library; the product (the-algorithm_b200/) never links, imports or calls it.
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for ann/ and its
 * vector arithmetic (com.twitter.ml.api.embedding.EmbeddingMath) is an unshipped monorepo target
 * (ann/src/main/scala/com/twitter/ann/common/Api.scala:4-6,12; common/BUILD:15).  The JVM cannot run
 * here either.  What this file pins is therefore (i) the control flow the reference DOES ship and
 * (ii) the arithmetic conventions C1..C7 stated in SURVEY.md section 8(c) / DESIGN.md.
 *
 * What each function follows (paths relative to /root/reference/ann/src/main/scala/com/twitter/ann):
 *   total order on distances ........ common/Metric.scala:17-36   (Ordering.Float.compare)
 *   L2 / Cosine / InnerProduct ...... common/Metric.scala:88-94, 119-125, 150-158, 263-290
 *   append, scan, bounded heap ...... brute_force/BruteForceIndex.scala:26-92
 *   heap mechanics .................. scala.collection.mutable.PriorityQueue, Scala 2.12.x stdlib
 *                                     (third-party; brute_force/BUILD:4 pins platform java8 / 2.12):
 *                                     1-indexed array heap, `+=` = append + sift-up with strict `<`,
 *                                     `dequeue` = move last to root + sift-down preferring the right
 *                                     child only when strictly greater, stop when parent >= child.
 *   shard merge ..................... common/ShardApi.scala:58-87  (flatten, stable sort, take k)
 *
 * Arithmetic conventions (switchable where noted):
 *   C1  sums accumulate in fp64 over i = 0..d-1 in index order and round ONCE to fp32
 *       (accum=1 switches to fp32 sequential accumulation, used only for the sensitivity report)
 *   C2  InnerProduct = 1.0f - (float)dot                              Metric.scala:157, :268
 *   C3  Cosine = 1.0f - (float)(dot / (sqrt(na) * sqrt(nb)))          Metric.scala:124, :282
 *       zero norm => 0/0 = NaN, which Float.compare orders last
 *   C4  L2 = (float)sqrt(sum (a-b)^2)                                 Metric.scala:93, :275
 *       (l2_squared=1 skips the sqrt, Faiss style)
 *   C5  canonical order = ascending (Float.compare(distance), id)
 *   C6  argument order distance(row, query)                           BruteForceIndex.scala:81
 *   C7  duplicates allowed, nothing validated
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -pthread).  -ffp-contract=off matters: the
 * GPU rescoring kernel uses explicit __dmul_rn/__dadd_rn and must see the same roundings.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define ORACLE_L2 0 /* thrift ordinals, ann/src/main/thrift/.../ann_common.thrift:16-19 */
#define ORACLE_COSINE 1
#define ORACLE_INNER_PRODUCT 2

/* ------------------------------------------------------------------------------------------ */
/* tiny pthread parallel-for over queries (dynamic, one query per grab): the reference runs each */
/* query as one task on a FuturePool(threads) (UnsafeQueryIndexServer.scala:26-32,64-68).       */
typedef void (*query_fn)(int q, void *ctx);
typedef struct {
    query_fn fn;
    void *ctx;
    int b;
    volatile int next;
} pfor_t;

static void *pfor_worker(void *arg) {
    pfor_t *p = (pfor_t *)arg;
    for (;;) {
        int q = __sync_fetch_and_add(&p->next, 1);
        if (q >= p->b) break;
        p->fn(q, p->ctx);
    }
    return NULL;
}

int oracle_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

static void parallel_for(int b, int nthreads, query_fn fn, void *ctx) {
    if (nthreads <= 0) nthreads = oracle_max_threads();
    if (nthreads > b) nthreads = b;
    pfor_t p = {fn, ctx, b, 0};
    if (nthreads <= 1) {
        pfor_worker(&p);
        return;
    }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    int started = 0;
    for (int t = 0; t < nthreads; ++t)
        if (pthread_create(&th[started], NULL, pfor_worker, &p) == 0) ++started;
    if (!started) pfor_worker(&p);
    for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
    free(th);
}

/* ------------------------------------------------------------------------------------------ */
/* java.lang.Float.compare as an unsigned key: -0.0 < +0.0, every NaN equal and greatest.      */
static inline uint32_t float_order_key(float f) {
    uint32_t b;
    if (f != f) return 0xFFFFFFFFu; /* NaN: above +inf (0xFF800000) */
    memcpy(&b, &f, 4);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

uint32_t oracle_float_order_key(float f) { return float_order_key(f); }

static inline int float_compare(float a, float b) {
    uint32_t ka = float_order_key(a), kb = float_order_key(b);
    return (ka > kb) - (ka < kb);
}

/* ------------------------------------------------------------------------------------------ */
/* metric.distance(row, query)                                                                 */
static float distance_f64(int metric, const float *a, const float *b, int d, int l2_squared) {
    if (metric == ORACLE_L2) {
        double acc = 0.0;
        for (int i = 0; i < d; ++i) {
            double diff = (double)a[i] - (double)b[i];
            double sq = diff * diff;
            acc = acc + sq;
        }
        return (float)(l2_squared ? acc : sqrt(acc));
    }
    double dot = 0.0;
    for (int i = 0; i < d; ++i) dot = dot + (double)a[i] * (double)b[i];
    if (metric == ORACLE_INNER_PRODUCT) return 1.0f - (float)dot;
    double na = 0.0, nb = 0.0;
    for (int i = 0; i < d; ++i) na = na + (double)a[i] * (double)a[i];
    for (int i = 0; i < d; ++i) nb = nb + (double)b[i] * (double)b[i];
    double cs = dot / (sqrt(na) * sqrt(nb));
    return 1.0f - (float)cs;
}

static float distance_f32(int metric, const float *a, const float *b, int d, int l2_squared) {
    if (metric == ORACLE_L2) {
        float acc = 0.0f;
        for (int i = 0; i < d; ++i) {
            float diff = a[i] - b[i];
            float sq = diff * diff;
            acc = acc + sq;
        }
        return l2_squared ? acc : sqrtf(acc);
    }
    float dot = 0.0f;
    for (int i = 0; i < d; ++i) {
        float p = a[i] * b[i];
        dot = dot + p;
    }
    if (metric == ORACLE_INNER_PRODUCT) return 1.0f - dot;
    float na = 0.0f, nb = 0.0f;
    for (int i = 0; i < d; ++i) {
        float p = a[i] * a[i];
        na = na + p;
    }
    for (int i = 0; i < d; ++i) {
        float p = b[i] * b[i];
        nb = nb + p;
    }
    return 1.0f - dot / (sqrtf(na) * sqrtf(nb));
}

static inline float distance(int metric, const float *a, const float *b, int d, int accum, int l2sq) {
    return accum ? distance_f32(metric, a, b, d, l2sq) : distance_f64(metric, a, b, d, l2sq);
}

float oracle_distance(int metric, const float *row, const float *query, int d, int accum, int l2_squared) {
    return distance(metric, row, query, d, accum, l2_squared);
}

/* MetricUtil.norm (Metric.scala:285-289) -> EmbeddingMath.Float.normalize (unshipped): L2-normalise a vector, the step
 * HNSW / Faiss use to turn Cosine into InnerProduct (DistanceFunctionGenerator.scala:11-15; Hnsw.scala:149-155).
 * Convention C8 (unpinned like C1-C4): the squared norm accumulates in fp64 in index order, each element is divided in
 * fp64 by sqrt(norm2) and rounded once to fp32.  A zero vector gives 0/0 = NaN in every element. */
void oracle_normalize(const float *v, int d, float *out) {
    double n2 = 0.0;
    for (int i = 0; i < d; ++i) n2 = n2 + (double)v[i] * (double)v[i];
    double nrm = sqrt(n2);
    for (int i = 0; i < d; ++i) out[i] = (float)((double)v[i] / nrm);
}

/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int64_t id;
    float dist;
    int64_t seq; /* insertion slot, used only by the faithful heap for bookkeeping */
} neighbor_t;

/* canonical order C5: (Float.compare(distance), id) ascending */
static inline int canonical_less(const neighbor_t *x, const neighbor_t *y) {
    int c = float_compare(x->dist, y->dist);
    if (c) return c < 0;
    return x->id < y->id;
}

/* bounded max-heap under the canonical order (worst on top); deterministic by construction     */
static void canon_sift_down(neighbor_t *h, int n, int i) {
    for (;;) {
        int l = 2 * i + 1, r = l + 1, m = i;
        if (l < n && canonical_less(&h[m], &h[l])) m = l;
        if (r < n && canonical_less(&h[m], &h[r])) m = r;
        if (m == i) return;
        neighbor_t t = h[i];
        h[i] = h[m];
        h[m] = t;
        i = m;
    }
}

static void canon_sift_up(neighbor_t *h, int i) {
    while (i > 0) {
        int p = (i - 1) / 2;
        if (!canonical_less(&h[p], &h[i])) return;
        neighbor_t t = h[i];
        h[i] = h[p];
        h[p] = t;
        i = p;
    }
}

static int canon_cmp_qsort(const void *a, const void *b) {
    const neighbor_t *x = (const neighbor_t *)a, *y = (const neighbor_t *)b;
    if (canonical_less(x, y)) return -1;
    if (canonical_less(y, x)) return 1;
    return 0;
}

/*
 * Canonical exact top-k over a contiguous corpus.  Same scan as BruteForceIndex.scala:76-86 but the
 * selection order is the deterministic refinement C5, so the answer does not depend on heap history.
 * out_* are [b*k]; unused slots get id=-1, dist=+inf; out_count[q] = min(max(k,0), n).
 */
typedef struct {
    int metric, d, k, accum, l2_squared;
    const float *corpus, *queries;
    const int64_t *ids;
    int64_t n;
    int64_t *out_ids;
    float *out_dist;
    int32_t *out_count;
} canon_ctx_t;

static void canon_one_query(int q, void *vctx) {
    canon_ctx_t *c = (canon_ctx_t *)vctx;
    int k = c->k, d = c->d;
    const float *qv = c->queries + (size_t)q * d;
    neighbor_t *heap = (neighbor_t *)malloc(sizeof(neighbor_t) * (size_t)(k > 0 ? k : 1));
    int hn = 0;
    if (k > 0) {
        for (int64_t r = 0; r < c->n; ++r) {
            neighbor_t nb;
            nb.id = c->ids ? c->ids[r] : r;
            nb.dist = distance(c->metric, c->corpus + (size_t)r * d, qv, d, c->accum, c->l2_squared);
            nb.seq = r;
            if (hn < k) {
                heap[hn] = nb;
                canon_sift_up(heap, hn);
                ++hn;
            } else if (canonical_less(&nb, &heap[0])) {
                heap[0] = nb;
                canon_sift_down(heap, hn, 0);
            }
        }
        qsort(heap, (size_t)hn, sizeof(neighbor_t), canon_cmp_qsort);
    }
    for (int j = 0; j < k; ++j) {
        c->out_ids[(size_t)q * k + j] = j < hn ? heap[j].id : -1;
        c->out_dist[(size_t)q * k + j] = j < hn ? heap[j].dist : INFINITY;
    }
    if (c->out_count) c->out_count[q] = hn;
    free(heap);
}

int oracle_query_canonical(int metric, const float *corpus, const int64_t *ids, int64_t n, int d,
                           const float *queries, int b, int k, int64_t *out_ids, float *out_dist,
                           int32_t *out_count, int accum, int l2_squared, int nthreads) {
    if (k < 0) k = 0;
    canon_ctx_t c = {metric, d, k, accum, l2_squared, corpus, queries, ids, n, out_ids, out_dist, out_count};
    parallel_for(b, nthreads, canon_one_query, &c);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Reference-faithful index: one heap-allocated row per node in an insertion-ordered linked list */
/* (ConcurrentLinkedQueue[EntityEmbedding[T]], BruteForceIndex.scala:34-36,48-52).              */
typedef struct node_s {
    struct node_s *next;
    int64_t id;
    float *embedding; /* separately allocated, like Embedding[Float]'s backing array */
} node_t;

typedef struct {
    int metric;
    int d;
    int64_t n;
    node_t *head, *tail;
} oracle_index_t;

oracle_index_t *oracle_index_create(int metric, int d) {
    oracle_index_t *ix = (oracle_index_t *)calloc(1, sizeof(*ix));
    ix->metric = metric;
    ix->d = d;
    return ix;
}

/* append: linkedQueue.add(embedding), one row per call in the reference; batched here for speed  */
int oracle_index_append(oracle_index_t *ix, const int64_t *ids, const float *rows, int64_t n) {
    for (int64_t r = 0; r < n; ++r) {
        node_t *nd = (node_t *)malloc(sizeof(node_t));
        nd->embedding = (float *)malloc(sizeof(float) * (size_t)ix->d);
        memcpy(nd->embedding, rows + (size_t)r * ix->d, sizeof(float) * (size_t)ix->d);
        nd->id = ids ? ids[r] : ix->n;
        nd->next = NULL;
        if (ix->tail) ix->tail->next = nd;
        else ix->head = nd;
        ix->tail = nd;
        ix->n += 1;
    }
    return 0;
}

int64_t oracle_index_size(const oracle_index_t *ix) { return ix->n; }

void oracle_index_destroy(oracle_index_t *ix) {
    if (!ix) return;
    node_t *nd = ix->head;
    while (nd) {
        node_t *nx = nd->next;
        free(nd->embedding);
        free(nd);
        nd = nx;
    }
    free(ix);
}

/* Scala 2.12 mutable.PriorityQueue, restated: 1-indexed array heap, max on top under `ord`.     */
/* ord here is Ordering.by(_.distance): ONLY the distance takes part, ids never do (finding F6). */
static inline int pq_lt(const neighbor_t *x, const neighbor_t *y) { return float_compare(x->dist, y->dist) < 0; }
static inline int pq_ge(const neighbor_t *x, const neighbor_t *y) { return float_compare(x->dist, y->dist) >= 0; }

static void pq_fix_up(neighbor_t *as, int m) {
    int k = m;
    while (k > 1 && pq_lt(&as[k / 2], &as[k])) {
        neighbor_t t = as[k];
        as[k] = as[k / 2];
        as[k / 2] = t;
        k = k / 2;
    }
}

static void pq_fix_down(neighbor_t *as, int m, int n) {
    int k = m;
    while (n >= 2 * k) {
        int j = 2 * k;
        if (j < n && pq_lt(&as[j], &as[j + 1])) j += 1;
        if (pq_ge(&as[k], &as[j])) return;
        neighbor_t t = as[k];
        as[k] = as[j];
        as[j] = t;
        k = j;
    }
}

/*
 * queryWithDistance, BruteForceIndex.scala:66-91, line for line: walk the list, push EVERY row,
 * pop the max whenever size > k, finally dequeueAll (descending) and reverse.
 * k <= 0 yields an empty list (every push is immediately popped).
 */
typedef struct {
    const oracle_index_t *ix;
    const float *queries;
    int kk, accum, l2_squared;
    int64_t *out_ids;
    float *out_dist;
    int32_t *out_count;
} faithful_ctx_t;

static void faithful_one_query(int q, void *vctx) {
    faithful_ctx_t *c = (faithful_ctx_t *)vctx;
    const oracle_index_t *ix = c->ix;
    int kk = c->kk;
    const float *qv = c->queries + (size_t)q * ix->d;
    /* slot 0 unused; at most k+1 live entries */
    neighbor_t *as = (neighbor_t *)malloc(sizeof(neighbor_t) * (size_t)(kk + 3));
    int size0 = 1; /* resarr.p_size0: next free slot */
    int64_t seq = 0;
    for (const node_t *nd = ix->head; nd; nd = nd->next, ++seq) {
        neighbor_t nb;
        nb.id = nd->id;
        nb.dist = distance(ix->metric, nd->embedding, qv, ix->d, c->accum, c->l2_squared);
        nb.seq = seq;
        as[size0] = nb; /* += */
        pq_fix_up(as, size0);
        size0 += 1;
        if (size0 - 1 > kk) { /* dequeue() */
            size0 -= 1;
            as[1] = as[size0];
            pq_fix_down(as, 1, size0 - 1);
        }
    }
    int cnt = size0 - 1;
    /* dequeueAll: max first; .reverse => write from the back */
    for (int j = cnt - 1; j >= 0; --j) {
        neighbor_t top = as[1];
        size0 -= 1;
        as[1] = as[size0];
        pq_fix_down(as, 1, size0 - 1);
        c->out_ids[(size_t)q * kk + j] = top.id;
        c->out_dist[(size_t)q * kk + j] = top.dist;
    }
    for (int j = cnt; j < kk; ++j) {
        c->out_ids[(size_t)q * kk + j] = -1;
        c->out_dist[(size_t)q * kk + j] = INFINITY;
    }
    if (c->out_count) c->out_count[q] = cnt;
    free(as);
}

int oracle_index_query(const oracle_index_t *ix, const float *queries, int b, int k, int64_t *out_ids,
                       float *out_dist, int32_t *out_count, int accum, int l2_squared, int nthreads) {
    faithful_ctx_t c = {ix, queries, k < 0 ? 0 : k, accum, l2_squared, out_ids, out_dist, out_count};
    parallel_for(b, nthreads, faithful_one_query, &c);
    return 0;
}

/* Contiguous, unrolled, multi-threaded CPU scan under the canonical order: the "fair CPU" figure */
/* (BASELINE.md section 4).  fp32 lane-parallel accumulation for speed, so it is NOT a parity      */
/* oracle; it is only ever timed.                                                                  */
static void fast_one_query(int q, void *vctx) {
    canon_ctx_t *c = (canon_ctx_t *)vctx;
    int d = c->d, k = c->k, metric = c->metric;
    const float *qv = c->queries + (size_t)q * d;
    float nb2 = 0.f;
    for (int i = 0; i < d; ++i) nb2 += qv[i] * qv[i];
    neighbor_t *heap = (neighbor_t *)malloc(sizeof(neighbor_t) * (size_t)k);
    int hn = 0;
    for (int64_t r = 0; r < c->n; ++r) {
        const float *a = c->corpus + (size_t)r * d;
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, na[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        int i = 0;
        if (metric == ORACLE_L2) {
            for (; i + 8 <= d; i += 8)
                for (int j = 0; j < 8; ++j) {
                    float t = a[i + j] - qv[i + j];
                    acc[j] += t * t;
                }
            for (; i < d; ++i) {
                float t = a[i] - qv[i];
                acc[0] += t * t;
            }
        } else if (metric == ORACLE_INNER_PRODUCT) {
            for (; i + 8 <= d; i += 8)
                for (int j = 0; j < 8; ++j) acc[j] += a[i + j] * qv[i + j];
            for (; i < d; ++i) acc[0] += a[i] * qv[i];
        } else {
            for (; i + 8 <= d; i += 8)
                for (int j = 0; j < 8; ++j) {
                    acc[j] += a[i + j] * qv[i + j];
                    na[j] += a[i + j] * a[i + j];
                }
            for (; i < d; ++i) {
                acc[0] += a[i] * qv[i];
                na[0] += a[i] * a[i];
            }
        }
        float s = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
        float sa = ((na[0] + na[1]) + (na[2] + na[3])) + ((na[4] + na[5]) + (na[6] + na[7]));
        neighbor_t nb;
        nb.id = c->ids ? c->ids[r] : r;
        nb.seq = r;
        nb.dist = metric == ORACLE_L2 ? sqrtf(s)
                  : metric == ORACLE_INNER_PRODUCT ? 1.0f - s
                                                   : 1.0f - s / (sqrtf(sa) * sqrtf(nb2));
        if (hn < k) {
            heap[hn] = nb;
            canon_sift_up(heap, hn);
            ++hn;
        } else if (canonical_less(&nb, &heap[0])) {
            heap[0] = nb;
            canon_sift_down(heap, hn, 0);
        }
    }
    qsort(heap, (size_t)hn, sizeof(neighbor_t), canon_cmp_qsort);
    for (int j = 0; j < k; ++j) {
        c->out_ids[(size_t)q * k + j] = j < hn ? heap[j].id : -1;
        c->out_dist[(size_t)q * k + j] = j < hn ? heap[j].dist : INFINITY;
    }
    free(heap);
}

int oracle_query_fast_cpu(int metric, const float *corpus, const int64_t *ids, int64_t n, int d,
                          const float *queries, int b, int k, int64_t *out_ids, float *out_dist,
                          int nthreads) {
    if (k <= 0) return 0;
    canon_ctx_t c = {metric, d, k, 1, 0, corpus, queries, ids, n, out_ids, out_dist, NULL};
    parallel_for(b, nthreads, fast_one_query, &c);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* ComposedQueryable.queryWithDistance, ShardApi.scala:72-86.                                   */
/* in_* hold s shard lists of k slots each with in_count[s] valid entries.                       */
/* faithful=1: list.flatten.sorted(ordering).take(k): STABLE sort on distance only (shard order, */
/*             then intra-shard order, survives among ties).                                     */
/* faithful=0: canonical (distance, id) order -- what the multi-GPU merge must reproduce.         */
static int merge_cmp_stable(const void *a, const void *b) {
    const neighbor_t *x = (const neighbor_t *)a, *y = (const neighbor_t *)b;
    int c = float_compare(x->dist, y->dist);
    if (c) return c;
    return (x->seq > y->seq) - (x->seq < y->seq);
}

int oracle_merge(const int64_t *in_ids, const float *in_dist, const int32_t *in_count, int shards, int k,
                 int64_t *out_ids, float *out_dist, int32_t *out_count, int faithful) {
    if (k < 0) k = 0;
    neighbor_t *all = (neighbor_t *)malloc(sizeof(neighbor_t) * (size_t)(shards * k + 1));
    int m = 0;
    for (int s = 0; s < shards; ++s)
        for (int j = 0; j < in_count[s] && j < k; ++j) {
            all[m].id = in_ids[(size_t)s * k + j];
            all[m].dist = in_dist[(size_t)s * k + j];
            all[m].seq = m;
            ++m;
        }
    qsort(all, (size_t)m, sizeof(neighbor_t), faithful ? merge_cmp_stable : canon_cmp_qsort);
    int cnt = m < k ? m : k;
    for (int j = 0; j < k; ++j) {
        out_ids[j] = j < cnt ? all[j].id : -1;
        out_dist[j] = j < cnt ? all[j].dist : INFINITY;
    }
    if (out_count) *out_count = cnt;
    free(all);
    return 0;
}

