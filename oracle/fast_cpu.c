/*
 * fast_cpu.c -- the "fair CPU" figure of bench.py: what a competent CPU implementation of the same exact top-k would
 * do on the GPU box's host cores.  TIMED ONLY -- never a parity oracle (fp32 lane-parallel accumulation, -O3, fused
 * multiply-add allowed), and test infrastructure like the rest of oracle/: only tests/ and bench.py's cpu_baseline leg
 * may load it.
 *
 * Unlike the reference's loop (BruteForceIndex.scala:76-86: one thread walks the whole corpus per query, a heap push per
 * row) this is query-blocked and row-parallel:
 *   - threads split the ROWS (each streams its row range from DRAM once per query block, not once per query);
 *   - a block of QB queries is scored against a tile of rows that stays in L1/L2 (a small GEMM, SIMD over the dimension,
 *     8 independent accumulators per dot product);
 *   - a row enters a per-(thread, query) bounded heap only if it beats the heap's current worst;
 *   - per-thread heaps are merged at the end under the canonical (distance, id) order.
 * Function multiversioning picks AVX-512 / AVX2 / baseline code at load time on whatever host runs it.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#define FC_L2 0
#define FC_COSINE 1
#define FC_IP 2
#define QB 16      /* queries per block */
#define RT 64      /* rows per tile */

typedef struct {
    float dist;
    int64_t id;
} ent_t;

static inline int ent_less(const ent_t *a, const ent_t *b) {
    if (a->dist != b->dist) return a->dist < b->dist;
    return a->id < b->id;
}

static void sift_down(ent_t *h, int n, int i) {
    for (;;) {
        int l = 2 * i + 1, r = l + 1, m = i;
        if (l < n && ent_less(&h[m], &h[l])) m = l;
        if (r < n && ent_less(&h[m], &h[r])) m = r;
        if (m == i) return;
        ent_t t = h[i];
        h[i] = h[m];
        h[m] = t;
        i = m;
    }
}

static void sift_up(ent_t *h, int i) {
    while (i > 0) {
        int p = (i - 1) / 2;
        if (!ent_less(&h[p], &h[i])) return;
        ent_t t = h[i];
        h[i] = h[p];
        h[p] = t;
        i = p;
    }
}

/* scores[q][r] = a_r . b_q (and optionally |a_r|^2) for one tile; the compiler vectorises the inner loops */
__attribute__((target_clones("avx512f", "avx2,fma", "default")))
static void tile_scores(const float *restrict rows, int nr, int d, const float *restrict qs, int nq, int metric,
                        float *restrict out /* [QB][RT] */, float *restrict an2 /* [RT] */) {
    for (int r = 0; r < nr; ++r) {
        const float *a = rows + (size_t)r * d;
        if (metric == FC_COSINE) {
            float s = 0.f;
            for (int i = 0; i < d; ++i) s += a[i] * a[i];
            an2[r] = s;
        }
        int q = 0;
        for (; q + 4 <= nq; q += 4) {   /* four queries per pass over the row: four independent accumulator chains */
            const float *b0 = qs + (size_t)q * d, *b1 = b0 + d, *b2 = b1 + d, *b3 = b2 + d;
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
            if (metric == FC_L2) {
                for (int i = 0; i < d; ++i) {
                    const float ai = a[i];
                    const float t0 = ai - b0[i], t1 = ai - b1[i], t2 = ai - b2[i], t3 = ai - b3[i];
                    s0 += t0 * t0;
                    s1 += t1 * t1;
                    s2 += t2 * t2;
                    s3 += t3 * t3;
                }
            } else {
                for (int i = 0; i < d; ++i) {
                    const float ai = a[i];
                    s0 += ai * b0[i];
                    s1 += ai * b1[i];
                    s2 += ai * b2[i];
                    s3 += ai * b3[i];
                }
            }
            out[(q + 0) * RT + r] = s0;
            out[(q + 1) * RT + r] = s1;
            out[(q + 2) * RT + r] = s2;
            out[(q + 3) * RT + r] = s3;
        }
        for (; q < nq; ++q) {
            const float *b = qs + (size_t)q * d;
            float s = 0.f;
            if (metric == FC_L2) {
                for (int i = 0; i < d; ++i) {
                    float t = a[i] - b[i];
                    s += t * t;
                }
            } else {
                for (int i = 0; i < d; ++i) s += a[i] * b[i];
            }
            out[q * RT + r] = s;
        }
    }
}

typedef struct {
    int metric, d, k, nq, tid, nthreads;
    const float *corpus, *queries;
    const int64_t *ids;
    int64_t n;
    ent_t *heaps; /* [nthreads][nq][k] */
    int *heap_n;  /* [nthreads][nq] */
} ctx_t;

static void *worker(void *arg) {
    ctx_t *c = (ctx_t *)arg;
    const int d = c->d, k = c->k;
    const int64_t lo = c->n * c->tid / c->nthreads, hi = c->n * (c->tid + 1) / c->nthreads;
    float *sc = (float *)malloc(sizeof(float) * QB * RT);
    float an2[RT];
    float *qn = (float *)malloc(sizeof(float) * (size_t)c->nq);
    for (int q = 0; q < c->nq; ++q) {
        float s = 0.f;
        for (int i = 0; i < d; ++i) s += c->queries[(size_t)q * d + i] * c->queries[(size_t)q * d + i];
        qn[q] = sqrtf(s);
    }
    for (int q0 = 0; q0 < c->nq; q0 += QB) {
        const int nq = c->nq - q0 < QB ? c->nq - q0 : QB;
        for (int64_t r0 = lo; r0 < hi; r0 += RT) {
            const int nr = (int)(hi - r0 < RT ? hi - r0 : RT);
            tile_scores(c->corpus + (size_t)r0 * d, nr, d, c->queries + (size_t)q0 * d, nq, c->metric, sc, an2);
            for (int q = 0; q < nq; ++q) {
                ent_t *h = c->heaps + ((size_t)c->tid * c->nq + q0 + q) * k;
                int *hn = c->heap_n + (size_t)c->tid * c->nq + q0 + q;
                for (int r = 0; r < nr; ++r) {
                    float s = sc[q * RT + r], dist;
                    if (c->metric == FC_L2) dist = sqrtf(s);
                    else if (c->metric == FC_IP) dist = 1.0f - s;
                    else dist = 1.0f - s / (sqrtf(an2[r]) * qn[q0 + q]);
                    ent_t e = {dist, c->ids ? c->ids[r0 + r] : r0 + r};
                    if (*hn < k) {
                        h[*hn] = e;
                        sift_up(h, *hn);
                        ++*hn;
                    } else if (ent_less(&e, &h[0])) {
                        h[0] = e;
                        sift_down(h, k, 0);
                    }
                }
            }
        }
    }
    free(sc);
    free(qn);
    return NULL;
}

static int cmp_ent(const void *a, const void *b) {
    const ent_t *x = (const ent_t *)a, *y = (const ent_t *)b;
    if (ent_less(x, y)) return -1;
    if (ent_less(y, x)) return 1;
    return 0;
}

int fastcpu_query(int metric, const float *corpus, const int64_t *ids, int64_t n, int d, const float *queries, int nq, int k,
                  int64_t *out_ids, float *out_dist, int nthreads) {
    if (k <= 0 || nq <= 0) return 0;
    if (nthreads <= 0) nthreads = (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (nthreads < 1) nthreads = 1;
    if ((int64_t)nthreads > n) nthreads = n > 0 ? (int)n : 1;
    ent_t *heaps = (ent_t *)malloc(sizeof(ent_t) * (size_t)nthreads * nq * k);
    int *heap_n = (int *)calloc((size_t)nthreads * nq, sizeof(int));
    ctx_t *cs = (ctx_t *)malloc(sizeof(ctx_t) * (size_t)nthreads);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    for (int t = 0; t < nthreads; ++t) {
        ctx_t c = {metric, d, k, nq, t, nthreads, corpus, queries, ids, n, heaps, heap_n};
        cs[t] = c;
        if (pthread_create(&th[t], NULL, worker, &cs[t]) != 0) {
            worker(&cs[t]);
            th[t] = 0;
        }
    }
    for (int t = 0; t < nthreads; ++t)
        if (th[t]) pthread_join(th[t], NULL);
    ent_t *all = (ent_t *)malloc(sizeof(ent_t) * (size_t)nthreads * k);
    for (int q = 0; q < nq; ++q) {
        int m = 0;
        for (int t = 0; t < nthreads; ++t) {
            int hn = heap_n[(size_t)t * nq + q];
            memcpy(all + m, heaps + ((size_t)t * nq + q) * k, sizeof(ent_t) * (size_t)hn);
            m += hn;
        }
        qsort(all, (size_t)m, sizeof(ent_t), cmp_ent);
        for (int j = 0; j < k; ++j) {
            out_ids[(size_t)q * k + j] = j < m ? all[j].id : -1;
            out_dist[(size_t)q * k + j] = j < m ? all[j].dist : INFINITY;
        }
    }
    free(all);
    free(th);
    free(cs);
    free(heap_n);
    free(heaps);
    return 0;
}
