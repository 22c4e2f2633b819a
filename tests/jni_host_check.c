/* jni_host_check.c -- runs every Java_com_twitter_ann_brute_1force_B200AnnNative_00024_* entry point of the JNI shim
 * (the-algorithm_b200/host/jni/b200ann_jni.c) against a MOCK JNIEnv, so the boundary a JVM would call is compiled and
 * executed although this image has no JDK.  The shim is compiled against tests/jni_stub/jni.h (a stand-in declaring the JNI
 * members the shim uses); direct ByteBuffers, int arrays and Strings are plain C memory behind `struct _jobject`.
 *
 *   gcc -std=gnu11 -Itests/jni_stub tests/jni_host_check.c -Lthe-algorithm_b200/lib -lb200ann ... -o jni_host_check
 *   ./jni_host_check <scratch directory>
 *
 * Without a GPU: `create` must return 0 and `lastError` must say why (no CPU fallback) -> prints "no device".
 * With a GPU: known answers through appendBatch / queryBatch / updateBatch / readRows / distancePairs / normalizeRows /
 * knnJoin, the sharded handle against the single index, a save/load round trip in the reference's directory format, and the
 * shim's own refusal of an undersized direct buffer -> prints "gpu ok".  Any failure returns a distinct exit code.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../the-algorithm_b200/host/jni/b200ann_jni.c"

#ifndef B200ANN_HAVE_JNI
#error "compile with -Itests/jni_stub so that the shim's body is built"
#endif

/* ---- mock JVM objects ------------------------------------------------------------------------------------------ */
struct _jobject {
    void *addr;      /* direct buffer address / int array elements */
    jlong capacity;  /* direct buffer capacity in bytes; -1 = not a direct buffer */
    jsize length;    /* array length */
    const char *utf; /* string contents */
};

static void *m_GetDirectBufferAddress(JNIEnv *env, jobject buf) { (void)env; return buf->capacity < 0 ? NULL : buf->addr; }
static jlong m_GetDirectBufferCapacity(JNIEnv *env, jobject buf) { (void)env; return buf->capacity; }
static jsize m_GetArrayLength(JNIEnv *env, jarray a) { (void)env; return a->length; }
static int g_int_arrays_out = 0, g_strings_out = 0;
static jint *m_GetIntArrayElements(JNIEnv *env, jintArray a, jboolean *is_copy) {
    (void)env;
    if (is_copy) *is_copy = JNI_FALSE;
    g_int_arrays_out++;
    return (jint *)a->addr;
}
static void m_ReleaseIntArrayElements(JNIEnv *env, jintArray a, jint *elems, jint mode) {
    (void)env; (void)a; (void)elems; (void)mode;
    g_int_arrays_out--;
}
static const char *m_GetStringUTFChars(JNIEnv *env, jstring s, jboolean *is_copy) {
    (void)env;
    if (is_copy) *is_copy = JNI_FALSE;
    g_strings_out++;
    return s->utf;
}
static void m_ReleaseStringUTFChars(JNIEnv *env, jstring s, const char *chars) {
    (void)env; (void)s; (void)chars;
    g_strings_out--;
}
static struct _jobject g_new_string;
static char g_new_string_text[1024];
static jstring m_NewStringUTF(JNIEnv *env, const char *utf) {
    (void)env;
    snprintf(g_new_string_text, sizeof(g_new_string_text), "%s", utf ? utf : "");
    g_new_string.utf = g_new_string_text;
    return &g_new_string;
}

static const struct JNINativeInterface_ g_table = {
    m_GetDirectBufferAddress, m_GetDirectBufferCapacity, m_GetArrayLength,       m_GetIntArrayElements,
    m_ReleaseIntArrayElements, m_GetStringUTFChars,      m_ReleaseStringUTFChars, m_NewStringUTF,
};
static JNIEnv g_env = &g_table;
#define ENV (&g_env)
#define J(name) Java_com_twitter_ann_brute_1force_B200AnnNative_00024_##name

static struct _jobject direct(void *p, size_t bytes) {
    struct _jobject o = {p, (jlong)bytes, 0, NULL};
    return o;
}
static struct _jobject jstr(const char *s) {
    struct _jobject o = {NULL, -1, 0, s};
    return o;
}
static struct _jobject jints(jint *v, int n) {
    struct _jobject o = {v, -1, n, NULL};
    return o;
}
static const char *last_error(void) { return J(lastError)(ENV, NULL)->utf; }

#define CHECK(cond, code)                                                              \
    do {                                                                               \
        if (!(cond)) {                                                                 \
            printf("FAILED %s (line %d): %s\n", #cond, __LINE__, last_error());        \
            return (code);                                                             \
        }                                                                              \
    } while (0)

int main(int argc, char **argv) {
    const char *scratch = argc > 1 ? argv[1] : "/tmp";
    CHECK(J(version)(ENV, NULL) == ann_version(), 2);

    /* BruteForceIndex.apply through the shim: InnerProduct, dim 2 */
    jlong h = J(create)(ENV, NULL, ANN_METRIC_INNER_PRODUCT, 2, 0, 0, 0);
    if (h == 0) {
        const char *why = last_error();
        if (strstr(why, "no CUDA device") || strstr(why, "sm_100a only") || strstr(why, "failed")) {
            /* every other handle-taking entry point must refuse a null handle instead of crashing */
            float q[2] = {1, 0};
            int64_t ids[1];
            float dist[1];
            int32_t cnt[1];
            struct _jobject bq = direct(q, sizeof q), bi = direct(ids, sizeof ids), bd = direct(dist, sizeof dist), bc = direct(cnt, sizeof cnt);
            CHECK(J(queryBatch)(ENV, NULL, 0, &bq, 1, 2, 1, &bi, &bd, &bc) == ANN_ERR_NULL_POINTER, 3);
            CHECK(J(appendBatch)(ENV, NULL, 0, &bi, &bq, 1) == ANN_ERR_NULL_POINTER, 4);
            CHECK(J(size)(ENV, NULL, 0) == -1, 5);
            J(destroy)(ENV, NULL, 0);
            jint dev[2] = {0, 0};
            struct _jobject jd = jints(dev, 2);
            CHECK(J(shardedCreate)(ENV, NULL, ANN_METRIC_L2, 3, 0, 0, &jd) == 0 && g_int_arrays_out == 0, 6);
            struct _jobject dir = jstr(scratch);
            CHECK(J(loadDirectory)(ENV, NULL, ANN_METRIC_L2, 0, 0, 0, &dir, 0) == 0 && g_strings_out == 0, 7);
            printf("no device: %s\n", why);
            return 0;
        }
        printf("unexpected create failure: %s\n", why);
        return 8;
    }

    /* ---- Appendable / Queryable known answer (the one tests/cpp_host_check.cpp uses) ---- */
    int64_t ids3[3] = {7, 8, 9};
    float rows3[6] = {1, 0, 0, 1, 2, 0};
    struct _jobject b_ids3 = direct(ids3, sizeof ids3), b_rows3 = direct(rows3, sizeof rows3);
    CHECK(J(appendBatch)(ENV, NULL, h, &b_ids3, &b_rows3, 3) == ANN_OK, 10);
    CHECK(J(size)(ENV, NULL, h) == 3, 11);
    float q1[2] = {1, 0};
    int64_t out_ids[2];
    float out_dist[2];
    int32_t out_cnt[1];
    struct _jobject b_q1 = direct(q1, sizeof q1), b_oi = direct(out_ids, sizeof out_ids), b_od = direct(out_dist, sizeof out_dist),
                    b_oc = direct(out_cnt, sizeof out_cnt);
    CHECK(J(queryBatch)(ENV, NULL, h, &b_q1, 1, 2, 2, &b_oi, &b_od, &b_oc) == ANN_OK, 12);
    CHECK(out_cnt[0] == 2 && out_ids[0] == 9 && out_dist[0] == -1.0f && out_ids[1] == 7 && out_dist[1] == 0.0f, 13);

    /* the shim refuses buffers that are too small for the call, before the library could overrun them */
    struct _jobject b_small = direct(out_ids, 8);
    CHECK(J(queryBatch)(ENV, NULL, h, &b_q1, 1, 2, 2, &b_small, &b_od, &b_oc) == ANN_ERR_INVALID_ARGUMENT, 14);
    CHECK(strstr(last_error(), "outIds") != NULL, 15);
    struct _jobject b_heap = {out_ids, -1, 0, NULL}; /* a heap (non-direct) ByteBuffer has no address */
    CHECK(J(queryBatch)(ENV, NULL, h, &b_q1, 1, 2, 2, &b_heap, &b_od, &b_oc) == ANN_ERR_INVALID_ARGUMENT, 16);
    CHECK(J(appendBatch)(ENV, NULL, h, &b_ids3, &b_small, 3) == ANN_ERR_INVALID_ARGUMENT && J(size)(ENV, NULL, h) == 3, 17);
    /* library-side errors keep their codes and messages */
    float q3[3] = {1, 0, 0};
    struct _jobject b_q3 = direct(q3, sizeof q3);
    CHECK(J(queryBatch)(ENV, NULL, h, &b_q3, 1, 3, 2, &b_oi, &b_od, &b_oc) == ANN_ERR_DIMENSION_MISMATCH, 18);
    CHECK(strstr(last_error(), "dimension") != NULL, 19);

    /* ---- Updatable.update: slot 0 (id 7) becomes (5, 0) and is now the nearest under InnerProduct ---- */
    int64_t slot0[1] = {0};
    float new_row[2] = {5, 0};
    struct _jobject b_slot = direct(slot0, sizeof slot0), b_new = direct(new_row, sizeof new_row);
    CHECK(J(updateBatch)(ENV, NULL, h, &b_slot, &b_new, 1) == ANN_OK, 20);
    CHECK(J(queryBatch)(ENV, NULL, h, &b_q1, 1, 2, 2, &b_oi, &b_od, &b_oc) == ANN_OK, 21);
    CHECK(out_ids[0] == 7 && out_dist[0] == -4.0f && out_ids[1] == 9, 22);
    int64_t rd_ids[3];
    float rd_rows[6];
    struct _jobject b_ri = direct(rd_ids, sizeof rd_ids), b_rr = direct(rd_rows, sizeof rd_rows);
    CHECK(J(readRows)(ENV, NULL, h, 0, 3, &b_ri, &b_rr) == ANN_OK, 23);
    CHECK(rd_ids[0] == 7 && rd_ids[2] == 9 && rd_rows[0] == 5.0f && rd_rows[4] == 2.0f, 24);

    /* ---- options / stats by name ---- */
    struct _jobject s_dim = jstr("dim"), s_path = jstr("path"), s_bogus = jstr("no_such_option");
    CHECK(J(getStat)(ENV, NULL, h, &s_dim) == 2, 25);
    CHECK(J(setOption)(ENV, NULL, h, &s_path, 1) == ANN_OK && J(setOption)(ENV, NULL, h, &s_path, 0) == ANN_OK, 26);
    CHECK(J(setOption)(ENV, NULL, h, &s_bogus, 1) == ANN_ERR_UNKNOWN_OPTION && g_strings_out == 0, 27);

    /* ---- Metric.distance / MetricUtil.norm for plain vectors ---- */
    float pa[4] = {0, 0, 1, 2}, pb[4] = {3, 4, 4, 5}, pout[2];
    struct _jobject b_pa = direct(pa, sizeof pa), b_pb = direct(pb, sizeof pb), b_po = direct(pout, sizeof pout);
    CHECK(J(distancePairs)(ENV, NULL, ANN_METRIC_L2, 0, 2, &b_pa, &b_pb, 2, &b_po, 0) == ANN_OK, 30);
    CHECK(pout[0] == 5.0f && pout[1] == (float)sqrt(18.0), 31);
    float nr[2] = {3, 4}, nout[2];
    struct _jobject b_nr = direct(nr, sizeof nr), b_no = direct(nout, sizeof nout);
    CHECK(J(normalizeRows)(ENV, NULL, 2, &b_nr, 1, &b_no, 0) == ANN_OK && nout[0] == 0.6f && nout[1] == 0.8f, 32);

    /* ---- a larger corpus: sharded handle (two shards on device 0) and knnJoin against the single index ---- */
    enum { N = 5000, D = 3, B = 2, K = 10 };
    static int64_t sids[N];
    static float srows[N * D];
    for (int i = 0; i < N; ++i) {
        sids[i] = 100000 - i;
        srows[i * D + 0] = (float)(i % 71);
        srows[i * D + 1] = (float)(i % 13);
        srows[i * D + 2] = (float)(i / 1000);
    }
    struct _jobject b_sids = direct(sids, sizeof sids), b_srows = direct(srows, sizeof srows);
    jlong one = J(create)(ENV, NULL, ANN_METRIC_L2, D, 0, 0, 0);
    CHECK(one != 0 && J(appendBatch)(ENV, NULL, one, &b_sids, &b_srows, N) == ANN_OK, 40);
    jint dev[2] = {0, 0};
    struct _jobject jd = jints(dev, 2);
    jlong sx = J(shardedCreate)(ENV, NULL, ANN_METRIC_L2, D, 0, 0, &jd);
    CHECK(sx != 0 && g_int_arrays_out == 0, 41);
    CHECK(J(shardedAppendBatch)(ENV, NULL, sx, &b_sids, &b_srows, N) == ANN_OK && J(shardedSize)(ENV, NULL, sx) == N, 42);
    float qs[B * D] = {3.f, 4.f, 1.f, 70.f, 12.f, 4.f};
    int64_t a_ids[B * K], s_ids[B * K], j_ids[B * K];
    float a_d[B * K], s_d[B * K], j_d[B * K];
    int32_t a_c[B], s_c[B], j_c[B];
    struct _jobject b_qs = direct(qs, sizeof qs);
    struct _jobject b_ai = direct(a_ids, sizeof a_ids), b_ad = direct(a_d, sizeof a_d), b_ac = direct(a_c, sizeof a_c);
    struct _jobject b_si = direct(s_ids, sizeof s_ids), b_sd = direct(s_d, sizeof s_d), b_sc = direct(s_c, sizeof s_c);
    struct _jobject b_ji = direct(j_ids, sizeof j_ids), b_jd = direct(j_d, sizeof j_d), b_jc = direct(j_c, sizeof j_c);
    CHECK(J(queryBatch)(ENV, NULL, one, &b_qs, B, D, K, &b_ai, &b_ad, &b_ac) == ANN_OK, 43);
    CHECK(J(shardedQueryBatch)(ENV, NULL, sx, &b_qs, B, D, K, &b_si, &b_sd, &b_sc) == ANN_OK, 44);
    CHECK(J(knnJoin)(ENV, NULL, ANN_METRIC_L2, D, 0, 0, &b_sids, &b_srows, N, &b_qs, B, K, 2048, 0, &b_ji, &b_jd, &b_jc) == ANN_OK, 45);
    CHECK(a_c[0] == K && a_c[1] == K, 46);
    CHECK(memcmp(a_ids, s_ids, sizeof a_ids) == 0 && memcmp(a_d, s_d, sizeof a_d) == 0 && memcmp(a_c, s_c, sizeof a_c) == 0, 47);
    CHECK(memcmp(a_ids, j_ids, sizeof a_ids) == 0 && memcmp(a_d, j_d, sizeof a_d) == 0 && memcmp(a_c, j_c, sizeof a_c) == 0, 48);
    /* exact ties are ordered by id: every (distance, id) pair must ascend */
    for (int q = 0; q < B; ++q)
        for (int j = 1; j < K; ++j) {
            const float d0 = a_d[q * K + j - 1], d1 = a_d[q * K + j];
            CHECK(d0 < d1 || (d0 == d1 && a_ids[q * K + j - 1] < a_ids[q * K + j]), 49);
        }

    /* ---- the reference's directory format: save, load, same answer; sharded save/load likewise ---- */
    char dir1[900], dir2[900];
    snprintf(dir1, sizeof dir1, "%s/jni_single", scratch);
    snprintf(dir2, sizeof dir2, "%s/jni_sharded", scratch);
    struct _jobject s_dir1 = jstr(dir1), s_dir2 = jstr(dir2);
    CHECK(J(saveDirectory)(ENV, NULL, one, &s_dir1, ANN_ID_INT64_BE, ANN_LAYOUT_FLOAT_TENSOR) == ANN_OK, 50);
    jlong back = J(loadDirectory)(ENV, NULL, ANN_METRIC_L2, 0, 0, 0, &s_dir1, ANN_ID_AUTO);
    CHECK(back != 0 && J(size)(ENV, NULL, back) == N, 51);
    CHECK(J(queryBatch)(ENV, NULL, back, &b_qs, B, D, K, &b_si, &b_sd, &b_sc) == ANN_OK, 52);
    CHECK(memcmp(a_ids, s_ids, sizeof a_ids) == 0 && memcmp(a_d, s_d, sizeof a_d) == 0, 53);
    CHECK(J(shardedSaveDirectory)(ENV, NULL, sx, &s_dir2, ANN_ID_INT64_BE, ANN_LAYOUT_FLOAT_TENSOR) == ANN_OK, 54);
    jlong sback = J(shardedLoadDirectory)(ENV, NULL, ANN_METRIC_L2, 0, 0, &s_dir2, ANN_ID_AUTO, &jd);
    CHECK(sback != 0 && J(shardedSize)(ENV, NULL, sback) == N && g_int_arrays_out == 0 && g_strings_out == 0, 55);
    CHECK(J(shardedQueryBatch)(ENV, NULL, sback, &b_qs, B, D, K, &b_si, &b_sd, &b_sc) == ANN_OK, 56);
    CHECK(memcmp(a_ids, s_ids, sizeof a_ids) == 0 && memcmp(a_d, s_d, sizeof a_d) == 0, 57);

    J(shardedDestroy)(ENV, NULL, sback);
    J(shardedDestroy)(ENV, NULL, sx);
    J(destroy)(ENV, NULL, back);
    J(destroy)(ENV, NULL, one);
    J(destroy)(ENV, NULL, h);
    J(destroy)(ENV, NULL, 0); /* NULL handle is ignored, like swig Index.delete() on a released object */
    printf("gpu ok\n");
    return 0;
}
