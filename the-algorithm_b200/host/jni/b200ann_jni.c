/* b200ann_jni.c -- JNI shim between com.twitter.ann.brute_force.B200AnnNative (host/scala/GpuBruteForceIndex.scala) and
 * the C ABI (include/b200ann.h).  One function per entry point and one piece of logic: direct ByteBuffers are unwrapped to
 * plain pointers AFTER their capacity has been checked against what the call will read or write, so a buffer sized for the
 * wrong (b, k, dim) is refused (ANN_ERR_INVALID_ARGUMENT, message through lastError) instead of overrunning the JVM's
 * native heap.  Same boundary shape as the reference's SWIG Faiss binding
 * (ann/src/main/java/com/twitter/ann/faiss/swig/swigfaissJNI.java:267-269: opaque long handle + flat buffers).
 *
 * Build (where a JDK exists):  gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux \
 *                                  b200ann_jni.c -L../../lib -lb200ann -o libb200ann_jni.so
 * This image has no jni.h, so the body is compiled only when the header is present; the guard keeps `make` green.
 * tests/test_jni_host.py compiles the body against a stand-in header (tests/jni_stub/jni.h) and runs every entry point
 * below against a mock JNIEnv (tests/jni_host_check.c). */
#if defined(__has_include)
#if __has_include(<jni.h>)
#define B200ANN_HAVE_JNI 1
#endif
#endif

#ifdef B200ANN_HAVE_JNI
#include <jni.h>

#include "../../../include/b200ann.h"

#define NATIVE(ret, name) JNIEXPORT ret JNICALL Java_com_twitter_ann_brute_1force_B200AnnNative_00024_##name

static void *addr(JNIEnv *env, jobject buf) { return buf ? (*env)->GetDirectBufferAddress(env, buf) : NULL; }

/* message of the LAST native call on this thread when the shim itself refused it; every native clears it on entry, so
 * lastError reports the shim's message only for the call that produced it and the library's (ann_last_error) otherwise */
static __thread const char *g_shim_error = NULL;

/* 1 when `buf` is a direct buffer with room for `bytes` (a NULL buffer passes: the C ABI decides whether it is optional) */
static int fits(JNIEnv *env, jobject buf, int64_t bytes, const char *what) {
    if (!buf || bytes <= 0) return 1;
    if ((*env)->GetDirectBufferAddress(env, buf) != NULL && (int64_t)(*env)->GetDirectBufferCapacity(env, buf) >= bytes) return 1;
    g_shim_error = what;
    return 0;
}
#define REQUIRE(buf, bytes, what) \
    do { if (!fits(env, (buf), (int64_t)(bytes), "b200ann_jni: " what " is not a direct ByteBuffer large enough for this call")) return ANN_ERR_INVALID_ARGUMENT; } while (0)
/* dimension of an existing handle (0 when it cannot be read: the C ABI then reports the real problem) */
static int64_t dim_of(jlong handle) {
    int64_t d = 0;
    return (handle && ann_get_stat((const ann_index *)(intptr_t)handle, "dim", &d) == ANN_OK) ? d : 0;
}
static int64_t pos(int64_t v) { return v > 0 ? v : 0; }

NATIVE(jlong, create)(JNIEnv *env, jobject self, jint metric, jint dim, jlong capacity_hint, jint device, jint flags) {
    g_shim_error = NULL;
    (void)env; (void)self;
    ann_config cfg = {metric, dim, capacity_hint, device, (uint32_t)flags};
    ann_index *ix = NULL;
    return ann_create(&cfg, &ix) == ANN_OK ? (jlong)(intptr_t)ix : 0;
}

NATIVE(void, destroy)(JNIEnv *env, jobject self, jlong handle) {
    g_shim_error = NULL;
    (void)env; (void)self;
    ann_destroy((ann_index *)(intptr_t)handle);
}

NATIVE(jint, appendBatch)(JNIEnv *env, jobject self, jlong handle, jobject ids, jobject rows, jlong n) {
    g_shim_error = NULL;
    (void)self;
    REQUIRE(ids, pos(n) * 8, "ids");
    REQUIRE(rows, pos(n) * dim_of(handle) * 4, "rows");
    return ann_append_batch((ann_index *)(intptr_t)handle, (const int64_t *)addr(env, ids), (const float *)addr(env, rows), n);
}

NATIVE(jlong, size)(JNIEnv *env, jobject self, jlong handle) {
    g_shim_error = NULL;
    (void)env; (void)self;
    int64_t n = 0;
    return ann_size((const ann_index *)(intptr_t)handle, &n) == ANN_OK ? (jlong)n : -1;
}

NATIVE(jint, queryBatch)(JNIEnv *env, jobject self, jlong handle, jobject queries, jint b, jint dim, jint k, jobject out_ids,
                         jobject out_dist, jobject out_count) {
    g_shim_error = NULL;
    (void)self;
    REQUIRE(queries, pos(b) * pos(dim) * 4, "queries");
    REQUIRE(out_ids, pos(b) * pos(k) * 8, "outIds");
    REQUIRE(out_dist, pos(b) * pos(k) * 4, "outDist");
    REQUIRE(out_count, pos(b) * 4, "outCount");
    return ann_query_batch((ann_index *)(intptr_t)handle, (const float *)addr(env, queries), b, dim, k,
                           (int64_t *)addr(env, out_ids), (float *)addr(env, out_dist), (int32_t *)addr(env, out_count));
}

/* KnnHelper.findNearestNeighbours in one native call (ann_knn_join): host buffers in, host buffers out */
NATIVE(jint, knnJoin)(JNIEnv *env, jobject self, jint metric, jint dim, jint device, jint flags, jobject corpus_ids,
                      jobject corpus_rows, jlong n, jobject queries, jlong nq, jint k, jlong corpus_tile_rows, jint query_tile,
                      jobject out_ids, jobject out_dist, jobject out_count) {
    g_shim_error = NULL;
    (void)self;
    REQUIRE(corpus_ids, pos(n) * 8, "corpusIds");
    REQUIRE(corpus_rows, pos(n) * pos(dim) * 4, "corpusRows");
    REQUIRE(queries, pos(nq) * pos(dim) * 4, "queries");
    REQUIRE(out_ids, pos(nq) * pos(k) * 8, "outIds");
    REQUIRE(out_dist, pos(nq) * pos(k) * 4, "outDist");
    REQUIRE(out_count, pos(nq) * 4, "outCount");
    ann_config cfg = {metric, dim, 0, device, (uint32_t)flags};
    return ann_knn_join(&cfg, (const int64_t *)addr(env, corpus_ids), (const float *)addr(env, corpus_rows), n,
                        (const float *)addr(env, queries), nq, k, corpus_tile_rows, query_tile, (int64_t *)addr(env, out_ids),
                        (float *)addr(env, out_dist), (int32_t *)addr(env, out_count));
}

/* Metric.distance for n plain pairs / MetricUtil.norm (Metric.scala:76-86, 285-289): host buffers, computed on the device */
NATIVE(jint, distancePairs)(JNIEnv *env, jobject self, jint metric, jint flags, jint dim, jobject a, jobject b, jlong n,
                            jobject out, jint device) {
    g_shim_error = NULL;
    (void)self;
    REQUIRE(a, pos(n) * pos(dim) * 4, "a");
    REQUIRE(b, pos(n) * pos(dim) * 4, "b");
    REQUIRE(out, pos(n) * 4, "out");
    return ann_distance_pairs(metric, (uint32_t)flags, dim, (const float *)addr(env, a), (const float *)addr(env, b), n,
                              (float *)addr(env, out), device);
}

NATIVE(jint, normalizeRows)(JNIEnv *env, jobject self, jint dim, jobject rows, jlong n, jobject out, jint device) {
    g_shim_error = NULL;
    (void)self;
    REQUIRE(rows, pos(n) * pos(dim) * 4, "rows");
    REQUIRE(out, pos(n) * pos(dim) * 4, "out");
    return ann_normalize_rows(dim, (const float *)addr(env, rows), n, (float *)addr(env, out), device);
}

/* The shard-side query in two halves (ann_query_seed_device / ann_query_finish_device).  Device addresses travel as
 * longs: the query batch, the key arrays and the result block live in device memory the host mapped across its GPUs
 * (CUDA IPC); peerSeedKeys is a direct buffer holding `world` 64-bit device addresses. */
NATIVE(jint, querySeedDevice)(JNIEnv *env, jobject self, jlong handle, jlong d_queries, jint b, jint dim, jint k,
                              jlong d_seed_keys, jlong stream) {
    g_shim_error = NULL;
    (void)env; (void)self;
    return ann_query_seed_device((ann_index *)(intptr_t)handle, (const float *)(intptr_t)d_queries, b, dim, k,
                                 (uint32_t *)(intptr_t)d_seed_keys, (void *)(intptr_t)stream);
}

NATIVE(jint, queryFinishDevice)(JNIEnv *env, jobject self, jlong handle, jlong d_queries, jint b, jint dim, jint k,
                                jobject peer_seed_keys, jint world, jlong d_out_ids, jlong d_out_dist, jlong d_out_count,
                                jlong stream) {
    g_shim_error = NULL;
    (void)self;
    REQUIRE(peer_seed_keys, pos(world) * 8, "peerSeedKeys");
    return ann_query_finish_device((ann_index *)(intptr_t)handle, (const float *)(intptr_t)d_queries, b, dim, k,
                                   (const uint32_t *const *)addr(env, peer_seed_keys), world, (int64_t *)(intptr_t)d_out_ids,
                                   (float *)(intptr_t)d_out_dist, (int32_t *)(intptr_t)d_out_count, (void *)(intptr_t)stream);
}

/* ---- one process, several GPUs: ShardedAppendable + ComposedQueryable as one native handle (ann_sharded_*) ---- */
NATIVE(jlong, shardedCreate)(JNIEnv *env, jobject self, jint metric, jint dim, jlong capacity_hint, jint flags, jintArray devices) {
    g_shim_error = NULL;
    (void)self;
    ann_config cfg = {metric, dim, capacity_hint, 0, (uint32_t)flags};
    jsize n = (*env)->GetArrayLength(env, devices);
    jint *d = (*env)->GetIntArrayElements(env, devices, NULL);
    ann_sharded_index *sx = NULL;
    int rc = ann_sharded_create(&cfg, (const int32_t *)d, (int32_t)n, &sx);
    (*env)->ReleaseIntArrayElements(env, devices, d, JNI_ABORT);
    return rc == ANN_OK ? (jlong)(intptr_t)sx : 0;
}

NATIVE(void, shardedDestroy)(JNIEnv *env, jobject self, jlong handle) {
    g_shim_error = NULL;
    (void)env; (void)self;
    ann_sharded_destroy((ann_sharded_index *)(intptr_t)handle);
}

NATIVE(jint, shardedAppendBatch)(JNIEnv *env, jobject self, jlong handle, jobject ids, jobject rows, jlong n) {
    g_shim_error = NULL;
    (void)self;
    int64_t d = 0;
    if (handle && ann_sharded_get_stat((const ann_sharded_index *)(intptr_t)handle, "dim", &d) != ANN_OK) d = 0;
    REQUIRE(ids, pos(n) * 8, "ids");
    REQUIRE(rows, pos(n) * d * 4, "rows");
    return ann_sharded_append_batch((ann_sharded_index *)(intptr_t)handle, (const int64_t *)addr(env, ids),
                                    (const float *)addr(env, rows), n);
}

NATIVE(jlong, shardedSize)(JNIEnv *env, jobject self, jlong handle) {
    g_shim_error = NULL;
    (void)env; (void)self;
    int64_t n = 0;
    return ann_sharded_size((const ann_sharded_index *)(intptr_t)handle, &n) == ANN_OK ? (jlong)n : -1;
}

NATIVE(jint, shardedQueryBatch)(JNIEnv *env, jobject self, jlong handle, jobject queries, jint b, jint dim, jint k,
                                jobject out_ids, jobject out_dist, jobject out_count) {
    g_shim_error = NULL;
    (void)self;
    REQUIRE(queries, pos(b) * pos(dim) * 4, "queries");
    REQUIRE(out_ids, pos(b) * pos(k) * 8, "outIds");
    REQUIRE(out_dist, pos(b) * pos(k) * 4, "outDist");
    REQUIRE(out_count, pos(b) * 4, "outCount");
    return ann_sharded_query_batch((ann_sharded_index *)(intptr_t)handle, (const float *)addr(env, queries), b, dim, k,
                                   (int64_t *)addr(env, out_ids), (float *)addr(env, out_dist), (int32_t *)addr(env, out_count));
}

/* ---- the reference's on-disk format: BruteForceFileData thrift stream, shard_<i>/ directories (csrc/persist.cu) ---- */
NATIVE(jint, saveDirectory)(JNIEnv *env, jobject self, jlong handle, jstring dir, jint id_format, jint layout) {
    g_shim_error = NULL;
    (void)self;
    const char *d = (*env)->GetStringUTFChars(env, dir, NULL);
    int rc = ann_save_directory((ann_index *)(intptr_t)handle, d, id_format, layout);
    (*env)->ReleaseStringUTFChars(env, dir, d);
    return rc;
}

NATIVE(jlong, loadDirectory)(JNIEnv *env, jobject self, jint metric, jint dim, jint device, jint flags, jstring dir, jint id_format) {
    g_shim_error = NULL;
    (void)self;
    ann_config cfg = {metric, dim, 0, device, (uint32_t)flags};
    const char *d = (*env)->GetStringUTFChars(env, dir, NULL);
    ann_index *ix = NULL;
    int rc = ann_load_directory(&cfg, d, id_format, &ix);
    (*env)->ReleaseStringUTFChars(env, dir, d);
    return rc == ANN_OK ? (jlong)(intptr_t)ix : 0;
}

NATIVE(jint, shardedSaveDirectory)(JNIEnv *env, jobject self, jlong handle, jstring dir, jint id_format, jint layout) {
    g_shim_error = NULL;
    (void)self;
    const char *d = (*env)->GetStringUTFChars(env, dir, NULL);
    int rc = ann_sharded_save_directory((ann_sharded_index *)(intptr_t)handle, d, id_format, layout);
    (*env)->ReleaseStringUTFChars(env, dir, d);
    return rc;
}

NATIVE(jlong, shardedLoadDirectory)(JNIEnv *env, jobject self, jint metric, jint dim, jint flags, jstring dir, jint id_format,
                                    jintArray devices) {
    g_shim_error = NULL;
    (void)self;
    ann_config cfg = {metric, dim, 0, 0, (uint32_t)flags};
    const char *d = (*env)->GetStringUTFChars(env, dir, NULL);
    jsize n = (*env)->GetArrayLength(env, devices);
    jint *dv = (*env)->GetIntArrayElements(env, devices, NULL);
    ann_sharded_index *sx = NULL;
    int rc = ann_sharded_load_directory(&cfg, d, id_format, (const int32_t *)dv, (int32_t)n, &sx);
    (*env)->ReleaseIntArrayElements(env, devices, dv, JNI_ABORT);
    (*env)->ReleaseStringUTFChars(env, dir, d);
    return rc == ANN_OK ? (jlong)(intptr_t)sx : 0;
}

/* Updatable.update (Api.scala:148-150), batched: overwrite the rows stored at `slots` (ann_update_batch) */
NATIVE(jint, updateBatch)(JNIEnv *env, jobject self, jlong handle, jobject slots, jobject rows, jlong n) {
    g_shim_error = NULL;
    (void)self;
    REQUIRE(slots, pos(n) * 8, "slots");
    REQUIRE(rows, pos(n) * dim_of(handle) * 4, "rows");
    return ann_update_batch((ann_index *)(intptr_t)handle, (const int64_t *)addr(env, slots), (const float *)addr(env, rows), n);
}

/* rows [start, start + n) and their ids back into direct buffers (ann_read_rows): what toDirectory iterates */
NATIVE(jint, readRows)(JNIEnv *env, jobject self, jlong handle, jlong start, jlong n, jobject out_ids, jobject out_rows) {
    g_shim_error = NULL;
    (void)self;
    REQUIRE(out_ids, pos(n) * 8, "outIds");
    REQUIRE(out_rows, pos(n) * dim_of(handle) * 4, "outRows");
    return ann_read_rows((ann_index *)(intptr_t)handle, start, n, (int64_t *)addr(env, out_ids), (float *)addr(env, out_rows));
}

NATIVE(jint, setOption)(JNIEnv *env, jobject self, jlong handle, jstring name, jlong value) {
    g_shim_error = NULL;
    (void)self;
    const char *s = name ? (*env)->GetStringUTFChars(env, name, NULL) : NULL;
    int rc = ann_set_option((ann_index *)(intptr_t)handle, s, value);
    if (s) (*env)->ReleaseStringUTFChars(env, name, s);
    return rc;
}

/* a stat by name, or Long.MIN_VALUE when the call failed (see lastError) */
NATIVE(jlong, getStat)(JNIEnv *env, jobject self, jlong handle, jstring name) {
    g_shim_error = NULL;
    (void)self;
    const char *s = name ? (*env)->GetStringUTFChars(env, name, NULL) : NULL;
    int64_t v = 0;
    int rc = ann_get_stat((const ann_index *)(intptr_t)handle, s, &v);
    if (s) (*env)->ReleaseStringUTFChars(env, name, s);
    return rc == ANN_OK ? (jlong)v : (jlong)INT64_MIN;
}

NATIVE(jint, version)(JNIEnv *env, jobject self) {
    g_shim_error = NULL;
    (void)env; (void)self;
    return ann_version();
}

NATIVE(jstring, lastError)(JNIEnv *env, jobject self) {
    (void)self;
    return (*env)->NewStringUTF(env, g_shim_error ? g_shim_error : ann_last_error());
}
#else
/* no JDK in this image: nothing to compile */
typedef int b200ann_jni_not_built_here;
#endif
