/* jni.h -- TEST STAND-IN, not the JDK header.
 *
 * This image has no JDK, so the JNI shim (the-algorithm_b200/host/jni/b200ann_jni.c) would otherwise never be compiled.
 * This file declares, from the public JNI specification, exactly the types, macros and function-table members the shim
 * uses, so that `-Itests/jni_stub` type-checks the shim's body and tests/jni_host_check.c can EXECUTE every
 * Java_com_twitter_ann_brute_1force_B200AnnNative_00024_* entry point against a mock JNIEnv (direct buffers, int arrays
 * and strings backed by plain C memory).  The real header lays the function table out differently (its members sit at
 * fixed slots among ~230 others); that does not matter here because shim and mock are compiled against the same
 * declaration, and against a real JDK the shim uses the same member NAMES.
 */
#ifndef B200ANN_TEST_JNI_STUB_H_
#define B200ANN_TEST_JNI_STUB_H_

#include <stdint.h>

typedef int32_t jint;
typedef int64_t jlong;
typedef int8_t jbyte;
typedef uint8_t jboolean;
typedef float jfloat;
typedef double jdouble;
typedef jint jsize;

struct _jobject; /* defined by the mock (tests/jni_host_check.c) */
typedef struct _jobject *jobject;
typedef jobject jclass;
typedef jobject jstring;
typedef jobject jarray;
typedef jarray jintArray;

struct JNINativeInterface_;
typedef const struct JNINativeInterface_ *JNIEnv;

struct JNINativeInterface_ {
    void *(*GetDirectBufferAddress)(JNIEnv *env, jobject buf);
    jlong (*GetDirectBufferCapacity)(JNIEnv *env, jobject buf);
    jsize (*GetArrayLength)(JNIEnv *env, jarray array);
    jint *(*GetIntArrayElements)(JNIEnv *env, jintArray array, jboolean *is_copy);
    void (*ReleaseIntArrayElements)(JNIEnv *env, jintArray array, jint *elems, jint mode);
    const char *(*GetStringUTFChars)(JNIEnv *env, jstring str, jboolean *is_copy);
    void (*ReleaseStringUTFChars)(JNIEnv *env, jstring str, const char *chars);
    jstring (*NewStringUTF)(JNIEnv *env, const char *utf);
};

#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_FALSE 0
#define JNI_TRUE 1
#define JNI_COMMIT 1
#define JNI_ABORT 2

#endif /* B200ANN_TEST_JNI_STUB_H_ */
