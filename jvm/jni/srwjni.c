/* srwjni.c -- JNI shim between au.csiro.data61.randomwalk.nativewalk.SrwNative and libsrw (include/srw.h).
 * NOT compiled in this repository: the build image has no JDK (no jni.h).  Build line in jvm/README.md. */
#include <jni.h>
#include <stdint.h>
#include <stdlib.h>

#include "srw.h"

static void throw_srw(JNIEnv *env) {
  jclass cls = (*env)->FindClass(env, "java/lang/RuntimeException");
  if (cls) (*env)->ThrowNew(env, cls, srw_last_error());
}

JNIEXPORT jint JNICALL Java_au_csiro_data61_randomwalk_nativewalk_SrwNative_00024_deviceCount(JNIEnv *env, jobject self) {
  (void)env; (void)self;
  return (jint)srw_device_count();
}

/* Main.main / runJob for --cmd randomwalk (Main.scala:18-27, 109-127) */
JNIEXPORT jint JNICALL Java_au_csiro_data61_randomwalk_nativewalk_SrwNative_00024_runRandomWalk(JNIEnv *env, jobject self, jobjectArray jargv) {
  (void)self;
  const jsize n = (*env)->GetArrayLength(env, jargv);
  const char **argv = (const char **)calloc((size_t)(n > 0 ? n : 1), sizeof(char *));
  jstring *held = (jstring *)calloc((size_t)(n > 0 ? n : 1), sizeof(jstring));
  if (!argv || !held) { free(argv); free(held); return -1; }
  for (jsize i = 0; i < n; ++i) {
    held[i] = (jstring)(*env)->GetObjectArrayElement(env, jargv, i);
    argv[i] = (*env)->GetStringUTFChars(env, held[i], NULL);
  }
  const int rc = srw_main((int)n, argv);
  for (jsize i = 0; i < n; ++i) (*env)->ReleaseStringUTFChars(env, held[i], argv[i]);
  free(argv); free(held);
  if (rc != 0) throw_srw(env);
  return (jint)rc;
}

/* loadGraph() (URW:17-88 / VRW:13-98 from arrays) + randomWalk() (RW:75-176) */
JNIEXPORT jobjectArray JNICALL Java_au_csiro_data61_randomwalk_nativewalk_SrwNative_00024_walk(
    JNIEnv *env, jobject self, jintArray jsrc, jintArray jdst, jfloatArray jw, jintArray jpid, jboolean directed, jint walkLength,
    jint numWalks, jdouble p, jdouble q, jlong seed, jint sampler) {
  (void)self;
  const jsize n = (*env)->GetArrayLength(env, jsrc);
  jint *src = (*env)->GetIntArrayElements(env, jsrc, NULL), *dst = (*env)->GetIntArrayElements(env, jdst, NULL);
  jfloat *w = jw ? (*env)->GetFloatArrayElements(env, jw, NULL) : NULL;
  jint *pid = jpid ? (*env)->GetIntArrayElements(env, jpid, NULL) : NULL;
  srw_params prm;
  srw_params_default(&prm);
  prm.walk_length = walkLength; prm.num_walks = numWalks; prm.p = p; prm.q = q; prm.seed = (uint64_t)seed;
  prm.sampler = sampler; prm.directed = directed ? 1 : 0;
  srw_graph *g = NULL;
  srw_paths *paths = NULL;
  srw_status st = srw_graph_from_edges((int64_t)n, (const int32_t *)src, (const int32_t *)dst, w, (const int32_t *)pid, directed ? 1 : 0,
                                       sampler == SRW_SAMPLER_EXACT ? SRW_BUILD_ALL : SRW_BUILD_ALIAS, &g);
  (*env)->ReleaseIntArrayElements(env, jsrc, src, JNI_ABORT);
  (*env)->ReleaseIntArrayElements(env, jdst, dst, JNI_ABORT);
  if (w) (*env)->ReleaseFloatArrayElements(env, jw, w, JNI_ABORT);
  if (pid) (*env)->ReleaseIntArrayElements(env, jpid, pid, JNI_ABORT);
  if (st == SRW_OK) st = srw_walk(g, &prm, &paths);
  if (st != SRW_OK) { srw_graph_free(g); throw_srw(env); return NULL; }
  int64_t n_paths = 0;
  const int32_t *ids = NULL;
  const int64_t *offs = NULL;
  srw_paths_view(paths, &n_paths, &ids, &offs);
  jobjectArray out = (*env)->NewObjectArray(env, (jsize)n_paths, (*env)->FindClass(env, "[I"), NULL);
  for (int64_t i = 0; out && i < n_paths; ++i) {
    const jsize len = (jsize)(offs[i + 1] - offs[i]);
    jintArray row = (*env)->NewIntArray(env, len);
    if (!row) break;
    (*env)->SetIntArrayRegion(env, row, 0, len, (const jint *)(ids + offs[i]));
    (*env)->SetObjectArrayElement(env, out, (jsize)i, row);
    (*env)->DeleteLocalRef(env, row);
  }
  srw_paths_free(paths);
  srw_graph_free(g);
  return out;
}
