/* kassign_jni.c — JNI shim over include/kassign.h for NativeKafkaTopicAssigner.java.
 * NOT compiled in this repository: the build image has no jni.h. Build on a box with a JDK:
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude kafka_assigner_b200/jni/kassign_jni.c \
 *       -Lkafka_assigner_b200/csrc -lkassign -o libkassign_jni.so */
#include <jni.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "kassign.h"

static void throw_status(JNIEnv* env, const ka_status* st, jobjectArray names) {
    char msg[512]; const char* cls = "java/lang/IllegalStateException"; const char* t = "?";
    jstring js = NULL;
    if (st->topic_index >= 0) { js = (*env)->GetObjectArrayElement(env, names, st->topic_index); t = (*env)->GetStringUTFChars(env, js, 0); }
    switch (st->code) {
    case KA_ERR_RF_MISMATCH:     snprintf(msg, sizeof msg, "Topic %s has partition %d with unexpected replication factor %d", t, st->partition, st->a); break;   /* KTA:58-60 */
    case KA_ERR_RF_NOT_POSITIVE: snprintf(msg, sizeof msg, "Topic %s does not have a positive replication factor!", t); break;                                   /* KTA:65-66 */
    case KA_ERR_RF_GT_BROKERS:   snprintf(msg, sizeof msg, "Topic %s has a higher replication factor (%d) than available brokers!", t, st->a); break;             /* KTA:67-69 */
    case KA_ERR_UNASSIGNABLE:    snprintf(msg, sizeof msg, "Partition %d could not be fully assigned!", st->partition); break;                                    /* KAS:183-184 */
    case KA_ERR_HASH_INDEX:      snprintf(msg, sizeof msg, "%d", st->a); cls = "java/lang/ArrayIndexOutOfBoundsException"; break;                                 /* KAS:190-192 */
    default:                     snprintf(msg, sizeof msg, "kassign error %d", st->code); cls = "java/lang/RuntimeException";
    }
    if (js) (*env)->ReleaseStringUTFChars(env, js, t);
    (*env)->ThrowNew(env, (*env)->FindClass(env, cls), msg);
}

JNIEXPORT jlong JNICALL Java_siftscience_kafka_tools_NativeAssigner_create(JNIEnv* e, jclass c, jint dev) { return (jlong)(intptr_t)ka_ctx_create(dev); }
JNIEXPORT void  JNICALL Java_siftscience_kafka_tools_NativeAssigner_destroy(JNIEnv* e, jclass c, jlong h) { ka_ctx_destroy((ka_ctx*)(intptr_t)h); }

JNIEXPORT void JNICALL Java_siftscience_kafka_tools_NativeAssigner_setBrokers(JNIEnv* e, jclass c, jlong h, jintArray ids, jobjectArray racks) {
    jsize n = (*e)->GetArrayLength(e, ids);
    jint* id = (*e)->GetIntArrayElements(e, ids, 0);
    const char** names = calloc(n ? n : 1, sizeof *names); jstring* js = calloc(n ? n : 1, sizeof *js); int32_t* idx = malloc((n ? n : 1) * sizeof *idx);
    int32_t rc = (names && js && idx) ? KA_OK : KA_ERR_BAD_ARG;
    if (rc == KA_OK) {
        for (jsize i = 0; i < n; i++) { js[i] = (*e)->GetObjectArrayElement(e, racks, i); names[i] = js[i] ? (*e)->GetStringUTFChars(e, js[i], 0) : NULL; }
        rc = ka_rack_indices(n, (const int32_t*)id, names, idx);            /* KAS:81-94 string-keyed racks */
        if (rc == KA_OK) rc = ka_ctx_set_brokers((ka_ctx*)(intptr_t)h, n, (const int32_t*)id, idx);
        for (jsize i = 0; i < n; i++) if (js[i]) (*e)->ReleaseStringUTFChars(e, js[i], names[i]);
    }
    free(names); free(js); free(idx); (*e)->ReleaseIntArrayElements(e, ids, id, JNI_ABORT);
    if (rc != KA_OK) {   /* the Java side updates its cached broker table only when this call returns normally */
        char msg[96];
        snprintf(msg, sizeof msg, "kassign: broker table upload failed (code %d)", (int)rc);
        (*e)->ThrowNew(e, (*e)->FindClass(e, "java/lang/RuntimeException"), msg);
    }
}

JNIEXPORT void JNICALL Java_siftscience_kafka_tools_NativeAssigner_solve(JNIEnv* e, jclass c, jlong h, jobjectArray names, jintArray hash,
        jlongArray partOff, jintArray partId, jlongArray repOff, jintArray cur, jint desiredRf, jint stride, jintArray outLen, jintArray out) {
    jint *ph = (*e)->GetIntArrayElements(e, hash, 0), *pp = (*e)->GetIntArrayElements(e, partId, 0), *pc = (*e)->GetIntArrayElements(e, cur, 0);
    jlong *po = (*e)->GetLongArrayElements(e, partOff, 0), *pr = (*e)->GetLongArrayElements(e, repOff, 0);
    jint *ol = (*e)->GetIntArrayElements(e, outLen, 0), *ob = (*e)->GetIntArrayElements(e, out, 0);
    ka_status st;
    ka_solve((ka_ctx*)(intptr_t)h, (*e)->GetArrayLength(e, hash), (const int32_t*)ph, (const int64_t*)po, (const int32_t*)pp,
             (const int64_t*)pr, (const int32_t*)pc, desiredRf, stride, (int32_t*)ol, (int32_t*)ob, &st);
    (*e)->ReleaseIntArrayElements(e, hash, ph, JNI_ABORT); (*e)->ReleaseIntArrayElements(e, partId, pp, JNI_ABORT);
    (*e)->ReleaseIntArrayElements(e, cur, pc, JNI_ABORT);  (*e)->ReleaseLongArrayElements(e, partOff, po, JNI_ABORT);
    (*e)->ReleaseLongArrayElements(e, repOff, pr, JNI_ABORT);
    (*e)->ReleaseIntArrayElements(e, outLen, ol, 0); (*e)->ReleaseIntArrayElements(e, out, ob, 0);
    if (st.code != KA_OK) throw_status(e, &st, names);
}
