/* Reference-side JNI glue for libimm3gpu.so (include/imm3.h); see INTEGRATION.md.  Not compiled in this repository's build image (no JDK headers). */
#include <jni.h>
#include <stdint.h>
#include "imm3.h"
/* jni/imm3_jni.c — 40 lines of glue, no logic */
JNIEXPORT jlong JNICALL Java_immutabledb_engine_gpu_Imm3Jni_open(JNIEnv* e, jobject o, jstring dir, jint dev, jint rank, jint world) {
    const char* d = (*e)->GetStringUTFChars(e, dir, 0);
    imm3_open_opts opts = {dev, rank, world, 0};
    imm3_db* db = NULL;
    int rc = imm3_open(d, &opts, &db);
    (*e)->ReleaseStringUTFChars(e, dir, d);
    if (rc) { (*e)->ThrowNew(e, (*e)->FindClass(e, "java/lang/RuntimeException"), imm3_last_error()); return 0; }
    return (jlong)(intptr_t)db;
}
JNIEXPORT jobject JNICALL Java_immutabledb_engine_gpu_Imm3Jni_column(JNIEnv* e, jobject o, jlong res, jint col) {
    const imm3_result* r = (const imm3_result*)(intptr_t)res;
    return (*e)->NewDirectByteBuffer(e, (void*)imm3_result_col_data(r, col), imm3_result_nrows(r) * imm3_result_col_width(r, col));
}
