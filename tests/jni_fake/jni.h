/* tests/jni_fake/jni.h — a minimal stand-in for the JDK's jni.h: just the slice of JNIEnv that
 * lz4-java_b200/jni/b200_jni.c uses, so the shim can be compiled and driven from C in an
 * environment without a JDK.  Java objects are plain structs here. */
#ifndef FAKE_JNI_H
#define FAKE_JNI_H
#include <stdint.h>
#include <stddef.h>
#define JNIEXPORT
#define JNICALL
#define JNI_ABORT 2
typedef int32_t jint; typedef int64_t jlong; typedef int8_t jbyte; typedef uint8_t jboolean;
typedef struct fake_obj { void* data; jlong len; } *jobject;
typedef jobject jclass; typedef jobject jbyteArray; typedef jobject jarray;
struct JNINativeInterface_;
typedef const struct JNINativeInterface_* JNIEnv;
struct JNINativeInterface_ {
    jclass (*FindClass)(JNIEnv*, const char*);
    jobject (*NewGlobalRef)(JNIEnv*, jobject);
    jint (*ThrowNew)(JNIEnv*, jclass, const char*);
    void* (*GetPrimitiveArrayCritical)(JNIEnv*, jarray, jboolean*);
    void (*ReleasePrimitiveArrayCritical)(JNIEnv*, jarray, void*, jint);
    void* (*GetDirectBufferAddress)(JNIEnv*, jobject);
    jlong (*GetDirectBufferCapacity)(JNIEnv*, jobject);
};
#endif
