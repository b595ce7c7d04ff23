/*
 * b200_jni.c — JNI shim binding net.jpountz.lz4.LZ4B200JNI / net.jpountz.xxhash.XXHashB200JNI to
 * libb200lz4.so.  It is the twin of the reference's src/jni/net_jpountz_lz4_LZ4JNI.c and
 * src/jni/net_jpountz_xxhash_XXHashJNI.c: same argument lists (array-or-direct-buffer operands with
 * offsets), same return conventions, but each call lands in the CUDA backend.
 *
 * NOT COMPILED IN THIS REPOSITORY'S CI: the build container has no JDK (no jni.h).  Build on a
 * machine with a JDK:
 *   gcc -O2 -fPIC -shared -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -I../../include \
 *       b200_jni.c -L.. -lb200lz4 -o liblz4-java-b200.so
 *
 * Differences from the reference shim, all forced by the device boundary:
 *  - byte[] operands use Get/ReleaseByteArrayElements-free access via GetPrimitiveArrayCritical like the
 *    reference (LZ4JNI.c:54,65), but a CUDA copy + sync inside a critical section stalls the GC
 *    (SURVEY.md §8b "Ownership"); prefer direct ByteBuffers registered once with b200lz4_host_register.
 *  - LZ4_decompress_fast passes the readable source length (srcAvail) because a device copy needs a size.
 *  - the `out == NULL` path releases `in` (the reference leaks the critical section there, LZ4JNI.c:70-73).
 */
#include <jni.h>
#include <stdint.h>
#include "b200lz4.h"

static jclass OutOfMemoryError;

JNIEXPORT void JNICALL Java_net_jpountz_lz4_LZ4B200JNI_init(JNIEnv* env, jclass cls)
{
    OutOfMemoryError = (*env)->NewGlobalRef(env, (*env)->FindClass(env, "java/lang/OutOfMemoryError"));
}
static void throw_OOM(JNIEnv* env) { (*env)->ThrowNew(env, OutOfMemoryError, "Out of memory"); }

static char* acquire(JNIEnv* env, jbyteArray arr, jobject buf)
{
    return arr != NULL ? (char*)(*env)->GetPrimitiveArrayCritical(env, arr, 0)
                       : (char*)(*env)->GetDirectBufferAddress(env, buf);
}
static void release(JNIEnv* env, jbyteArray arr, char* p, jint mode)
{
    if (arr != NULL && p != NULL) (*env)->ReleasePrimitiveArrayCritical(env, arr, p, mode);
}

#define ACQUIRE2()                                                           \
    char* in = acquire(env, srcArray, srcBuffer);                            \
    if (in == NULL) { throw_OOM(env); return 0; }                            \
    char* out = acquire(env, destArray, destBuffer);                         \
    if (out == NULL) { release(env, srcArray, in, JNI_ABORT); throw_OOM(env); return 0; }
#define RELEASE2()                                                           \
    release(env, srcArray, in, JNI_ABORT);                                   \
    release(env, destArray, out, 0);

JNIEXPORT jint JNICALL Java_net_jpountz_lz4_LZ4B200JNI_LZ4_1compress_1limitedOutput
  (JNIEnv* env, jclass cls, jbyteArray srcArray, jobject srcBuffer, jint srcOff, jint srcLen,
   jbyteArray destArray, jobject destBuffer, jint destOff, jint maxDestLen)
{   /* reference: LZ4JNI.c:46-86 */
    ACQUIRE2();
    jint r = b200lz4_compress_default(in + srcOff, out + destOff, srcLen, maxDestLen);
    RELEASE2();
    return r;
}

JNIEXPORT jint JNICALL Java_net_jpountz_lz4_LZ4B200JNI_LZ4_1compressHC
  (JNIEnv* env, jclass cls, jbyteArray srcArray, jobject srcBuffer, jint srcOff, jint srcLen,
   jbyteArray destArray, jobject destBuffer, jint destOff, jint maxDestLen, jint level)
{   /* reference: LZ4JNI.c:93-133 */
    ACQUIRE2();
    jint r = b200lz4_compress_HC(in + srcOff, out + destOff, srcLen, maxDestLen, level);
    RELEASE2();
    return r;
}

JNIEXPORT jint JNICALL Java_net_jpountz_lz4_LZ4B200JNI_LZ4_1decompress_1fast
  (JNIEnv* env, jclass cls, jbyteArray srcArray, jobject srcBuffer, jint srcOff, jint srcAvail,
   jbyteArray destArray, jobject destBuffer, jint destOff, jint destLen)
{   /* reference: LZ4JNI.c:140-180 (+ srcAvail) */
    ACQUIRE2();
    jint r = b200lz4_decompress_fast_bounded(in + srcOff, srcAvail, out + destOff, destLen);
    RELEASE2();
    return r;
}

JNIEXPORT jint JNICALL Java_net_jpountz_lz4_LZ4B200JNI_LZ4_1decompress_1safe
  (JNIEnv* env, jclass cls, jbyteArray srcArray, jobject srcBuffer, jint srcOff, jint srcLen,
   jbyteArray destArray, jobject destBuffer, jint destOff, jint maxDestLen)
{   /* reference: LZ4JNI.c:187-227 */
    ACQUIRE2();
    jint r = b200lz4_decompress_safe(in + srcOff, out + destOff, srcLen, maxDestLen);
    RELEASE2();
    return r;
}

JNIEXPORT jint JNICALL Java_net_jpountz_lz4_LZ4B200JNI_LZ4_1compressBound(JNIEnv* env, jclass cls, jint len)
{   /* reference: LZ4JNI.c:234-239 */
    return b200lz4_compressBound(len);
}

/* batch entry points over direct buffers (no reference counterpart; SURVEY.md §7 hard part 1).
 * offsets/lengths/results are direct IntBuffer/LongBuffer views so nothing is copied on the Java side. */
JNIEXPORT jint JNICALL Java_net_jpountz_lz4_LZ4B200JNI_compressBatch
  (JNIEnv* env, jclass cls, jobject src, jobject srcOff, jobject srcLen, jobject dst, jobject dstOff, jobject dstCap,
   jobject result, jint n, jint maxSrcLen)
{
    return b200lz4_compress_fast_batch_host(
        (const uint8_t*)(*env)->GetDirectBufferAddress(env, src), (const uint64_t*)(*env)->GetDirectBufferAddress(env, srcOff),
        (const int32_t*)(*env)->GetDirectBufferAddress(env, srcLen), (uint8_t*)(*env)->GetDirectBufferAddress(env, dst),
        (const uint64_t*)(*env)->GetDirectBufferAddress(env, dstOff), (const int32_t*)(*env)->GetDirectBufferAddress(env, dstCap),
        (int32_t*)(*env)->GetDirectBufferAddress(env, result), (size_t)n, maxSrcLen);
}
JNIEXPORT jint JNICALL Java_net_jpountz_lz4_LZ4B200JNI_decompressFastBatch
  (JNIEnv* env, jclass cls, jobject src, jobject srcOff, jobject srcAvail, jobject dst, jobject dstOff, jobject dstLen,
   jobject result, jint n)
{
    return b200lz4_decompress_fast_batch_host(
        (const uint8_t*)(*env)->GetDirectBufferAddress(env, src), (const uint64_t*)(*env)->GetDirectBufferAddress(env, srcOff),
        (const int32_t*)(*env)->GetDirectBufferAddress(env, srcAvail), (uint8_t*)(*env)->GetDirectBufferAddress(env, dst),
        (const uint64_t*)(*env)->GetDirectBufferAddress(env, dstOff), (const int32_t*)(*env)->GetDirectBufferAddress(env, dstLen),
        (int32_t*)(*env)->GetDirectBufferAddress(env, result), (size_t)n);
}
JNIEXPORT jint JNICALL Java_net_jpountz_lz4_LZ4B200JNI_decompressSafeBatch
  (JNIEnv* env, jclass cls, jobject src, jobject srcOff, jobject srcLen, jobject dst, jobject dstOff, jobject dstCap,
   jobject result, jint n)
{
    return b200lz4_decompress_safe_batch_host(
        (const uint8_t*)(*env)->GetDirectBufferAddress(env, src), (const uint64_t*)(*env)->GetDirectBufferAddress(env, srcOff),
        (const int32_t*)(*env)->GetDirectBufferAddress(env, srcLen), (uint8_t*)(*env)->GetDirectBufferAddress(env, dst),
        (const uint64_t*)(*env)->GetDirectBufferAddress(env, dstOff), (const int32_t*)(*env)->GetDirectBufferAddress(env, dstCap),
        (int32_t*)(*env)->GetDirectBufferAddress(env, result), (size_t)n);
}
/* the same three calls range-sharded over several GPUs from this one JVM (b200lz4_*_batch_host_multi): `devices` is a
 * direct IntBuffer of device indices, or null for devices 0..ndev-1 */
static const int* device_list(JNIEnv* env, jobject devices)
{ return devices != NULL ? (const int*)(*env)->GetDirectBufferAddress(env, devices) : NULL; }
JNIEXPORT jint JNICALL Java_net_jpountz_lz4_LZ4B200JNI_compressBatchMulti
  (JNIEnv* env, jclass cls, jobject src, jobject srcOff, jobject srcLen, jobject dst, jobject dstOff, jobject dstCap,
   jobject result, jint n, jint maxSrcLen, jobject devices, jint ndev)
{
    return b200lz4_compress_fast_batch_host_multi(
        (const uint8_t*)(*env)->GetDirectBufferAddress(env, src), (const uint64_t*)(*env)->GetDirectBufferAddress(env, srcOff),
        (const int32_t*)(*env)->GetDirectBufferAddress(env, srcLen), (uint8_t*)(*env)->GetDirectBufferAddress(env, dst),
        (const uint64_t*)(*env)->GetDirectBufferAddress(env, dstOff), (const int32_t*)(*env)->GetDirectBufferAddress(env, dstCap),
        (int32_t*)(*env)->GetDirectBufferAddress(env, result), (size_t)n, maxSrcLen, device_list(env, devices), ndev);
}
JNIEXPORT jint JNICALL Java_net_jpountz_lz4_LZ4B200JNI_decompressFastBatchMulti
  (JNIEnv* env, jclass cls, jobject src, jobject srcOff, jobject srcAvail, jobject dst, jobject dstOff, jobject dstLen,
   jobject result, jint n, jobject devices, jint ndev)
{
    return b200lz4_decompress_fast_batch_host_multi(
        (const uint8_t*)(*env)->GetDirectBufferAddress(env, src), (const uint64_t*)(*env)->GetDirectBufferAddress(env, srcOff),
        (const int32_t*)(*env)->GetDirectBufferAddress(env, srcAvail), (uint8_t*)(*env)->GetDirectBufferAddress(env, dst),
        (const uint64_t*)(*env)->GetDirectBufferAddress(env, dstOff), (const int32_t*)(*env)->GetDirectBufferAddress(env, dstLen),
        (int32_t*)(*env)->GetDirectBufferAddress(env, result), (size_t)n, device_list(env, devices), ndev);
}
JNIEXPORT jint JNICALL Java_net_jpountz_lz4_LZ4B200JNI_decompressSafeBatchMulti
  (JNIEnv* env, jclass cls, jobject src, jobject srcOff, jobject srcLen, jobject dst, jobject dstOff, jobject dstCap,
   jobject result, jint n, jobject devices, jint ndev)
{
    return b200lz4_decompress_safe_batch_host_multi(
        (const uint8_t*)(*env)->GetDirectBufferAddress(env, src), (const uint64_t*)(*env)->GetDirectBufferAddress(env, srcOff),
        (const int32_t*)(*env)->GetDirectBufferAddress(env, srcLen), (uint8_t*)(*env)->GetDirectBufferAddress(env, dst),
        (const uint64_t*)(*env)->GetDirectBufferAddress(env, dstOff), (const int32_t*)(*env)->GetDirectBufferAddress(env, dstCap),
        (int32_t*)(*env)->GetDirectBufferAddress(env, result), (size_t)n, device_list(env, devices), ndev);
}
/* packed output per GPU shard (b200lz4_compress_fast_compact_host_multi): outOff / shardBase / shardTotal are direct LongBuffers */
JNIEXPORT jint JNICALL Java_net_jpountz_lz4_LZ4B200JNI_compressPackedMulti
  (JNIEnv* env, jclass cls, jobject src, jobject srcOff, jobject srcLen, jobject dst, jobject outOff, jobject result, jint n,
   jint maxSrcLen, jobject devices, jint ndev, jobject shardBase, jobject shardTotal)
{
    return b200lz4_compress_fast_compact_host_multi(
        (const uint8_t*)(*env)->GetDirectBufferAddress(env, src), (const uint64_t*)(*env)->GetDirectBufferAddress(env, srcOff),
        (const int32_t*)(*env)->GetDirectBufferAddress(env, srcLen), (uint8_t*)(*env)->GetDirectBufferAddress(env, dst),
        (size_t)(*env)->GetDirectBufferCapacity(env, dst), (uint64_t*)(*env)->GetDirectBufferAddress(env, outOff),
        (int32_t*)(*env)->GetDirectBufferAddress(env, result), (size_t)n, maxSrcLen, device_list(env, devices), ndev,
        (uint64_t*)(*env)->GetDirectBufferAddress(env, shardBase), (uint64_t*)(*env)->GetDirectBufferAddress(env, shardTotal));
}
JNIEXPORT jint JNICALL Java_net_jpountz_lz4_LZ4B200JNI_deviceCount(JNIEnv* env, jclass cls)
{ (void)env; (void)cls; return b200lz4_device_count(); }
JNIEXPORT jint JNICALL Java_net_jpountz_lz4_LZ4B200JNI_registerDirectBuffer(JNIEnv* env, jclass cls, jobject buf)
{
    return b200lz4_host_register((*env)->GetDirectBufferAddress(env, buf), (size_t)(*env)->GetDirectBufferCapacity(env, buf));
}

/* ---------------------------------------------------------------- XXHash (reference: XXHashJNI.c:42-255) */
JNIEXPORT void JNICALL Java_net_jpountz_xxhash_XXHashB200JNI_init(JNIEnv* env, jclass cls) { (void)env; (void)cls; }

/* The hash calls return the hash value, so a backend failure (no device, CUDA error) cannot travel in the return
 * value: it is thrown as java.lang.IllegalStateException carrying b200lz4_last_error() — never a silent 0. */
static void throw_if_failed(JNIEnv* env)
{
    if (b200lz4_last_status() != 0) {
        jclass ise = (*env)->FindClass(env, "java/lang/IllegalStateException");
        if (ise != NULL) (*env)->ThrowNew(env, ise, b200lz4_last_error());
    }
}

JNIEXPORT jint JNICALL Java_net_jpountz_xxhash_XXHashB200JNI_XXH32
  (JNIEnv* env, jclass cls, jbyteArray buf, jint off, jint len, jint seed)
{
    char* in = (char*)(*env)->GetPrimitiveArrayCritical(env, buf, 0);
    if (in == NULL) return 0;
    jint h = (jint)b200xxh32(in + off, (size_t)len, (uint32_t)seed);
    (*env)->ReleasePrimitiveArrayCritical(env, buf, in, JNI_ABORT);
    throw_if_failed(env);
    return h;
}
JNIEXPORT jint JNICALL Java_net_jpountz_xxhash_XXHashB200JNI_XXH32BB
  (JNIEnv* env, jclass cls, jobject buf, jint off, jint len, jint seed)
{
    char* in = (char*)(*env)->GetDirectBufferAddress(env, buf);
    if (in == NULL) return 0;
    jint h = (jint)b200xxh32(in + off, (size_t)len, (uint32_t)seed);
    throw_if_failed(env);
    return h;
}
JNIEXPORT jlong JNICALL Java_net_jpountz_xxhash_XXHashB200JNI_XXH32_1init(JNIEnv* env, jclass cls, jint seed)
{ return (jlong)(intptr_t)b200xxh32_create((uint32_t)seed); }
JNIEXPORT void JNICALL Java_net_jpountz_xxhash_XXHashB200JNI_XXH32_1reset(JNIEnv* env, jclass cls, jlong state, jint seed)
{ b200xxh32_reset((void*)(intptr_t)state, (uint32_t)seed); }
JNIEXPORT void JNICALL Java_net_jpountz_xxhash_XXHashB200JNI_XXH32_1update
  (JNIEnv* env, jclass cls, jlong state, jbyteArray src, jint off, jint len)
{
    char* in = (char*)(*env)->GetPrimitiveArrayCritical(env, src, 0);
    if (in == NULL) return;
    int rc = b200xxh32_update((void*)(intptr_t)state, in + off, (size_t)len);
    (*env)->ReleasePrimitiveArrayCritical(env, src, in, JNI_ABORT);
    if (rc != 0) throw_if_failed(env);
}
JNIEXPORT jint JNICALL Java_net_jpountz_xxhash_XXHashB200JNI_XXH32_1digest(JNIEnv* env, jclass cls, jlong state)
{ jint h = (jint)b200xxh32_digest((void*)(intptr_t)state); throw_if_failed(env); return h; }
JNIEXPORT void JNICALL Java_net_jpountz_xxhash_XXHashB200JNI_XXH32_1free(JNIEnv* env, jclass cls, jlong state)
{ b200xxh32_free((void*)(intptr_t)state); }

JNIEXPORT jlong JNICALL Java_net_jpountz_xxhash_XXHashB200JNI_XXH64
  (JNIEnv* env, jclass cls, jbyteArray buf, jint off, jint len, jlong seed)
{
    char* in = (char*)(*env)->GetPrimitiveArrayCritical(env, buf, 0);
    if (in == NULL) return 0;
    jlong h = (jlong)b200xxh64(in + off, (size_t)len, (uint64_t)seed);
    (*env)->ReleasePrimitiveArrayCritical(env, buf, in, JNI_ABORT);
    throw_if_failed(env);
    return h;
}
JNIEXPORT jlong JNICALL Java_net_jpountz_xxhash_XXHashB200JNI_XXH64BB
  (JNIEnv* env, jclass cls, jobject buf, jint off, jint len, jlong seed)
{
    char* in = (char*)(*env)->GetDirectBufferAddress(env, buf);
    if (in == NULL) return 0;
    jlong h = (jlong)b200xxh64(in + off, (size_t)len, (uint64_t)seed);
    throw_if_failed(env);
    return h;
}
JNIEXPORT jlong JNICALL Java_net_jpountz_xxhash_XXHashB200JNI_XXH64_1init(JNIEnv* env, jclass cls, jlong seed)
{ return (jlong)(intptr_t)b200xxh64_create((uint64_t)seed); }
JNIEXPORT void JNICALL Java_net_jpountz_xxhash_XXHashB200JNI_XXH64_1reset(JNIEnv* env, jclass cls, jlong state, jlong seed)
{ b200xxh64_reset((void*)(intptr_t)state, (uint64_t)seed); }
JNIEXPORT void JNICALL Java_net_jpountz_xxhash_XXHashB200JNI_XXH64_1update
  (JNIEnv* env, jclass cls, jlong state, jbyteArray src, jint off, jint len)
{
    char* in = (char*)(*env)->GetPrimitiveArrayCritical(env, src, 0);
    if (in == NULL) return;
    int rc = b200xxh64_update((void*)(intptr_t)state, in + off, (size_t)len);
    (*env)->ReleasePrimitiveArrayCritical(env, src, in, JNI_ABORT);
    if (rc != 0) throw_if_failed(env);
}
JNIEXPORT jlong JNICALL Java_net_jpountz_xxhash_XXHashB200JNI_XXH64_1digest(JNIEnv* env, jclass cls, jlong state)
{ jlong h = (jlong)b200xxh64_digest((void*)(intptr_t)state); throw_if_failed(env); return h; }
JNIEXPORT void JNICALL Java_net_jpountz_xxhash_XXHashB200JNI_XXH64_1free(JNIEnv* env, jclass cls, jlong state)
{ b200xxh64_free((void*)(intptr_t)state); }

/* batch hashing over a direct buffer (config 5: 100 M x 4 KiB) */
JNIEXPORT jint JNICALL Java_net_jpountz_xxhash_XXHashB200JNI_XXH64Batch
  (JNIEnv* env, jclass cls, jobject buf, jobject off, jobject len, jlong seed, jobject out, jint n)
{
    return b200xxh64_batch_host((const uint8_t*)(*env)->GetDirectBufferAddress(env, buf),
        (const uint64_t*)(*env)->GetDirectBufferAddress(env, off), (const int32_t*)(*env)->GetDirectBufferAddress(env, len),
        (uint64_t)seed, (uint64_t*)(*env)->GetDirectBufferAddress(env, out), (size_t)n);
}
JNIEXPORT jint JNICALL Java_net_jpountz_xxhash_XXHashB200JNI_XXH32Batch
  (JNIEnv* env, jclass cls, jobject buf, jobject off, jobject len, jint seed, jobject out, jint n)
{
    return b200xxh32_batch_host((const uint8_t*)(*env)->GetDirectBufferAddress(env, buf),
        (const uint64_t*)(*env)->GetDirectBufferAddress(env, off), (const int32_t*)(*env)->GetDirectBufferAddress(env, len),
        (uint32_t)seed, (uint32_t*)(*env)->GetDirectBufferAddress(env, out), (size_t)n);
}
