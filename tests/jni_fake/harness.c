/* Drives the JNI shim through a fake JNIEnv: byte[] and direct-ByteBuffer operands, offsets,
 * return conventions, critical-section pairing.  Prints "ok" or a diagnostic. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "jni.h"

static int pinned = 0, thrown = 0;
static jclass f_FindClass(JNIEnv* e, const char* n) { static struct fake_obj c; return &c; }
static jobject f_NewGlobalRef(JNIEnv* e, jobject o) { return o; }
static jint f_ThrowNew(JNIEnv* e, jclass c, const char* m) { thrown++; return 0; }
static void* f_GetCrit(JNIEnv* e, jarray a, jboolean* c) { pinned++; return a->data; }
static void f_RelCrit(JNIEnv* e, jarray a, void* p, jint m) { pinned--; }
static void* f_DBA(JNIEnv* e, jobject o) { return o->data; }
static jlong f_DBC(JNIEnv* e, jobject o) { return o->len; }
static const struct JNINativeInterface_ table = { f_FindClass, f_NewGlobalRef, f_ThrowNew, f_GetCrit, f_RelCrit, f_DBA, f_DBC };

void Java_net_jpountz_lz4_LZ4B200JNI_init(JNIEnv*, jclass);
jint Java_net_jpountz_lz4_LZ4B200JNI_LZ4_1compress_1limitedOutput(JNIEnv*, jclass, jbyteArray, jobject, jint, jint, jbyteArray, jobject, jint, jint);
jint Java_net_jpountz_lz4_LZ4B200JNI_LZ4_1decompress_1fast(JNIEnv*, jclass, jbyteArray, jobject, jint, jint, jbyteArray, jobject, jint, jint);
jint Java_net_jpountz_lz4_LZ4B200JNI_LZ4_1decompress_1safe(JNIEnv*, jclass, jbyteArray, jobject, jint, jint, jbyteArray, jobject, jint, jint);
jint Java_net_jpountz_lz4_LZ4B200JNI_LZ4_1compressBound(JNIEnv*, jclass, jint);
jint Java_net_jpountz_xxhash_XXHashB200JNI_XXH32(JNIEnv*, jclass, jbyteArray, jint, jint, jint);
jlong Java_net_jpountz_xxhash_XXHashB200JNI_XXH64BB(JNIEnv*, jclass, jobject, jint, jint, jlong);
jlong Java_net_jpountz_xxhash_XXHashB200JNI_XXH64_1init(JNIEnv*, jclass, jlong);
void Java_net_jpountz_xxhash_XXHashB200JNI_XXH64_1reset(JNIEnv*, jclass, jlong, jlong);
void Java_net_jpountz_xxhash_XXHashB200JNI_XXH64_1update(JNIEnv*, jclass, jlong, jbyteArray, jint, jint);
jlong Java_net_jpountz_xxhash_XXHashB200JNI_XXH64_1digest(JNIEnv*, jclass, jlong);
void Java_net_jpountz_xxhash_XXHashB200JNI_XXH64_1free(JNIEnv*, jclass, jlong);

int main(void)
{
    JNIEnv envp = &table; JNIEnv* env = &envp;
    enum { N = 20000, OFF = 7 };
    unsigned char* raw = malloc(N + OFF), *comp = malloc(N * 2 + OFF), *back = malloc(N + OFF);
    struct fake_obj a_raw = { raw, N + OFF }, a_comp = { comp, N * 2 + OFF }, a_back = { back, N + OFF };
    unsigned s = 1; int i;
    for (i = 0; i < N + OFF; i++) { s = s * 1103515245u + 12345u; raw[i] = (i % 97 < 60) ? (unsigned char)(i % 13) : (unsigned char)(s >> 24); }
    Java_net_jpountz_lz4_LZ4B200JNI_init(env, NULL);
    jint bound = Java_net_jpountz_lz4_LZ4B200JNI_LZ4_1compressBound(env, NULL, N);
    if (bound != N + N / 255 + 16) { printf("bound %d\n", bound); return 1; }
    /* byte[] -> direct buffer, with offsets (exactly one of array/buffer is non-null per operand) */
    jint c = Java_net_jpountz_lz4_LZ4B200JNI_LZ4_1compress_1limitedOutput(env, NULL, &a_raw, NULL, OFF, N, NULL, &a_comp, OFF, bound);
    if (c <= 0 || c >= N) { printf("compress returned %d\n", c); return 1; }
    jint d = Java_net_jpountz_lz4_LZ4B200JNI_LZ4_1decompress_1safe(env, NULL, NULL, &a_comp, OFF, c, &a_back, NULL, OFF, N);
    if (d != N || memcmp(raw + OFF, back + OFF, N)) { printf("safe returned %d\n", d); return 1; }
    memset(back, 0, N + OFF);
    jint r = Java_net_jpountz_lz4_LZ4B200JNI_LZ4_1decompress_1fast(env, NULL, NULL, &a_comp, OFF, N * 2, &a_back, NULL, OFF, N);
    if (r != c || memcmp(raw + OFF, back + OFF, N)) { printf("fast returned %d (want %d)\n", r, c); return 1; }
    /* too-small destination -> 0, which the Java wrapper maps to LZ4Exception */
    if (Java_net_jpountz_lz4_LZ4B200JNI_LZ4_1compress_1limitedOutput(env, NULL, &a_raw, NULL, OFF, N, NULL, &a_comp, OFF, 10) != 0) { printf("limit\n"); return 1; }
    /* hashes: one-shot over byte[] and direct buffer agree with the streaming state */
    jlong h64 = Java_net_jpountz_xxhash_XXHashB200JNI_XXH64BB(env, NULL, &a_raw, OFF, N, 42);
    jlong st = Java_net_jpountz_xxhash_XXHashB200JNI_XXH64_1init(env, NULL, 42);
    Java_net_jpountz_xxhash_XXHashB200JNI_XXH64_1update(env, NULL, st, &a_raw, OFF, 1234);
    Java_net_jpountz_xxhash_XXHashB200JNI_XXH64_1update(env, NULL, st, &a_raw, OFF + 1234, N - 1234);
    jlong hs = Java_net_jpountz_xxhash_XXHashB200JNI_XXH64_1digest(env, NULL, st);
    /* getValue() is idempotent, and reset() re-seeds the SAME device state in place (B200StreamState.reset) */
    if (Java_net_jpountz_xxhash_XXHashB200JNI_XXH64_1digest(env, NULL, st) != hs) { printf("digest not idempotent\n"); return 1; }
    Java_net_jpountz_xxhash_XXHashB200JNI_XXH64_1reset(env, NULL, st, 42);
    Java_net_jpountz_xxhash_XXHashB200JNI_XXH64_1update(env, NULL, st, &a_raw, OFF, N);
    if (Java_net_jpountz_xxhash_XXHashB200JNI_XXH64_1digest(env, NULL, st) != hs) { printf("reset + update differs\n"); return 1; }
    Java_net_jpountz_xxhash_XXHashB200JNI_XXH64_1free(env, NULL, st);
    if (h64 != hs) { printf("xxh64 %llx != %llx\n", (long long)h64, (long long)hs); return 1; }
    jint h32 = Java_net_jpountz_xxhash_XXHashB200JNI_XXH32(env, NULL, &a_raw, OFF, N, 7);
    if (pinned != 0 || thrown != 0) { printf("critical sections unbalanced: %d, thrown %d\n", pinned, thrown); return 1; }
    printf("ok c=%d xxh64=%016llx xxh32=%08x\n", c, (unsigned long long)h64, (unsigned)h32);
    return 0;
}
