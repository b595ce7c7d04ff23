package net.jpountz.xxhash;

import java.nio.ByteBuffer;
import java.nio.IntBuffer;
import java.nio.LongBuffer;

/** JNI bindings of the B200 XXHash kernels (twin of XXHashJNI.java:24-45). */
enum XXHashB200JNI {
  ;

  static {
    System.loadLibrary("b200lz4");
    System.loadLibrary("lz4-java-b200");
    init();
  }

  private static native void init();
  static native int XXH32(byte[] input, int offset, int len, int seed);
  static native int XXH32BB(ByteBuffer input, int offset, int len, int seed);
  static native long XXH32_init(int seed);
  static native void XXH32_reset(long state, int seed);      // no reference counterpart: re-seeds the device state in place
  static native void XXH32_update(long state, byte[] input, int offset, int len);
  static native int XXH32_digest(long state);
  static native void XXH32_free(long state);
  static native long XXH64(byte[] input, int offset, int len, long seed);
  static native long XXH64BB(ByteBuffer input, int offset, int len, long seed);
  static native long XXH64_init(long seed);
  static native void XXH64_reset(long state, long seed);
  static native void XXH64_update(long state, byte[] input, int offset, int len);
  static native long XXH64_digest(long state);
  static native void XXH64_free(long state);
  static native int XXH32Batch(ByteBuffer buf, LongBuffer off, IntBuffer len, int seed, IntBuffer out, int n);
  static native int XXH64Batch(ByteBuffer buf, LongBuffer off, IntBuffer len, long seed, LongBuffer out, int n);
}
