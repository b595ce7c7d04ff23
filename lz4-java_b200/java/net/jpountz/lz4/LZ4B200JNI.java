package net.jpountz.lz4;

import java.nio.ByteBuffer;
import java.nio.IntBuffer;
import java.nio.LongBuffer;

/**
 * JNI bindings to libb200lz4 (hand-written sm_100a CUDA), the "B200" backend.
 * Twin of {@code LZ4JNI} (src/java/net/jpountz/lz4/LZ4JNI.java:28-42); the native side is
 * lz4-java_b200/jni/b200_jni.c.  NOT compiled in the b200 repository (no JDK there).
 */
enum LZ4B200JNI {
  ;

  static {
    System.loadLibrary("b200lz4");          // the CUDA library (include/b200lz4.h)
    System.loadLibrary("lz4-java-b200");    // the JNI shim
    init();
  }

  static native void init();
  static native int LZ4_compress_limitedOutput(byte[] srcArray, ByteBuffer srcBuffer, int srcOff, int srcLen, byte[] destArray, ByteBuffer destBuffer, int destOff, int maxDestLen);
  static native int LZ4_compressHC(byte[] srcArray, ByteBuffer srcBuffer, int srcOff, int srcLen, byte[] destArray, ByteBuffer destBuffer, int destOff, int maxDestLen, int compressionLevel);
  /** Unlike LZ4JNI, the readable source length is passed: a device copy needs a size. */
  static native int LZ4_decompress_fast(byte[] srcArray, ByteBuffer srcBuffer, int srcOff, int srcAvail, byte[] destArray, ByteBuffer destBuffer, int destOff, int destLen);
  static native int LZ4_decompress_safe(byte[] srcArray, ByteBuffer srcBuffer, int srcOff, int srcLen, byte[] destArray, ByteBuffer destBuffer, int destOff, int maxDestLen);
  static native int LZ4_compressBound(int len);

  /* batch calls over DIRECT buffers: one launch for n independent blocks */
  static native int compressBatch(ByteBuffer src, LongBuffer srcOff, IntBuffer srcLen, ByteBuffer dst, LongBuffer dstOff, IntBuffer dstCap, IntBuffer result, int n, int maxSrcLen);
  static native int decompressFastBatch(ByteBuffer src, LongBuffer srcOff, IntBuffer srcAvail, ByteBuffer dst, LongBuffer dstOff, IntBuffer dstLen, IntBuffer result, int n);
  static native int decompressSafeBatch(ByteBuffer src, LongBuffer srcOff, IntBuffer srcLen, ByteBuffer dst, LongBuffer dstOff, IntBuffer dstCap, IntBuffer result, int n);
  /* the same, range-sharded over several GPUs from this one JVM; devices: direct IntBuffer of device indices, or null for 0..ndev-1 */
  static native int compressBatchMulti(ByteBuffer src, LongBuffer srcOff, IntBuffer srcLen, ByteBuffer dst, LongBuffer dstOff, IntBuffer dstCap, IntBuffer result, int n, int maxSrcLen, IntBuffer devices, int ndev);
  static native int decompressFastBatchMulti(ByteBuffer src, LongBuffer srcOff, IntBuffer srcAvail, ByteBuffer dst, LongBuffer dstOff, IntBuffer dstLen, IntBuffer result, int n, IntBuffer devices, int ndev);
  static native int decompressSafeBatchMulti(ByteBuffer src, LongBuffer srcOff, IntBuffer srcLen, ByteBuffer dst, LongBuffer dstOff, IntBuffer dstCap, IntBuffer result, int n, IntBuffer devices, int ndev);
  /** packed output per GPU shard: shard g's blocks back to back from dst[shardBase[g]], shardTotal[g] bytes; outOff absolute */
  static native int compressPackedMulti(ByteBuffer src, LongBuffer srcOff, IntBuffer srcLen, ByteBuffer dst, LongBuffer outOff, IntBuffer result, int n, int maxSrcLen, IntBuffer devices, int ndev, LongBuffer shardBase, LongBuffer shardTotal);
  static native int deviceCount();
  static native int registerDirectBuffer(ByteBuffer buf);
}
