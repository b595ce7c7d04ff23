package net.jpountz.lz4;

import java.nio.ByteBuffer;
import java.nio.ByteOrder;
import java.nio.IntBuffer;
import java.nio.LongBuffer;

/**
 * Batch API of the B200 backend: n independent blocks per call over DIRECT ByteBuffers
 * (register them once with {@link #pin(ByteBuffer)} so the device DMAs straight from/to them).
 * This is what makes a GPU backend worthwhile: the per-block {@code LZ4Compressor.compress} API
 * pays one launch and two PCIe round trips for 64 KiB (SURVEY.md §7).
 */
public final class LZ4B200Batch {
  private LZ4B200Batch() {}

  public static void pin(ByteBuffer direct) {
    if (!direct.isDirect()) throw new IllegalArgumentException("direct buffer required");
    if (LZ4B200JNI.registerDirectBuffer(direct) != 0) throw new LZ4Exception("cudaHostRegister failed");
  }

  static LongBuffer longs(int n) { return ByteBuffer.allocateDirect(8 * n).order(ByteOrder.nativeOrder()).asLongBuffer(); }
  static IntBuffer ints(int n) { return ByteBuffer.allocateDirect(4 * n).order(ByteOrder.nativeOrder()).asIntBuffer(); }

  /** Number of usable B200 devices (b200lz4_device_count). */
  public static int deviceCount() {
    final int n = LZ4B200JNI.deviceCount();
    if (n < 0) throw new LZ4Exception("B200 backend error " + n);
    return n;
  }

  /** {@link #compressUniform(ByteBuffer, int, int, ByteBuffer)} range-sharded over the first {@code gpus} devices: GPU g
   *  compresses blocks [g*n/gpus, (g+1)*n/gpus) on its own streams; same bytes as the single-GPU call. */
  public static int[] compressUniform(ByteBuffer src, int blockSize, int n, ByteBuffer dst, int gpus) {
    final int bound = LZ4Utils.maxCompressedLength(blockSize);
    final LongBuffer so = longs(n), dof = longs(n);
    final IntBuffer sl = ints(n), dc = ints(n), res = ints(n);
    for (int i = 0; i < n; i++) { so.put(i, (long) i * blockSize); sl.put(i, blockSize); dof.put(i, (long) i * bound); dc.put(i, bound); }
    final int rc = LZ4B200JNI.compressBatchMulti(src, so, sl, dst, dof, dc, res, n, blockSize, null, gpus);
    if (rc != 0) throw new LZ4Exception("B200 backend error " + rc);
    final int[] out = new int[n];
    res.get(out);
    for (int r : out) if (r <= 0) throw new LZ4Exception("maxDestLen is too small");
    return out;
  }

  /** What {@link #compressPacked} returns: per-block offsets (absolute in dst) and sizes, and the packed piece of each GPU. */
  public static final class Packed {
    public final long[] offset; public final int[] length; public final long[] shardBase; public final long[] shardLength;
    Packed(long[] o, int[] l, long[] b, long[] t) { offset = o; length = l; shardBase = b; shardLength = t; }
  }

  /** Compresses n equally sized blocks on {@code gpus} devices and packs each device's blocks back to back in dst (only
   *  compressed bytes cross PCIe).  The pieces [shardBase[g], shardBase[g] + shardLength[g]) written one after the other
   *  are the packed stream; dst must hold n slots of maxCompressedLength(blockSize) rounded up to 16. */
  public static Packed compressPacked(ByteBuffer src, int blockSize, int n, ByteBuffer dst, int gpus) {
    final LongBuffer so = longs(n), oo = longs(n), sb = longs(gpus), st = longs(gpus);
    final IntBuffer sl = ints(n), res = ints(n);
    for (int i = 0; i < n; i++) { so.put(i, (long) i * blockSize); sl.put(i, blockSize); }
    final int rc = LZ4B200JNI.compressPackedMulti(src, so, sl, dst, oo, res, n, blockSize, null, gpus, sb, st);
    if (rc != 0) throw new LZ4Exception("B200 backend error " + rc);
    final long[] off = new long[n]; final int[] len = new int[n]; final long[] base = new long[gpus]; final long[] tot = new long[gpus];
    oo.get(off); res.get(len); sb.get(base); st.get(tot);
    for (int r : len) if (r <= 0) throw new LZ4Exception("maxDestLen is too small");
    return new Packed(off, len, base, tot);
  }

  /** Inverse of the sharded {@link #compressUniform(ByteBuffer, int, int, ByteBuffer, int)}. */
  public static void decompressUniform(ByteBuffer src, int[] compressedLen, int blockSize, ByteBuffer dst, int gpus) {
    final int n = compressedLen.length;
    final int bound = LZ4Utils.maxCompressedLength(blockSize);
    final LongBuffer so = longs(n), dof = longs(n);
    final IntBuffer sa = ints(n), dl = ints(n), res = ints(n);
    for (int i = 0; i < n; i++) { so.put(i, (long) i * bound); sa.put(i, bound); dof.put(i, (long) i * blockSize); dl.put(i, blockSize); }
    final int rc = LZ4B200JNI.decompressFastBatchMulti(src, so, sa, dst, dof, dl, res, n, null, gpus);
    if (rc != 0) throw new LZ4Exception("B200 backend error " + rc);
    for (int i = 0; i < n; i++) if (res.get(i) != compressedLen[i]) throw new LZ4Exception("Error decoding block " + i);
  }

  /** Compresses n equally sized blocks laid out back to back in src into bound-sized slots of dst; returns per-block sizes. */
  public static int[] compressUniform(ByteBuffer src, int blockSize, int n, ByteBuffer dst) {
    final int bound = LZ4Utils.maxCompressedLength(blockSize);
    final LongBuffer so = longs(n), dof = longs(n);
    final IntBuffer sl = ints(n), dc = ints(n), res = ints(n);
    for (int i = 0; i < n; i++) { so.put(i, (long) i * blockSize); sl.put(i, blockSize); dof.put(i, (long) i * bound); dc.put(i, bound); }
    final int rc = LZ4B200JNI.compressBatch(src, so, sl, dst, dof, dc, res, n, blockSize);
    if (rc != 0) throw new LZ4Exception("B200 backend error " + rc);
    final int[] out = new int[n];
    res.get(out);
    for (int r : out) if (r <= 0) throw new LZ4Exception("maxDestLen is too small");
    return out;
  }

  /** Inverse of {@link #compressUniform}: fast-decompresses n slots into blockSize-byte blocks. */
  public static void decompressUniform(ByteBuffer src, int[] compressedLen, int blockSize, ByteBuffer dst) {
    final int n = compressedLen.length;
    final int bound = LZ4Utils.maxCompressedLength(blockSize);
    final LongBuffer so = longs(n), dof = longs(n);
    final IntBuffer sa = ints(n), dl = ints(n), res = ints(n);
    for (int i = 0; i < n; i++) { so.put(i, (long) i * bound); sa.put(i, bound); dof.put(i, (long) i * blockSize); dl.put(i, blockSize); }
    final int rc = LZ4B200JNI.decompressFastBatch(src, so, sa, dst, dof, dl, res, n);
    if (rc != 0) throw new LZ4Exception("B200 backend error " + rc);
    for (int i = 0; i < n; i++) if (res.get(i) != compressedLen[i]) throw new LZ4Exception("Error decoding block " + i);
  }
}
