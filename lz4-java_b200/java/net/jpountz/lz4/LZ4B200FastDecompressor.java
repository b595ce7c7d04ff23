package net.jpountz.lz4;

import static net.jpountz.util.ByteBufferUtils.checkNotReadOnly;
import static net.jpountz.util.ByteBufferUtils.checkRange;
import static net.jpountz.util.SafeUtils.checkRange;

import java.nio.ByteBuffer;

/** {@link LZ4FastDecompressor} on the B200 backend (twin of LZ4JNIFastDecompressor.java:29-84). Returns bytes READ. */
final class LZ4B200FastDecompressor extends LZ4FastDecompressor {

  public static final LZ4B200FastDecompressor INSTANCE = new LZ4B200FastDecompressor();

  @Override
  public final int decompress(byte[] src, int srcOff, byte[] dest, int destOff, int destLen) {
    checkRange(src, srcOff);
    checkRange(dest, destOff, destLen);
    final int result = LZ4B200JNI.LZ4_decompress_fast(src, null, srcOff, src.length - srcOff, dest, null, destOff, destLen);
    if (result <= Integer.MIN_VALUE + 3) {
      throw new LZ4Exception("B200 backend error " + result);   // B200LZ4_E_* (no device / CUDA error): never a CPU fallback
    }
    if (result < 0) {
      throw new LZ4Exception("Error decoding offset " + (srcOff - result) + " of input buffer");
    }
    return result;
  }

  @Override
  public int decompress(ByteBuffer src, int srcOff, ByteBuffer dest, int destOff, int destLen) {
    checkNotReadOnly(dest);
    checkRange(src, srcOff);
    checkRange(dest, destOff, destLen);
    if (!(src.hasArray() || src.isDirect()) || !(dest.hasArray() || dest.isDirect())) {
      throw new LZ4Exception("B200 backend needs array-backed or direct buffers");
    }
    final byte[] srcArr = src.hasArray() ? src.array() : null;
    final byte[] destArr = dest.hasArray() ? dest.array() : null;
    final int so = srcOff + (srcArr != null ? src.arrayOffset() : 0);
    final int dof = destOff + (destArr != null ? dest.arrayOffset() : 0);
    final int avail = src.capacity() - srcOff;
    final int result = LZ4B200JNI.LZ4_decompress_fast(srcArr, srcArr == null ? src : null, so, avail,
        destArr, destArr == null ? dest : null, dof, destLen);
    if (result <= Integer.MIN_VALUE + 3) {
      throw new LZ4Exception("B200 backend error " + result);   // B200LZ4_E_* (no device / CUDA error): never a CPU fallback
    }
    if (result < 0) {
      throw new LZ4Exception("Error decoding offset " + (srcOff - result) + " of input buffer");
    }
    return result;
  }
}
