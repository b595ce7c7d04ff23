package net.jpountz.lz4;

import static net.jpountz.util.ByteBufferUtils.checkNotReadOnly;
import static net.jpountz.util.ByteBufferUtils.checkRange;
import static net.jpountz.util.SafeUtils.checkRange;

import java.nio.ByteBuffer;

/**
 * Fast {@link LZ4Compressor} on the B200 backend.  Resolved by {@code LZ4Factory.instance("B200")}
 * through the class-name convention (LZ4Factory.java:176-196): {@code LZ4<impl>Compressor.INSTANCE}.
 * No CPU fallback: buffers that are neither array-backed nor direct are rejected.
 */
final class LZ4B200Compressor extends LZ4Compressor {

  public static final LZ4Compressor INSTANCE = new LZ4B200Compressor();

  @Override
  public int compress(byte[] src, int srcOff, int srcLen, byte[] dest, int destOff, int maxDestLen) {
    checkRange(src, srcOff, srcLen);
    checkRange(dest, destOff, maxDestLen);
    final int result = LZ4B200JNI.LZ4_compress_limitedOutput(src, null, srcOff, srcLen, dest, null, destOff, maxDestLen);
    if (result <= 0) {
      throw new LZ4Exception(result <= Integer.MIN_VALUE + 3 ? "B200 backend error " + result : "maxDestLen is too small");
    }
    return result;
  }

  @Override
  public int compress(ByteBuffer src, int srcOff, int srcLen, ByteBuffer dest, int destOff, int maxDestLen) {
    checkNotReadOnly(dest);
    checkRange(src, srcOff, srcLen);
    checkRange(dest, destOff, maxDestLen);
    if (!(src.hasArray() || src.isDirect()) || !(dest.hasArray() || dest.isDirect())) {
      throw new LZ4Exception("B200 backend needs array-backed or direct buffers");
    }
    final byte[] srcArr = src.hasArray() ? src.array() : null;
    final byte[] destArr = dest.hasArray() ? dest.array() : null;
    final int so = srcOff + (srcArr != null ? src.arrayOffset() : 0);
    final int dof = destOff + (destArr != null ? dest.arrayOffset() : 0);
    final int result = LZ4B200JNI.LZ4_compress_limitedOutput(srcArr, srcArr == null ? src : null, so, srcLen,
        destArr, destArr == null ? dest : null, dof, maxDestLen);
    if (result <= 0) {
      throw new LZ4Exception(result <= Integer.MIN_VALUE + 3 ? "B200 backend error " + result : "maxDestLen is too small");
    }
    return result;
  }
}
