package net.jpountz.lz4;

import static net.jpountz.util.ByteBufferUtils.checkNotReadOnly;
import static net.jpountz.util.ByteBufferUtils.checkRange;
import static net.jpountz.util.SafeUtils.checkRange;

import java.nio.ByteBuffer;

/** {@link LZ4SafeDecompressor} on the B200 backend (twin of LZ4JNISafeDecompressor.java:29-83). Returns bytes WRITTEN. */
final class LZ4B200SafeDecompressor extends LZ4SafeDecompressor {

  public static final LZ4B200SafeDecompressor INSTANCE = new LZ4B200SafeDecompressor();

  @Override
  public final int decompress(byte[] src, int srcOff, int srcLen, byte[] dest, int destOff, int maxDestLen) {
    checkRange(src, srcOff, srcLen);
    checkRange(dest, destOff, maxDestLen);
    final int result = LZ4B200JNI.LZ4_decompress_safe(src, null, srcOff, srcLen, dest, null, destOff, maxDestLen);
    if (result <= Integer.MIN_VALUE + 3) {
      throw new LZ4Exception("B200 backend error " + result);   // B200LZ4_E_* (no device / CUDA error): never a CPU fallback
    }
    if (result < 0) {
      throw new LZ4Exception("Error decoding offset " + (srcOff - result) + " of input buffer");
    }
    return result;
  }

  @Override
  public int decompress(ByteBuffer src, int srcOff, int srcLen, ByteBuffer dest, int destOff, int maxDestLen) {
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
    final int result = LZ4B200JNI.LZ4_decompress_safe(srcArr, srcArr == null ? src : null, so, srcLen,
        destArr, destArr == null ? dest : null, dof, maxDestLen);
    if (result <= Integer.MIN_VALUE + 3) {
      throw new LZ4Exception("B200 backend error " + result);   // B200LZ4_E_* (no device / CUDA error): never a CPU fallback
    }
    if (result < 0) {
      throw new LZ4Exception("Error decoding offset " + (srcOff - result) + " of input buffer");
    }
    return result;
  }
}
