package net.jpountz.lz4;

import static net.jpountz.lz4.LZ4Constants.DEFAULT_COMPRESSION_LEVEL;
import static net.jpountz.util.ByteBufferUtils.checkNotReadOnly;
import static net.jpountz.util.ByteBufferUtils.checkRange;
import static net.jpountz.util.SafeUtils.checkRange;

import java.nio.ByteBuffer;

/** High-compression {@link LZ4Compressor} on the B200 backend (twin of LZ4HCJNICompressor.java:29-90). */
final class LZ4HCB200Compressor extends LZ4Compressor {

  public static final LZ4Compressor INSTANCE = new LZ4HCB200Compressor();

  private final int compressionLevel;

  LZ4HCB200Compressor() { this(DEFAULT_COMPRESSION_LEVEL); }
  LZ4HCB200Compressor(int compressionLevel) { this.compressionLevel = compressionLevel; }   // LZ4Factory.java:197-202

  @Override
  public int compress(byte[] src, int srcOff, int srcLen, byte[] dest, int destOff, int maxDestLen) {
    checkRange(src, srcOff, srcLen);
    checkRange(dest, destOff, maxDestLen);
    final int result = LZ4B200JNI.LZ4_compressHC(src, null, srcOff, srcLen, dest, null, destOff, maxDestLen, compressionLevel);
    if (result <= Integer.MIN_VALUE + 3) {
      throw new LZ4Exception("B200 backend error " + result);   // B200LZ4_E_*: no device / CUDA error
    }
    if (result <= 0) {
      throw new LZ4Exception();                                 // LZ4HCJNICompressor.java:47-49
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
    final int result = LZ4B200JNI.LZ4_compressHC(srcArr, srcArr == null ? src : null, so, srcLen,
        destArr, destArr == null ? dest : null, dof, maxDestLen, compressionLevel);
    if (result <= Integer.MIN_VALUE + 3) {
      throw new LZ4Exception("B200 backend error " + result);   // B200LZ4_E_*: no device / CUDA error
    }
    if (result <= 0) {
      throw new LZ4Exception();                                 // LZ4HCJNICompressor.java:47-49
    }
    return result;
  }
}
