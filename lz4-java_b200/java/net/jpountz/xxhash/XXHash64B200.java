package net.jpountz.xxhash;

import static net.jpountz.util.ByteBufferUtils.checkRange;
import static net.jpountz.util.SafeUtils.checkRange;

import java.nio.ByteBuffer;

/** {@link XXHash64} on the B200 backend (twin of XXHash64JNI.java:24-51); resolved as XXHash64<impl>.INSTANCE. */
final class XXHash64B200 extends XXHash64 {

  public static final XXHash64 INSTANCE = new XXHash64B200();

  @Override
  public long hash(byte[] buf, int off, int len, long seed) {
    checkRange(buf, off, len);
    return XXHashB200JNI.XXH64(buf, off, len, seed);
  }

  @Override
  public long hash(ByteBuffer buf, int off, int len, long seed) {
    if (buf.isDirect()) {
      checkRange(buf, off, len);
      return XXHashB200JNI.XXH64BB(buf, off, len, seed);
    } else if (buf.hasArray()) {
      return hash(buf.array(), off + buf.arrayOffset(), len, seed);
    }
    throw new IllegalArgumentException("B200 backend needs array-backed or direct buffers");
  }
}
