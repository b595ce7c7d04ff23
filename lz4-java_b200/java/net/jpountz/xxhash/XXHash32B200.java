package net.jpountz.xxhash;

import static net.jpountz.util.ByteBufferUtils.checkRange;
import static net.jpountz.util.SafeUtils.checkRange;

import java.nio.ByteBuffer;

/** {@link XXHash32} on the B200 backend (twin of XXHash32JNI.java:24-51); resolved as XXHash32<impl>.INSTANCE. */
final class XXHash32B200 extends XXHash32 {

  public static final XXHash32 INSTANCE = new XXHash32B200();

  @Override
  public int hash(byte[] buf, int off, int len, int seed) {
    checkRange(buf, off, len);
    return XXHashB200JNI.XXH32(buf, off, len, seed);
  }

  @Override
  public int hash(ByteBuffer buf, int off, int len, int seed) {
    if (buf.isDirect()) {
      checkRange(buf, off, len);
      return XXHashB200JNI.XXH32BB(buf, off, len, seed);
    } else if (buf.hasArray()) {
      return hash(buf.array(), off + buf.arrayOffset(), len, seed);
    }
    throw new IllegalArgumentException("B200 backend needs array-backed or direct buffers");
  }
}
