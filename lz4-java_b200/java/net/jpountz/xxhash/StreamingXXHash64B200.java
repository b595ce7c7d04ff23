package net.jpountz.xxhash;

/**
 * Streaming XXH64 of the B200 backend (replaces StreamingXXHash64JNI.java for the "B200" implementation string); see
 * {@link StreamingXXHash32B200} and {@link B200StreamState}.
 */
final class StreamingXXHash64B200 extends StreamingXXHash64 {

  static final class Factory implements StreamingXXHash64.Factory {
    public static final StreamingXXHash64.Factory INSTANCE = new Factory();

    @Override
    public StreamingXXHash64 newStreamingHash(long seed) {
      return new StreamingXXHash64B200(seed);
    }
  }

  private final B200StreamState device;

  StreamingXXHash64B200(long seed) {
    super(seed);
    device = new B200StreamState(true, seed);
  }

  @Override
  public void update(byte[] bytes, int off, int len) {
    device.update(bytes, off, len);
  }

  @Override
  public long getValue() {
    return device.digest();
  }

  @Override
  public void reset() {
    device.reset(seed);
  }

  @Override
  public void close() {
    super.close();
    device.release();
  }

  @Override
  protected void finalize() throws Throwable {
    try {
      device.release();
    } finally {
      super.finalize();
    }
  }
}
