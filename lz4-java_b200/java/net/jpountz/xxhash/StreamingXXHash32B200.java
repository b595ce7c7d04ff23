package net.jpountz.xxhash;

/**
 * Streaming XXH32 of the B200 backend: the accumulators stay on the device between update() calls (replaces
 * StreamingXXHash32JNI.java for the "B200" implementation string; resolved by XXHashFactory.java:179-182 through the
 * nested Factory).  State handling is shared with the 64-bit class in {@link B200StreamState}.
 */
final class StreamingXXHash32B200 extends StreamingXXHash32 {

  static final class Factory implements StreamingXXHash32.Factory {
    public static final StreamingXXHash32.Factory INSTANCE = new Factory();

    @Override
    public StreamingXXHash32 newStreamingHash(int seed) {
      return new StreamingXXHash32B200(seed);
    }
  }

  private final B200StreamState device;

  StreamingXXHash32B200(int seed) {
    super(seed);
    device = new B200StreamState(false, seed);
  }

  @Override
  public void update(byte[] bytes, int off, int len) {
    device.update(bytes, off, len);
  }

  @Override
  public int getValue() {
    return (int) device.digest();
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
