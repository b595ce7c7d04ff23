package net.jpountz.xxhash;

/** Streaming XXH32 whose state lives on the device (twin of StreamingXXHash32JNI.java:24-93). */
final class StreamingXXHash32B200 extends StreamingXXHash32 {

  static class Factory implements StreamingXXHash32.Factory {
    public static final StreamingXXHash32.Factory INSTANCE = new Factory();
    @Override
    public StreamingXXHash32 newStreamingHash(int seed) { return new StreamingXXHash32B200(seed); }
  }

  private long state;

  StreamingXXHash32B200(int seed) {
    super(seed);
    state = XXHashB200JNI.XXH32_init(seed);
  }

  private void checkState() {
    if (state == 0) throw new AssertionError("Already finalized");
  }

  @Override
  public synchronized void reset() {
    checkState();
    XXHashB200JNI.XXH32_free(state);
    state = XXHashB200JNI.XXH32_init(seed);
  }

  @Override
  public synchronized int getValue() {
    checkState();
    return XXHashB200JNI.XXH32_digest(state);
  }

  @Override
  public synchronized void update(byte[] bytes, int off, int len) {
    checkState();
    XXHashB200JNI.XXH32_update(state, bytes, off, len);
  }

  @Override
  public synchronized void close() {
    if (state != 0) {
      super.close();
      XXHashB200JNI.XXH32_free(state);
      state = 0;
    }
  }

  @Override
  protected synchronized void finalize() throws Throwable {
    super.finalize();
    if (state != 0) { XXHashB200JNI.XXH32_free(state); state = 0; }
  }
}
