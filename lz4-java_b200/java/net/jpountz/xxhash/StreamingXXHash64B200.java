package net.jpountz.xxhash;

/** Streaming XXH64 whose state lives on the device (twin of StreamingXXHash64JNI.java:24-93). */
final class StreamingXXHash64B200 extends StreamingXXHash64 {

  static class Factory implements StreamingXXHash64.Factory {
    public static final StreamingXXHash64.Factory INSTANCE = new Factory();
    @Override
    public StreamingXXHash64 newStreamingHash(long seed) { return new StreamingXXHash64B200(seed); }
  }

  private long state;

  StreamingXXHash64B200(long seed) {
    super(seed);
    state = XXHashB200JNI.XXH64_init(seed);
  }

  private void checkState() {
    if (state == 0) throw new AssertionError("Already finalized");
  }

  @Override
  public synchronized void reset() {
    checkState();
    XXHashB200JNI.XXH64_free(state);
    state = XXHashB200JNI.XXH64_init(seed);
  }

  @Override
  public synchronized long getValue() {
    checkState();
    return XXHashB200JNI.XXH64_digest(state);
  }

  @Override
  public synchronized void update(byte[] bytes, int off, int len) {
    checkState();
    XXHashB200JNI.XXH64_update(state, bytes, off, len);
  }

  @Override
  public synchronized void close() {
    if (state != 0) {
      super.close();
      XXHashB200JNI.XXH64_free(state);
      state = 0;
    }
  }

  @Override
  protected synchronized void finalize() throws Throwable {
    super.finalize();
    if (state != 0) { XXHashB200JNI.XXH64_free(state); state = 0; }
  }
}
