package net.jpountz.xxhash;

/**
 * Owner of one device-resident streaming-hash state: the handle b200xxh32_create / b200xxh64_create return
 * (include/b200lz4.h; a 48- or 88-byte accumulator struct in HBM, one CUDA stream, a pinned 8-byte result slot).
 *
 * Both streaming classes of the B200 backend keep their state here, so the rules the reference spells out per class
 * (StreamingXXHash32JNI.java:24-32: every use must exclude the finalizer that frees the native memory) live in one
 * place: all access goes through this object's monitor, a released state answers with the reference's
 * AssertionError("Already finalized"), and a state the backend could not create (no GPU, out of device memory) is an
 * exception at construction, not a zero handle that fails later.
 */
final class B200StreamState {

  private final boolean wide;     // true: XXH64, false: XXH32
  private long handle;

  B200StreamState(boolean wide, long seed) {
    this.wide = wide;
    this.handle = wide ? XXHashB200JNI.XXH64_init(seed) : XXHashB200JNI.XXH32_init((int) seed);
    if (handle == 0) {
      throw new IllegalStateException("B200 backend: cannot create a streaming hash state (no device / out of memory)");
    }
  }

  private long live() {
    if (handle == 0) {
      throw new AssertionError("Already finalized");
    }
    return handle;
  }

  /** Back to the seeded initial state WITHOUT giving the device allocation back (the reference frees and re-creates,
   *  StreamingXXHash32JNI.java:54-58; here that would be a cudaFree + cudaMalloc + stream creation per reset). */
  synchronized void reset(long seed) {
    if (wide) {
      XXHashB200JNI.XXH64_reset(live(), seed);
    } else {
      XXHashB200JNI.XXH32_reset(live(), (int) seed);
    }
  }

  synchronized void update(byte[] bytes, int off, int len) {
    if (wide) {
      XXHashB200JNI.XXH64_update(live(), bytes, off, len);
    } else {
      XXHashB200JNI.XXH32_update(live(), bytes, off, len);
    }
  }

  /** Idempotent: the digest kernel reads the accumulators and leaves them untouched (XXHash32Test.java:51-53). */
  synchronized long digest() {
    return wide ? XXHashB200JNI.XXH64_digest(live()) : XXHashB200JNI.XXH32_digest(live()) & 0xFFFFFFFFL;
  }

  /** Frees the device state once; later calls are no-ops (close() then finalize() is the normal sequence). */
  synchronized void release() {
    final long h = handle;
    handle = 0;
    if (h != 0) {
      if (wide) {
        XXHashB200JNI.XXH64_free(h);
      } else {
        XXHashB200JNI.XXH32_free(h);
      }
    }
  }
}
