/*
 * b200lz4.h — C ABI of libb200lz4.so, the B200 (sm_100a) LZ4 block codec + XXHash backend.
 *
 * This is the drop-in boundary for the one hot path of lz4-java: the functions below are
 * what a `net.jpountz.lz4.LZ4B200JNI` / `net.jpountz.xxhash.XXHashB200JNI` shim binds, in
 * the same way the reference's JNI shim binds the vendored C:
 *
 *   reference call site (under /root/reference)                       replaced by
 *   src/jni/net_jpountz_lz4_LZ4JNI.c:75   LZ4_compress_default        b200lz4_compress_default
 *   src/jni/net_jpountz_lz4_LZ4JNI.c:122  LZ4_compress_HC             b200lz4_compress_HC
 *   src/jni/net_jpountz_lz4_LZ4JNI.c:169  LZ4_decompress_fast         b200lz4_decompress_fast_bounded
 *   src/jni/net_jpountz_lz4_LZ4JNI.c:216  LZ4_decompress_safe         b200lz4_decompress_safe
 *   src/jni/net_jpountz_lz4_LZ4JNI.c:237  LZ4_compressBound           b200lz4_compressBound
 *   src/jni/net_jpountz_xxhash_XXHashJNI.c:54,78    XXH32             b200xxh32
 *   src/jni/net_jpountz_xxhash_XXHashJNI.c:164,188  XXH64             b200xxh64
 *   src/jni/net_jpountz_xxhash_XXHashJNI.c:89-145   XXH32_* state     b200xxh32_create/reset/update/digest/free
 *   src/jni/net_jpountz_xxhash_XXHashJNI.c:199-255  XXH64_* state     b200xxh64_create/reset/update/digest/free
 *
 * All arithmetic runs in hand-written CUDA kernels; there is NO CPU fallback: every entry
 * point returns B200LZ4_E_NODEVICE (or 0 for the hash one-shots, with b200lz4_last_error()
 * set) when no sm_100 device/driver is usable.
 *
 * Return conventions are the reference's:
 *   compress          > 0 compressed size, 0 if dst is too small / input too large
 *   decompress_safe   >= 0 decoded size, < 0 == -(error position)-1   (lz4.c:2337)
 *   decompress_fast   >= 0 compressed bytes consumed, -1 on any error  (lz4.c:1890)
 *
 * Besides the one-block-per-call functions (what the net.jpountz API needs, n = 1), the
 * library exposes BATCH calls: one launch over n independent blocks.  These are the calls
 * that make a GPU backend meaningful (SURVEY.md §7 "hard parts" 1); the single-block calls
 * are the n = 1 case of the host batch path.
 */
#ifndef B200LZ4_H
#define B200LZ4_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200LZ4_VERSION        100          /* 0.1.0 */
#define B200LZ4_E_NODEVICE     (-2147483647)       /* INT_MIN + 1: no usable CUDA device / driver            */
#define B200LZ4_E_CUDA         (-2147483646)       /* INT_MIN + 2: CUDA runtime error, see b200lz4_last_error */
#define B200LZ4_E_ARG          (-2147483645)       /* INT_MIN + 3: invalid argument                           */

/* ---------------------------------------------------------------- library / device */
int         b200lz4_version(void);
int         b200lz4_device_count(void);           /* >= 0, or B200LZ4_E_NODEVICE            */
int         b200lz4_set_device(int device);       /* device used by the calling thread      */
const char* b200lz4_last_error(void);             /* thread-local, never NULL                */
/* The hash entry points return the hash VALUE (like XXH32/XXH64, xxhash.c:392,855), so they cannot return an error
 * code: b200xxh32 / b200xxh64 / b200xxh*_digest clear this thread-local status on entry and leave a B200LZ4_E_* in it
 * when they could not compute (they then return 0).  A binding must check it after each of those calls and throw. */
int         b200lz4_last_status(void);
/* Pin / unpin a caller-owned host range (e.g. a Java DirectByteBuffer) so the host batch
 * calls DMA straight from/to it.  Optional: unpinned memory works, slower. */
int         b200lz4_host_register(void* p, size_t bytes);
int         b200lz4_host_unregister(void* p);

/* ---------------------------------------------------------------- one block per call, HOST buffers */
int b200lz4_compressBound(int inputSize);                                            /* lz4.h:212 */
int b200lz4_compress_default(const char* src, char* dst, int srcSize, int dstCapacity);
int b200lz4_compress_HC(const char* src, char* dst, int srcSize, int dstCapacity, int level);
int b200lz4_decompress_safe(const char* src, char* dst, int compressedSize, int dstCapacity);
/* LZ4_decompress_fast does not know the input size (lz4.c:1788); a device copy needs one.
 * `srcAvail` is the number of readable bytes at src (the Java wrapper knows it:
 * src.length - srcOff).  There is deliberately no entry point without it: the host would have to
 * read compressBound(originalSize) bytes from src, past the end of most callers' buffers.
 * One-block calls copy back exactly the bytes they produced: dst[result, dstCapacity) is not touched. */
int b200lz4_decompress_fast_bounded(const char* src, int srcAvail, char* dst, int originalSize);

uint32_t b200xxh32(const void* input, size_t len, uint32_t seed);
uint64_t b200xxh64(const void* input, size_t len, uint64_t seed);

/* streaming hash state: opaque handle (jlong on the Java side) */
void*    b200xxh32_create(uint32_t seed);
void     b200xxh32_reset(void* state, uint32_t seed);
int      b200xxh32_update(void* state, const void* input, size_t len);
uint32_t b200xxh32_digest(void* state);
void     b200xxh32_free(void* state);
void*    b200xxh64_create(uint64_t seed);
void     b200xxh64_reset(void* state, uint64_t seed);
int      b200xxh64_update(void* state, const void* input, size_t len);
uint64_t b200xxh64_digest(void* state);
void     b200xxh64_free(void* state);

/* ---------------------------------------------------------------- batches, DEVICE-resident
 * Every pointer is a device pointer on the current device; `stream` is a cudaStream_t
 * (NULL = default stream).  Block i reads  src_base + src_off[i]  (src_len[i] bytes) and
 * writes dst_base + dst_off[i] (at most dst_cap[i] bytes); result[i] follows the
 * single-block return convention.  The calls are asynchronous on `stream` and return 0 or
 * a B200LZ4_E_* code for launch failures.
 *
 * compress_fast: `max_src_len` is an upper bound on src_len[] chosen by the caller
 * (<= 65536 selects the 16-bit position table like lz4.c:1353; 0 = unknown -> 32-bit). */
int b200lz4_compress_fast_batch_dev(const uint8_t* src_base, const uint64_t* src_off, const int32_t* src_len,
                                    uint8_t* dst_base, const uint64_t* dst_off, const int32_t* dst_cap,
                                    int32_t* result, size_t n, int max_src_len, void* stream);
int b200lz4_compress_hc_batch_dev(const uint8_t* src_base, const uint64_t* src_off, const int32_t* src_len,
                                  uint8_t* dst_base, const uint64_t* dst_off, const int32_t* dst_cap,
                                  int32_t* result, size_t n, int level, void* stream);
int b200lz4_decompress_safe_batch_dev(const uint8_t* src_base, const uint64_t* src_off, const int32_t* src_len,
                                      uint8_t* dst_base, const uint64_t* dst_off, const int32_t* dst_cap,
                                      int32_t* result, size_t n, void* stream);
/* decompress_fast: src_avail[i] = readable bytes at the block's src; dst_len[i] = exact original size */
int b200lz4_decompress_fast_batch_dev(const uint8_t* src_base, const uint64_t* src_off, const int32_t* src_avail,
                                      uint8_t* dst_base, const uint64_t* dst_off, const int32_t* dst_len,
                                      int32_t* result, size_t n, void* stream);
/* hashes: len[] may be anything >= 0; one seed for the batch (the Java API passes one seed per call) */
int b200xxh32_batch_dev(const uint8_t* base, const uint64_t* off, const int32_t* len,
                        uint32_t seed, uint32_t* out, size_t n, void* stream);
int b200xxh64_batch_dev(const uint8_t* base, const uint64_t* off, const int32_t* len,
                        uint64_t seed, uint64_t* out, size_t n, void* stream);

/* Packing on the device -- what *_compact_host does before its copy back, for callers whose blocks stay in HBM: block i's
 * lens[i] bytes (<= 0 counts as none) move from slots + slot_off[i] to out + out_off[i]; out_off[] = the exclusive prefix
 * sums of lens[], *total = their sum (device pointers all; out needs sum(lens) bytes).  Asynchronous on `stream`. */
int b200lz4_compact_dev(const uint8_t* slots, const uint64_t* slot_off, const int32_t* lens, uint8_t* out, uint64_t* out_off,
                        uint64_t* total, size_t n, void* stream);
/* Device-side stitch of per-GPU packed shards into one stream: shard g is shard_total[g] packed bytes
 * at shard_ptr[g] in the memory of device shard_dev[g] (a HOST array of device pointers; totals on the host -- 8 bytes per
 * GPU, the only thing exchanged).  Shard g lands in dst (memory of device dst_dev, dst_capacity bytes) at the sum of the totals
 * before it, reported in shard_pos[g] (host, may be NULL): one peer copy per shard, each on a stream of its source device, all
 * in flight together (NVLink between peers, otherwise staged by the driver).  A block at out_off[i] inside shard g is at
 * shard_pos[g] + out_off[i] in dst.  The shards must be complete before the call (synchronise their producers); returns when
 * every copy has landed.  Needs no worker threads and no host buffer; the caller's current device is left as it was. */
int b200lz4_stitch_shards_dev(const void* const* shard_ptr, const int* shard_dev, const uint64_t* shard_total, int nshard,
                              void* dst, int dst_dev, size_t dst_capacity, uint64_t* shard_pos);

/* ---------------------------------------------------------------- batches, HOST buffers
 * Same contracts, but every pointer is a HOST pointer.  The library chunks the batch and
 * pipelines H2D copy / kernel / D2H copy on several streams of the current device.  Blocks
 * must be laid out in ascending, non-overlapping order in both src and dst (the natural
 * layout of a block list over a DirectByteBuffer).  Synchronous: returns when all results
 * and output bytes are in host memory.  Returns 0 or a B200LZ4_E_* code. */
int b200lz4_compress_fast_batch_host(const uint8_t* src_base, const uint64_t* src_off, const int32_t* src_len,
                                     uint8_t* dst_base, const uint64_t* dst_off, const int32_t* dst_cap,
                                     int32_t* result, size_t n, int max_src_len);
int b200lz4_compress_hc_batch_host(const uint8_t* src_base, const uint64_t* src_off, const int32_t* src_len,
                                   uint8_t* dst_base, const uint64_t* dst_off, const int32_t* dst_cap,
                                   int32_t* result, size_t n, int level);
int b200lz4_decompress_safe_batch_host(const uint8_t* src_base, const uint64_t* src_off, const int32_t* src_len,
                                       uint8_t* dst_base, const uint64_t* dst_off, const int32_t* dst_cap,
                                       int32_t* result, size_t n);
int b200lz4_decompress_fast_batch_host(const uint8_t* src_base, const uint64_t* src_off, const int32_t* src_avail,
                                       uint8_t* dst_base, const uint64_t* dst_off, const int32_t* dst_len,
                                       int32_t* result, size_t n);
int b200xxh32_batch_host(const uint8_t* base, const uint64_t* off, const int32_t* len,
                         uint32_t seed, uint32_t* out, size_t n);
int b200xxh64_batch_host(const uint8_t* base, const uint64_t* off, const int32_t* len,
                         uint64_t seed, uint64_t* out, size_t n);

/* Compact-output compress for pipelines that want one contiguous stream (e.g. a frame
 * writer): blocks are compressed and written back-to-back into dst_base; out_off[i] and
 * result[i] give each block's position and size; *total = bytes written.  dst_capacity
 * must be >= sum(compressBound(src_len[i])) only in the worst case; the call fails with
 * B200LZ4_E_ARG if the compacted stream does not fit. */
int b200lz4_compress_fast_compact_host(const uint8_t* src_base, const uint64_t* src_off, const int32_t* src_len,
                                       uint8_t* dst_base, size_t dst_capacity, uint64_t* out_off,
                                       int32_t* result, size_t n, int max_src_len, uint64_t* total);

/* ---------------------------------------------------------------- host batches, range-sharded over several GPUs
 * ONE call from ONE process (a JVM) drives `ndev` GPUs: GPU devices[g] takes the contiguous block range
 * [g*n/ndev, (g+1)*n/ndev) (sizes differ by at most one block) and runs the host pipeline above on its own streams
 * from its own worker thread; blocks are independent, so there is no exchange between the GPUs and no collective
 * (SURVEY.md 8e).  `devices` = NULL means devices 0..ndev-1.  Same layout rules and results as the single-GPU calls
 * (result[] / out[] are filled for all n blocks; byte-identical output whatever ndev is).  Returns 0, or the first
 * B200LZ4_E_* any shard reported (b200lz4_last_error() then names the device). */
int b200lz4_compress_fast_batch_host_multi(const uint8_t* src_base, const uint64_t* src_off, const int32_t* src_len,
                                           uint8_t* dst_base, const uint64_t* dst_off, const int32_t* dst_cap,
                                           int32_t* result, size_t n, int max_src_len, const int* devices, int ndev);
int b200lz4_compress_hc_batch_host_multi(const uint8_t* src_base, const uint64_t* src_off, const int32_t* src_len,
                                         uint8_t* dst_base, const uint64_t* dst_off, const int32_t* dst_cap,
                                         int32_t* result, size_t n, int level, const int* devices, int ndev);
int b200lz4_decompress_safe_batch_host_multi(const uint8_t* src_base, const uint64_t* src_off, const int32_t* src_len,
                                             uint8_t* dst_base, const uint64_t* dst_off, const int32_t* dst_cap,
                                             int32_t* result, size_t n, const int* devices, int ndev);
int b200lz4_decompress_fast_batch_host_multi(const uint8_t* src_base, const uint64_t* src_off, const int32_t* src_avail,
                                             uint8_t* dst_base, const uint64_t* dst_off, const int32_t* dst_len,
                                             int32_t* result, size_t n, const int* devices, int ndev);
/* Packed output from several GPUs: shard g's compressed blocks are packed back to back (like
 * b200lz4_compress_fast_compact_host) starting at dst_base + shard_base[g], where shard_base[g] is the sum of the 16-byte
 * aligned bounds of the blocks before the shard — known before anything is compressed, so the GPUs never wait for each
 * other.  out_off[i] is absolute in dst_base; shard_total[g] = packed bytes of shard g.  The stream is contiguous within a
 * shard and has a gap between shards: a caller writes the ndev pieces one after the other (gather write).  dst_capacity
 * must hold the aligned bounds of all blocks.  shard_base / shard_total: ndev entries each (may be NULL). */
int b200lz4_compress_fast_compact_host_multi(const uint8_t* src_base, const uint64_t* src_off, const int32_t* src_len,
                                             uint8_t* dst_base, size_t dst_capacity, uint64_t* out_off,
                                             int32_t* result, size_t n, int max_src_len, const int* devices, int ndev,
                                             uint64_t* shard_base, uint64_t* shard_total);
int b200xxh32_batch_host_multi(const uint8_t* base, const uint64_t* off, const int32_t* len, uint32_t seed,
                               uint32_t* out, size_t n, const int* devices, int ndev);
int b200xxh64_batch_host_multi(const uint8_t* base, const uint64_t* off, const int32_t* len, uint64_t seed,
                               uint64_t* out, size_t n, const int* devices, int ndev);

/* ---------------------------------------------------------------- LZ4 Frame batch decoder
 * LZ4FrameInputStream semantics (src/java/net/jpountz/lz4/LZ4FrameInputStream.java:132-321) over a buffer
 * of concatenated frames: the host indexes the container, the device verifies header/block/content
 * XXH32 checksums and decodes all blocks of all frames in batched launches.
 * Error codes (negative): -1 premature end, -2 bad magic, -3 descriptor checksum, -4 block larger than the
 * frame's maximum, -5 block checksum, -6 block decode error, -7 content checksum, -8 content size,
 * -9 dst too small, -10 unsupported descriptor, -11 (decode_dev only) every check passed but a frame has a short block
 * before its last one (flush()), so its content is not one run inside d_slots: read block b at block_off[b] for
 * block_len_out[b] bytes (b200lz4f_index_block_offsets).
 * Errors come in STREAM order, as the reader would meet them: frame by frame the descriptor hash, block by block its checksum
 * and its decode, at the EndMark content checksum then content size; a container that is cut short or malformed behind at
 * least one frame header still gets an index, what precedes the bad spot is decoded and verified first, and decode_dev then
 * returns the container's own code (-1, -2, -4, -10).
 * d_slots layout: a full block takes blockMaxSize bytes; a block that cannot fill it (a stored block, or a compressed one
 * of fewer than blockMaxSize/255 bytes) takes what it can decode to, rounded up to 16 -- slot_bytes does not grow with
 * blockMaxSize for a stream of tiny blocks. */
int64_t b200lz4f_decompress_host(const uint8_t* src, size_t srcSize, uint8_t* dst, size_t dstCapacity);
void*   b200lz4f_index_create(const uint8_t* src_host, size_t srcSize, uint64_t* slot_bytes, int* err);
/* readSingleFrame = true (LZ4FrameInputStream.java:83-91,118-123): reading stops behind the first non-skippable frame, the
 * rest of src is not looked at; *src_consumed (may be NULL) = how far the reader got. */
int64_t b200lz4f_decompress_host_single(const uint8_t* src, size_t srcSize, uint8_t* dst, size_t dstCapacity, size_t* src_consumed);
void*   b200lz4f_index_create_single(const uint8_t* src_host, size_t srcSize, uint64_t* slot_bytes, size_t* src_consumed, int* err);
/* getExpectedContentSize / isExpectedContentSizeDefined (:416-445): *content_size = what the first non-skippable frame's
 * descriptor declares, -1 if it declares none (or there are only skippable frames).  Returns 0 or a code from the list above
 * (the descriptor hash is checked, on the device like every hash here). */
int     b200lz4f_expected_content_size(const uint8_t* src, size_t srcSize, int64_t* content_size);
size_t  b200lz4f_index_frames(void* index);
size_t  b200lz4f_index_blocks(void* index);
void    b200lz4f_index_block_offsets(void* index, uint64_t* block_off);   /* b200lz4f_index_blocks() entries, bytes into d_slots */
int64_t b200lz4f_decode_dev(void* index, const uint8_t* d_src, uint8_t* d_slots, uint64_t* frame_off, uint64_t* frame_len,
                            int32_t* block_len_out, void* stream);
void    b200lz4f_index_free(void* index);

/* ---------------------------------------------------------------- lz4-java's containers as whole-buffer calls
 * LZ4 Frame writer (LZ4FrameOutputStream.java:178-251): independent blocks of 64 KiB..4 MiB (bsCode 4..7), blocks that
 * do not shrink are stored raw; flags bit0 = content checksum, bit1 = block checksums, bit2 = content size.
 * "LZ4Block" container (LZ4BlockOutputStream.java:203-266 / LZ4BlockInputStream.java:191-264): 21-byte block
 * headers, XXH32 (seed 0x9747b28c, 28-bit) of each original block, fast decompressor on the read side.
 * Length-prefixed blocks (LZ4CompressorWithLength / LZ4DecompressorWithLength).
 * Return: bytes written / decoded, or negative: -1 premature end, -2 corrupted, -9 dst too small, B200LZ4_E_*. */
size_t  b200lz4f_compress_bound(size_t srcSize, int bsCode);
int64_t b200lz4f_compress_host(const uint8_t* src, size_t srcSize, uint8_t* dst, size_t dstCapacity, int bsCode, int flags);
/* the writers' LZ4Compressor argument (LZ4FrameOutputStream.java:132-133, LZ4BlockOutputStream.java:96,124): hc_level 0 = the
 * fast compressor (what the calls without _hc use), 1..17 = LZ4_compress_HC at that level */
int64_t b200lz4f_compress_host_hc(const uint8_t* src, size_t srcSize, uint8_t* dst, size_t dstCapacity, int bsCode, int flags, int hc_level);
size_t  b200lz4block_compress_bound(size_t srcSize, int blockSize);
int64_t b200lz4block_compress_host(const uint8_t* src, size_t srcSize, uint8_t* dst, size_t dstCapacity, int blockSize);
int64_t b200lz4block_compress_host_hc(const uint8_t* src, size_t srcSize, uint8_t* dst, size_t dstCapacity, int blockSize, int hc_level);
/* stopOnEmptyBlock: LZ4BlockInputStream's constructor flag (LZ4BlockInputStream.java:60-72; its default is true): non-zero
 * ends at the first empty block and reports in *srcConsumed (may be NULL) how far it read; zero steps over empty blocks and
 * reads concatenated streams to the end of src. */
int64_t b200lz4block_decompress_host(const uint8_t* src, size_t srcSize, uint8_t* dst, size_t dstCapacity,
                                     int stopOnEmptyBlock, size_t* srcConsumed);
int     b200lz4_compress_with_length(const char* src, char* dst, int srcSize, int dstCapacity);
int     b200lz4_decompressed_length(const char* src);
int     b200lz4_decompress_with_length(const char* src, int srcAvail, char* dst, int dstCapacity);

/* kernel-launch counter (bench.py's "gpu_launches"): number of kernels this library has
 * launched from the calling process since load / since the last reset. */
uint64_t b200lz4_launch_count(void);
void     b200lz4_launch_count_reset(void);

/* pipeline contexts (3 streams + device staging each) created in this process so far.  A thread keeps one per device;
 * contexts of threads that exited are reused by new threads, so with the reference's usage (any number of Java threads
 * calling the singleton codecs, LZ4Compressor.java:25) this stays at the peak number of CONCURRENT callers. */
int      b200lz4_context_count(void);

#ifdef __cplusplus
}
#endif
#endif /* B200LZ4_H */
