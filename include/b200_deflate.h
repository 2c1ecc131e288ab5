/* b200_deflate.h -- C ABI of libb200deflate.so: the drop-in boundary for the DEFLATE / INFLATE hot
 * path of HyperBitGore/deflate.hpp, executed by hand-written sm_100a CUDA kernels.
 *
 * Plain pointers and sizes only; no C++ or torch types.  Every entry point names the reference
 * interface it replaces (file:line into the reference's include/).  The header-only C++17 drop-ins
 * include/deflate.hpp and include/inflate.hpp are thin wrappers over this file; INTEGRATION.md shows
 * the binding a maintainer of the reference would add.
 *
 * There is NO CPU fallback: every compute entry point returns B200_E_CUDA when no sm_100 device or
 * driver is usable.
 *
 * Stream format produced by the compressor (one valid RFC 1951 raw stream):
 *   input is cut into independent 64 KiB chunks; each chunk is ONE block (dynamic, fixed or stored,
 *   whichever is smallest by exact bit count; two blocks where the chunk's statistics change enough
 *   to pay for a second code table) that starts byte-aligned; every chunk except the last
 *   is followed by TWO empty non-final stored blocks (pad bits, 00 00 FF FF, 00, 00 00 FF FF) so
 *   that the next chunk is byte-aligned again and its start can be found by scanning for the 9-byte
 *   pattern 00 00 FF FF 00 00 00 FF FF (a single sync marker would turn up by chance in compressed
 *   data); the last chunk's block carries BFINAL.  No match crosses a chunk boundary.
 */
#ifndef B200_DEFLATE_H
#define B200_DEFLATE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_DEFLATE_ABI_VERSION 1

/* Return codes (0 = success). */
#define B200_OK 0
#define B200_E_OVERRUN 1  /* inflate: ran past the input  (reference: std::runtime_error("Reading bits \
                             beyond the alloted buffer size!"), inflate.hpp:82,98,107) */
#define B200_E_DATA 2     /* inflate: invalid stream (bad code lengths, bad symbol, bad block) */
#define B200_E_OUTPUT 3   /* output buffer too small (only where truncation is not the contract) */
#define B200_E_CUDA 4     /* CUDA runtime / driver error, or no sm_100 device: there is no CPU path */
#define B200_E_ARG 5      /* bad argument */
#define B200_E_NOMEM 6    /* host or device allocation failed */
#define B200_E_IO 7       /* file API: a file could not be opened, read or written */

/* Compression levels: the reference's `int compression_level` (deflate.hpp:675-680). */
#define B200_LEVEL_STORED 0  /* stored blocks only */
#define B200_LEVEL_HUFFMAN 1 /* Huffman only, no matching */
#define B200_LEVEL_FAST 2    /* greedy hash matcher      (reference "fast",   getMatches     :310) */
#define B200_LEVEL_BETTER 3  /* lazy + deeper matcher    (reference "better", getMatchesSlow :268) */

/* Flags for the compressor. */
#define B200_F_NOT_LAST 1u /* this buffer is a shard that is NOT the end of the stream: its last chunk \
                              is closed with the byte-aligning empty stored block instead of BFINAL */
#define B200_F_NO_INDEX 2u /* do not put the segment index (20 bytes per 4 KiB segment -- 320 for a full chunk -- of  \
                              empty stored blocks whose padding bits hold the bit length of every segment) in      \
                              front of chunks of two segments or more: the stream is 0.3-0.5 % smaller and         \
                              inflates one warp per chunk instead of one thread per segment */
#define B200_F_ZLIB 4u     /* wrap the stream in zlib framing (RFC 1950): 78 9C in front, the Adler-32 of the INPUT behind,  \
                              computed on the GPU (what zlib.decompress / inflate::decompressZlib read) */
#define B200_F_GZIP 8u     /* wrap it in one gzip member (RFC 1952): 10-byte header, CRC-32 of the input + ISIZE behind */
/* Flags for the inflater. */
#define B200_F_STRICT 1u   /* reject what the reference silently accepts: distance beyond the output \
                              produced so far, NLEN != ~LEN, BTYPE 3, and for zlib streams a bad header or   \
                              Adler-32 trailer (default: behave like the reference) */

#define B200_CHUNK_BYTES 65536u

typedef struct b200_ctx b200_ctx; /* per-device workspace; use one ctx from one thread at a time */

int b200_abi_version(void);
const char* b200_strerror(int code);
/* Number of kernel launches issued by this library since load (bench.py reports it as gpu_launches). */
uint64_t b200_launch_count(void);

/* ---- context ------------------------------------------------------------------------------- */
int b200_ctx_create(int device, b200_ctx** ctx);
void b200_ctx_destroy(b200_ctx* ctx);

/* Per-kernel timing for benchmarks: when enabled, every kernel launch made through `ctx` is bracketed
 * by CUDA events on its stream.  b200_ctx_profile(ctx, 1) clears earlier records and starts,
 * b200_ctx_profile(ctx, 0) stops (records are kept for reading);
 * b200_ctx_profile_read() synchronizes on the recorded events and returns the summed device time and the launch
 * count of one kernel (ids 0.. as named by b200_kernel_name(); NULL past the last id). */
int b200_ctx_profile(b200_ctx* ctx, int enable);
int b200_ctx_profile_read(b200_ctx* ctx, int kernel_id, double* total_ms, uint64_t* launches);
const char* b200_kernel_name(int kernel_id);

/* Worst-case compressed size for n input bytes (all chunks stored + framing). */
size_t b200_deflate_bound(size_t n);

/* ---- host-buffer API: what the header-only drop-ins call ------------------------------------ */

/* Replaces deflate::compress(char* data, size_t data_size, int compression_level) -> vector
 * (reference include/deflate.hpp:779) and the vector overload (:798).
 * *out is malloc'ed by the library; release with b200_free().  Never fails on valid arguments other
 * than for CUDA / memory errors (the reference compressor never throws to its caller either). */
int b200_deflate_compress(const void* in, size_t n, int level, void** out, size_t* out_n);

/* The same two calls with compressor flags (B200_F_NO_INDEX, B200_F_ZLIB, B200_F_GZIP): framed output for the wire formats
 * next to the path (the reference only has the inflate side, inflate.hpp:326-361). */
int b200_deflate_compress_ex(const void* in, size_t n, int level, unsigned flags, void** out, size_t* out_n);
int b200_deflate_compress_into_ex(const void* in, size_t n, int level, unsigned flags, void* out, size_t cap, size_t* out_n);

/* Same, into a caller buffer of `cap` bytes (B200_E_OUTPUT if cap < compressed size; cap >=
 * b200_deflate_bound(n) always suffices).  Pinned host memory is streamed without staging. */
int b200_deflate_compress_into(const void* in, size_t n, int level, void* out, size_t cap, size_t* out_n);

/* Replaces inflate::decompress(void* in, size_t in_size, void* out, size_t out_size) -> size_t
 * (reference include/inflate.hpp:338): decodes the raw RFC 1951 stream, writes at most `cap` bytes
 * and, like the reference (:345), silently truncates: *out_n = min(decoded size, cap).
 * *full_n (may be NULL) receives the full decoded size. */
int b200_inflate(const void* in, size_t n, void* out, size_t cap, size_t* out_n, size_t* full_n,
                 unsigned flags);

/* Replaces inflate::decompress(void* in, size_t in_size) -> vector (inflate.hpp:363) and the vector
 * overload (:376).  *out is malloc'ed; release with b200_free(). */
int b200_inflate_alloc(const void* in, size_t n, void** out, size_t* out_n, unsigned flags);

/* Replaces inflate::decompressZlib (inflate.hpp:326,352): skips the 2-byte zlib header (and, unlike
 * the reference whose FDICT test at :329/:355 can never fire, the 4-byte DICTID when FDICT is set);
 * the Adler-32 trailer is ignored like the reference does.  With B200_F_STRICT the header (CM = 8, check
 * bits) and the trailer are verified -- the Adler-32 is computed on the GPU over the decoded bytes -- and a
 * mismatch is B200_E_DATA (the output has been written all the same). */
int b200_inflate_zlib(const void* in, size_t n, void* out, size_t cap, size_t* out_n, size_t* full_n,
                      unsigned flags);
int b200_inflate_zlib_alloc(const void* in, size_t n, void** out, size_t* out_n, unsigned flags);

/* gzip (RFC 1952): first member of a .gz buffer; header fields FEXTRA / FNAME / FCOMMENT / FHCRC are skipped.  With
 * B200_F_STRICT the CRC-32 (computed on the GPU over the decoded bytes) and ISIZE must match, else B200_E_DATA. */
int b200_inflate_gzip(const void* in, size_t n, void* out, size_t cap, size_t* out_n, size_t* full_n,
                      unsigned flags);
int b200_inflate_gzip_alloc(const void* in, size_t n, void** out, size_t* out_n, unsigned flags);

void b200_free(void* p);

/* Views (what the header-only drop-ins use for the std::vector-returning overloads): the result is left in the library's
 * pinned host arena -- written there by the DMA engine, no page faults, no intermediate malloc -- and *view points at it
 * until b200_view_release(); the caller copies it once into its own storage (one std::vector::assign).  A successful
 * *_view call keeps the default context locked until b200_view_release() (same thread).  framing: 0 raw, 1 zlib. */
int b200_deflate_compress_view(const void* in, size_t n, int level, unsigned flags, const void** view, size_t* out_n);
int b200_inflate_view(const void* in, size_t n, unsigned flags, int framing, const void** view, size_t* out_n);
void b200_view_release(void);
/* Fault in the pages of fresh host memory [p, p + n) from several threads (one zero byte is written per page, so only for
 * memory whose contents do not matter yet): what the drop-in headers do to the std::vector they are about to fill. */
void b200_host_prefault(void* p, size_t n);

/* ---- file-path API -------------------------------------------------------------------------- */
/* Replaces deflate::compress(std::string file_path, std::string new_file, int level) (reference include/deflate.hpp:755-777):
 * the file is streamed through the GPU in slices of 64 MiB with bounded memory -- pinned host staging, slice k + 1 is
 * read while slice k is compressed and slice k - 1 is written.  The output is byte-identical to compressing the whole
 * file in memory.  flags: B200_F_NO_INDEX, B200_F_ZLIB, B200_F_GZIP (checksum accumulated on the device across slices).
 * *in_n / *out_n (may be NULL): bytes read / written.  B200_E_IO on file errors. */
int b200_deflate_compress_file(const char* in_path, const char* out_path, int level, unsigned flags, size_t* in_n,
                               size_t* out_n);
/* Replaces inflate::decompress(std::string file_path, std::string new_file) (include/inflate.hpp:390-408; the reference
 * fails on multi-block files).  Streams of this library's chunk format are decoded window by window with bounded
 * memory (the next window is read and the previous output written while the current one is on the GPU); any other raw
 * DEFLATE stream goes through the memory API as a whole. */
int b200_inflate_file(const char* in_path, const char* out_path, unsigned flags, size_t* in_n, size_t* out_n);

/* ---- device-resident API (benchmarks, pipelines, multi-GPU shards) -------------------------- */
/* All pointers prefixed d_ are device pointers on ctx's device; `stream` is a cudaStream_t passed as
 * void* (NULL = default stream).  Work is enqueued on `stream`.  If h_out_n is non-NULL the call
 * synchronizes the stream and stores the byte count there; d_out_n (may be NULL) receives it on the
 * device without a sync. */

/* Compress d_in[0..n) into d_out[0..cap).  cap must be >= b200_deflate_bound(n).
 * d_chunk_off (may be NULL): n_chunks+1 uint64 byte offsets of each chunk's compressed bytes inside
 * d_out (an optional index for chunk-parallel inflate; the stream is self-describing without it). */
int b200_deflate_compress_dev(b200_ctx* ctx, const void* d_in, size_t n, int level, unsigned flags,
                              void* d_out, size_t cap, uint64_t* d_out_n, size_t* h_out_n,
                              uint64_t* d_chunk_off, void* stream);

/* Batch compression: n_files independent inputs (input f = d_in + d_in_off[f], d_in_len[f] bytes; empty inputs
 * allowed) become n_files independent raw DEFLATE streams in one launch sequence -- what calling
 * deflate::compress (deflate.hpp:779) once per file produces, without the per-call launch latency.  Stream f is
 * d_out[d_out_off[f] .. d_out_off[f + 1]) (d_out_off has n_files + 1 entries, device memory); every stream ends
 * with its own BFINAL block.  cap must cover the sum of b200_deflate_bound(d_in_len[f]) (checked: B200_E_ARG).
 * *h_total (may be NULL) receives the total compressed size.  Synchronizes `stream` once (chunk count). */
int b200_deflate_compress_batch_dev(b200_ctx* ctx, const void* d_in, const uint64_t* d_in_off,
                                    const uint64_t* d_in_len, size_t n_files, int level, unsigned flags,
                                    void* d_out, size_t cap, uint64_t* d_out_off, size_t* h_total, void* stream);

/* Multi-GPU gather fused into the encoder.  Stage 1 runs tokenise / code construction / sizing for one
 * batch (n <= the batch size: 16384 chunks unless B200_BATCH_CHUNKS says otherwise) and leaves this shard's compressed byte count in *d_local_n (device).  The
 * caller exchanges the counts between ranks (an 8-byte all_gather) and computes *d_base, the offset of
 * this shard inside the joined stream.  Stage 2 bit-packs and writes every chunk of the shard at
 * d_out + *d_base + (offset inside the shard): d_out may be ANOTHER GPU's memory mapped over NVLink
 * (symmetric memory / cudaIpc), so the encoder's stores are the gather.  Same ctx, same stream, stage 2
 * directly after stage 1 of the same input. */
int b200_deflate_compress_stage1_dev(b200_ctx* ctx, const void* d_in, size_t n, int level, unsigned flags,
                                     uint64_t* d_local_n, void* stream);
int b200_deflate_compress_stage2_dev(b200_ctx* ctx, const void* d_in, size_t n, void* d_out,
                                     const uint64_t* d_base, void* stream);

/* Inflate one raw stream.  Streams produced by this library (or any stream whose blocks are joined
 * by byte-aligning pairs of empty stored blocks and whose chunks do not reference earlier chunks) are
 * decoded chunk-parallel: chunks that carry the segment index by one thread per 4 KiB segment plus a copy
 * pass (csrc/inflate_tp.cuh), chunks without it by one warp each; anything else is decoded by a single
 * warp.  Scratch: 128 KiB per chunk, at most 2 x 32768 chunks.  Writes at most cap bytes; *h_out_n /
 * *d_out_n = min(decoded, cap); status (B200_*) is the return value when h_out_n is given, else it is
 * left in d_status (int32, may be NULL).  Synchronizes `stream` internally. */
int b200_inflate_dev(b200_ctx* ctx, const void* d_in, size_t n, void* d_out, size_t cap,
                     uint64_t* d_out_n, size_t* h_out_n, size_t* h_full_n, int32_t* d_status,
                     unsigned flags, void* stream);

/* Multi-GPU inflate of one stream of this library's format: every GPU takes a WINDOW d_in[0..n) of the joined
 * stream -- its byte range plus slack for the chunk that straddles the range's end -- and decodes the chunks that
 * START at window offsets in [lo, hi) (hi >= n: up to the end).  first_is_start: a chunk begins at byte 0 of the
 * window (true for the window at offset 0 of the stream; lo is then ignored); otherwise chunk starts are the bytes
 * behind the chunk separators found in the window (the window must begin at least 9 bytes below lo).  ends_stream: the window contains the end of the stream (its last chunk carries BFINAL).
 * Chunk k of the window goes to d_out + k * 64 KiB.  *h_out_n = decoded bytes, *h_nchunks = chunks decoded,
 * *h_next_start = window offset of the first chunk NOT decoded (n if none).  B200_E_DATA if the window is not a run
 * of this library's chunks.  Synchronizes `stream`. */
int b200_inflate_shard_dev(b200_ctx* ctx, const void* d_in, size_t n, size_t lo, size_t hi, int first_is_start,
                           int ends_stream, void* d_out, size_t cap, size_t* h_out_n, size_t* h_nchunks,
                           size_t* h_next_start, unsigned flags, void* stream);

/* Batch inflate: n_streams independent raw streams, one warp each (BASELINE config 4).
 * Stream i is d_in + d_in_off[i], d_in_len[i] bytes; output goes to d_out + d_out_off[i], at most
 * d_out_cap[i] bytes (truncating); d_out_len[i] = full decoded size; d_status[i] = B200_* code. */
int b200_inflate_batch_dev(b200_ctx* ctx, const void* d_in, const uint64_t* d_in_off,
                           const uint64_t* d_in_len, void* d_out, const uint64_t* d_out_off,
                           const uint64_t* d_out_cap, uint64_t* d_out_len, int32_t* d_status,
                           size_t n_streams, unsigned flags, void* stream);

/* Adler-32 (RFC 1950) of d_data[0..n) computed on the device; the result goes to *h_out (synchronizes the
 * stream) and / or *d_out (device, no sync).  What B200_F_STRICT uses to check a zlib stream's trailer. */
int b200_adler32_dev(b200_ctx* ctx, const void* d_data, size_t n, uint32_t* h_out, uint32_t* d_out, void* stream);

/* CRC-32 (RFC 1952 / zlib crc32) of d_data[0..n) computed on the device: one CTA per 64 KiB block, blocks combined with the
 * GF(2) identity crc(A || B) = crc(A) * x^(8|B|) + crc(B).  Same result conventions as b200_adler32_dev. */
int b200_crc32_dev(b200_ctx* ctx, const void* d_data, size_t n, uint32_t* h_out, uint32_t* d_out, void* stream);

/* n_words (<= 32) 64-bit words from device memory to PINNED host memory (cudaHostAlloc / cudaHostRegister, mapped into the
 * device's address space as all pinned memory is under unified addressing), written by a kernel on `stream` -- not by a
 * copy: an 8-byte cudaMemcpyAsync is queued on the copy engine behind whatever bulk transfer is on the wire, and the
 * stream it is on waits with it.  The multi-GPU gather (deflate.hpp_b200/shard.py) passes the sizes of a round to the host
 * this way.  The words are visible to the host once work enqueued behind the call on `stream` has been waited for. */
int b200_publish_dev(b200_ctx* ctx, void* h_pinned_dst, const void* d_src, size_t n_words, void* stream);

/* Synthetic corpus of BASELINE config 3/5 (definition: oracle/corpus_oracle.c):
 * chunks first_chunk .. first_chunk+n_chunks-1 of 64 KiB each, written to d_out. */
int b200_corpus_generate_dev(void* d_out, uint64_t seed, uint64_t first_chunk, uint64_t n_chunks,
                             void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200_DEFLATE_H */
