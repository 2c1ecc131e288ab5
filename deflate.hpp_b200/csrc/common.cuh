// common.cuh -- shared constants, RFC 1951 symbol arithmetic and small device helpers.
//
// Replaces the reference's linear-scan RangeLookup tables (include/common.hpp:408-440, 508-575)
// with closed-form arithmetic (count-leading-zeros), and its token word (deflate.hpp:11-13) with a
// compact one-word-per-TOKEN format (the reference stores one word per input BYTE).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

constexpr uint32_t CHUNK = 65536;          // independent unit: one DEFLATE block, one CTA
constexpr uint32_t NSEG = 16;              // segments per chunk: one warp each (parse and encode)
constexpr uint32_t SEG = CHUNK / NSEG;     // 4096 bytes parsed by one warp; tokens never cross a segment
constexpr uint32_t NLIT = 288;             // literal/length alphabet slots (286 used)
constexpr uint32_t NDIST = 32;             // distance alphabet slots (30 used)
constexpr uint32_t NSYM = NLIT + NDIST;    // 320: [0,288) lit/len, [288,320) dist
constexpr uint32_t HDR_WORDS = 160;        // >= 4498 bits worst-case dynamic header
constexpr uint32_t MIN_MATCH = 3;
constexpr uint32_t MAX_MATCH = 258;
constexpr uint32_t MAX_DIST = 32768;
// Chunk separator: TWO empty non-final stored blocks.  Byte-aligned it reads 00 | 00 00 FF FF | 00 | 00 00 FF FF
// (after a Huffman block the first header's 3 bits share the block's last byte).  The 9-byte tail
// "00 00 FF FF 00 00 00 FF FF" is what the chunk-parallel inflater scans for: a single sync marker
// (4 bytes) shows up by chance about once per 4 GiB of compressed data, the doubled one never does.
constexpr uint32_t SYNC_BYTES_ALIGNED = 10;   // bytes the separator takes when it starts byte-aligned
constexpr uint32_t SYNC_PATTERN_BYTES = 9;

// Segment index (parallel inflate inside a chunk).  A Huffman-coded chunk of S segments (S = 16 for a full 64 KiB
// chunk, 8..15 for the short last chunk of a stream; shorter ones have none) is preceded by 4 S empty non-final stored blocks, each starting
// byte-aligned: byte 0 = 1pppp000 (BFINAL 0, BTYPE 00, then five "ignored up to the byte boundary" bits, RFC 1951 3.2.4),
// then 00 00 FF FF.  The four p bits of the groups spell S little-endian u16 words: word 0 = INDEX_MAGIC | (S - 1) << 10,
// word s (1..S-1) = the number of bits segment s-1's symbols take.  Every inflater skips these blocks; this
// one reads them and starts one thread per 4 KiB segment.  Bit 7 of byte 0 is always set so that a group can
// never look like a chunk separator (below) -- the separator's stored blocks have all-zero padding.
constexpr uint32_t INDEX_GROUPS = 64;                       // of a full chunk
constexpr uint32_t INDEX_BYTES = INDEX_GROUPS * 5;          // 320
constexpr uint32_t INDEX_BYTES_PER_SEG = 20;                // four groups of 5 bytes per 16-bit word
constexpr uint32_t INDEX_MIN_SEGS = 8;                      // chunks shorter than this many segments get no index (huffman.cuh)
constexpr uint32_t INDEX_MAGIC = 0x2B5;                     // 10 bits

// Token: literal -> byte value (dist field 0); match -> length in bits [0,9), distance in [16,32).
__host__ __device__ __forceinline__ uint32_t tok_match(uint32_t len, uint32_t dist) { return len | (dist << 16); }
__host__ __device__ __forceinline__ uint32_t tok_dist(uint32_t t) { return t >> 16; }
__host__ __device__ __forceinline__ uint32_t tok_len(uint32_t t) { return t & 0x1FFu; }

// Batch compression (many independent inputs in one launch sequence): chunk c of the batch is bytes
// [off, off + clen) of the input buffer; `last` bit 0 = it closes its stream (BFINAL, no separator after it), bit 1 = its
// input is a small one (< SMALL_INPUT_BYTES: the better level searches deeper there).
// Kernels take a `const ChunkSrc*` that is NULL for one contiguous input (chunk c = bytes [c * CHUNK, ...)).
struct ChunkSrc { uint64_t off; uint32_t clen; uint32_t last; };
constexpr uint64_t SMALL_INPUT_BYTES = 1ull << 20;
// ntok[] entry of a segment that consists of literals only: bit 31 set, no tokens in tok[] -- token i is input byte i
constexpr uint32_t NTOK_LITERALS = 0x80000000u;

// Block descriptor written by the Huffman kernel, read by the encoder.
struct BlockDesc {
    uint32_t btype;          // 0 stored, 1 fixed, 2 dynamic
    uint32_t hdr_bits;       // bits in hdr[] (3-bit block header + dynamic tables)
    uint32_t total_bits;     // whole Huffman block incl. header and EOB (btype 1/2)
    uint32_t nbytes;         // bytes this chunk occupies in the stream (incl. sync marker if any)
    uint32_t clen;           // uncompressed bytes in this chunk
    uint32_t last;           // 1 = carries BFINAL, no sync marker
    uint32_t eob;            // (unused)
    uint32_t index_bytes;    // 0, or 20 bytes per segment when the chunk is preceded by the segment index
    uint32_t seg_bitoff[NSEG];  // bit offset of each segment's first token, relative to block start
    // a chunk split into two blocks (huffman.cuh): segments [0, split_seg) use codes[0] / hdr[0], the rest codes[1] / hdr[1]
    uint32_t split_seg;      // 0 = one block
    uint32_t btype2;         // second block: 1 fixed, 2 dynamic
    uint32_t hdr_bits2;      // bits of the second block's header
    uint32_t block2_bit;     // bit offset of the second block's header, relative to the first block's start
};

// ---- RFC 1951 3.2.5 symbol arithmetic ------------------------------------------------------

// length 3..258 -> symbol index 0..28 (symbol = 257 + index), extra-bit count, extra value
__device__ __forceinline__ void len_symbol(uint32_t len, uint32_t& idx, uint32_t& nextra, uint32_t& extra) {
    uint32_t l = len - 3;
    if (l < 8) { idx = l; nextra = 0; extra = 0; return; }
    if (len == 258) { idx = 28; nextra = 0; extra = 0; return; }
    uint32_t k = 31 - __clz(l);               // >= 3
    nextra = k - 2;
    idx = 4 * (k - 1) + ((l >> nextra) & 3);
    extra = l & ((1u << nextra) - 1);
}

// distance 1..32768 -> symbol 0..29, extra-bit count, extra value
__device__ __forceinline__ void dist_symbol(uint32_t dist, uint32_t& sym, uint32_t& nextra, uint32_t& extra) {
    uint32_t d = dist - 1;
    if (d < 4) { sym = d; nextra = 0; extra = 0; return; }
    uint32_t k = 31 - __clz(d);               // >= 2
    nextra = k - 1;
    sym = 2 * k + ((d >> nextra) & 1);
    extra = d & ((1u << nextra) - 1);
}

__host__ __device__ __forceinline__ uint32_t len_extra_bits(uint32_t idx) {   // idx = symbol - 257
    return (idx < 8 || idx == 28) ? 0 : (idx - 4) >> 2;
}
__host__ __device__ __forceinline__ uint32_t dist_extra_bits(uint32_t sym) {
    return sym < 4 ? 0 : (sym - 2) >> 1;
}
__host__ __device__ __forceinline__ uint32_t fixed_lit_len(uint32_t sym) {   // common.hpp:442-482
    return sym < 144 ? 8 : sym < 256 ? 9 : sym < 280 ? 7 : 8;
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ uint32_t bitrev(uint32_t code, uint32_t len) { return __brev(code) >> (32 - len); }

// 4 bytes at an arbitrary byte offset of a shared-memory buffer (two aligned words + funnel shift).
__device__ __forceinline__ uint32_t ld4_unaligned(const uint8_t* base, uint32_t off) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(base) + (off >> 2);
    return __funnelshift_r(w[0], w[1], (off & 3) * 8);
}

}  // namespace b200
