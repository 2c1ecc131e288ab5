// inflate_tp.cuh -- K6 (fast path): two-pass INFLATE for many independent units (the chunks of one stream,
// or the streams of a batch).
//
// The one-warp-per-unit decoder in inflate.cuh runs the bit-serial symbol decode redundantly on all 32
// lanes: ncu shows it issue-bound at ~50 warp-instructions per symbol (profiles/, r01b rows).  Here the two
// halves of INFLATE are separated by what they parallelise over:
//
//   pass A  Huffman decode only (reference: decompressHuffmanBlock / decodeTree, include/inflate.hpp:136-275),
//           ONE THREAD per piece of the stream.  A thread reads its input through a 4-word cp.async ring in
//           shared memory, looks symbols up in u16 direct tables in shared memory, writes literals straight to
//           the output as merged aligned 32-bit words, and appends every back-reference (and stored block) to
//           an op list in HBM as one u64 {pos, len, dist}.  A warp instruction decodes up to 32 symbols of 32
//           different pieces instead of one symbol 32 times.
//             inflate_classify_kernel + inflate_segments_kernel  (chunks of one stream)  the piece is a 4 KiB
//               SEGMENT of a chunk: this library's compressor writes the bit length of every segment into the
//               stream (segment index, common.cuh), so 16 threads share one chunk and its tables.  Chunks
//               without index go to the one-warp decoder (or are finished at once if they are stored blocks).
//             inflate_symbols_kernel  (generic: the piece is a whole unit, each thread owns interleaved private
//               tables)  batch inflate with B200_BATCH_TP=1; on 1-64 KiB zlib streams one warp per stream is
//               faster, so this is opt-in.
//   pass B  inflate_copy_kernel, ONE WARP per unit.  Applies the unit's ops in order (reference: the
//           back-reference loop inflate.hpp:262-272 and the stored-block copy :294-303).  32 ops are loaded
//           per step; short ops whose source is final before the step starts are independent of one another
//           and are copied one per lane, the rest go through the cooperative (all lanes, period-aware) copy
//           in order.
//
// Pass A may leave garbage in bytes that pass B owns (it stores whole words); pass B runs after pass A
// (stream order) and only ever reads bytes below its current position, so every byte a copy reads is final.
// Units the fast path cannot take (no index, index that does not chain, op list full, > 4 GiB) are marked
// ST_FALLBACK and decoded by inflate_fallback_kernel = the one-warp decoder of inflate.cuh; results are
// identical either way (every GPU test stream goes through both).
#pragma once
#include "inflate.cuh"

namespace b200 {

constexpr uint32_t TP_LIT_BITS = 9;
constexpr uint32_t TP_DST_BITS = 8;
constexpr uint32_t SG_LIT_BITS = 10;                                       // segment kernel: one table per chunk, shared by 16 threads
constexpr uint32_t TP_THREADS = 128;                                       // units per CTA in pass A
constexpr uint32_t TP_ENTRIES = (1u << TP_LIT_BITS) + (1u << TP_DST_BITS); // u16 entries per thread
constexpr uint32_t TP_LUT_WORDS = 64;                                      // [0,32) length base|extra, [32,64) distance
constexpr uint32_t TP_SMEM_BYTES = TP_ENTRIES * 2 * TP_THREADS + TP_LUT_WORDS * 4;
constexpr uint32_t TP_LIT_RUN = 3;             // extra literals a thread may take per trip of the hot loop
constexpr int ST_FALLBACK = 3;                 // internal: unit must be redone by the one-warp decoder
constexpr uint32_t OPS_PER_CHUNK = CHUNK / 4;  // op-list capacity of one chunk (u64 each)

// op: [31:0] destination position inside the unit, [47:32] length, [63:48] distance.
// distance 0 + length > 0 = stored block; the NEXT slot holds the source byte offset inside the unit's input
// (pass A never lets such a pair straddle a 32-op step).  length 0 = no-op padding.
__device__ __forceinline__ uint64_t tp_op(uint32_t pos, uint32_t len, uint32_t dist) {
    return (uint64_t)pos | ((uint64_t)len << 32) | ((uint64_t)dist << 48);
}

// table entries (u16): [3:0] code length (0 = not in the table: longer code or no code), [15:4] payload
//   lit/len payload: < 256 literal byte; 0x100 end of block; 0x200 | idx length symbol 257 + idx; 0x300 invalid (286/287)
//   distance payload: symbol 0..31 (30, 31 invalid)
__device__ __forceinline__ uint32_t tp_lit_entry(uint32_t sym, uint32_t len) {
    uint32_t p = sym;
    if (sym > 256) p = sym > 285 ? 0x300u : (0x200u | (sym - 257));
    return (p << 4) | len;
}

struct TpTables {   // per-thread, in local memory (L1-resident); only touched for headers and long codes
    uint8_t lens[NSYM];
    uint16_t sorted[NSYM];
    uint16_t first[2][16], count[2][16], offs[2][16];
};

// Builds one direct-lookup table (single thread).  which: 0 lit/len -> tab[0 .. 2^tbits), 1 distance or
// precode -> tab[base ..).  Entries of thread t live at tab[index * NT] (tab already points at column t).
__device__ bool tp_build(uint16_t* tab, uint32_t NT, TpTables& T, const uint8_t* lens, uint32_t n, uint32_t which,
                         uint32_t tbits, bool lit_alphabet) {
    uint16_t* count = T.count[which];
    uint16_t* first = T.first[which];
    uint16_t* offs = T.offs[which];
    uint16_t* sorted = T.sorted + (which ? NLIT : 0);
    uint16_t next[16];
    for (uint32_t l = 0; l < 16; l++) { count[l] = 0; next[l] = 0; }
    for (uint32_t i = 0; i < n; i++) count[lens[i]]++;
    count[0] = 0;
    uint32_t code = 0, off = 0;
    int left = 1;
    for (uint32_t l = 1; l <= 15; l++) {
        code <<= 1;
        first[l] = (uint16_t)code;
        offs[l] = (uint16_t)off;
        code += count[l];
        off += count[l];
        left = (left << 1) - (int)count[l];
        if (left < 0) return false;
    }
    for (uint32_t i = 0; i < (1u << tbits); i++) tab[i * NT] = 0;
    for (uint32_t s = 0; s < n; s++) {
        const uint32_t l = lens[s];
        if (!l) continue;
        const uint32_t r = next[l]++;
        sorted[offs[l] + r] = (uint16_t)s;
        if (l <= tbits) {
            const uint32_t rev = bitrev(first[l] + r, l);
            const uint16_t e = (uint16_t)(lit_alphabet ? tp_lit_entry(s, l) : ((s << 4) | l));
            for (uint32_t k = rev; k < (1u << tbits); k += (1u << l)) tab[k * NT] = e;
        }
    }
    return true;
}

// canonical search for a code longer than the direct table: returns symbol | length << 16, or -1 (no code)
__device__ __noinline__ int tp_slow_symbol(const TpTables& T, uint32_t which, uint32_t bits32, uint32_t lmin) {
    const uint32_t rb = __brev(bits32);
    for (uint32_t l = lmin; l <= 15; l++) {
        const uint32_t code = rb >> (32 - l);
        const uint32_t f = T.first[which][l];
        if (code >= f && code - f < T.count[which][l])
            return (int)(T.sorted[(which ? NLIT : 0) + T.offs[which][l] + (code - f)] | (l << 16));
    }
    return -1;
}

__device__ __forceinline__ uint32_t lds_u16(uint32_t saddr) {
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(saddr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t saddr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}

// ---- per-thread bit reader: a small ring in shared memory, filled by cp.async, between HBM and the 64-bit buffer --
// ~800 streams per SM are read 4 bytes at a time: the L1 cannot hold a line per stream, 4 of 10 words come from
// DRAM.  A register queue between the load and its use does not help enough -- shifting the queue touches the
// register a load is still writing one refill after it was issued, and refills come in bursts (literal runs).
// cp.async has no destination register: word i + TB_RING is requested when word i is consumed, sits in the
// thread's ring slots in shared memory ([slot][thread]: conflict-free) and is waited for TB_RING - 1 refills later.
constexpr uint32_t TB_RING = 4;
struct TBits {
    const uint32_t* wp;   // 4-byte aligned address at or below the stream start
    uint32_t nw;          // words from wp that cover the stream
    uint32_t wi;          // words merged into bb so far == index of the word held in w0
    uint32_t w0;          // next word to merge
    uint32_t ring_sa;     // shared-memory address of this thread's slot 0
    uint32_t ring_stride; // bytes between slots (4 x threads sharing the ring array)
    uint32_t skip;        // bytes between wp and the stream start
    uint64_t bb;
    uint32_t bc;
};
// request word i into its ring slot (zero-filled past the end of the stream); one cp.async group per word
__device__ __forceinline__ void tb_request(const TBits& r, uint32_t i) {
    const uint32_t dst = r.ring_sa + (i % TB_RING) * r.ring_stride;
    const bool in_range = i < r.nw;
    const uint32_t* src = r.wp + (in_range ? i : 0u);
    const uint32_t nbytes = in_range ? 4u : 0u;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\ncp.async.commit_group;" :: "r"(dst), "l"(src), "r"(nbytes) : "memory");
}
__device__ __forceinline__ uint32_t tb_slot(const TBits& r, uint32_t i) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(r.ring_sa + (i % TB_RING) * r.ring_stride) : "memory");
    return v;
}
// merge w0 into the bit buffer (bc <= 32 on entry), request the word TB_RING ahead, fetch the next w0
__device__ __forceinline__ void tb_take(TBits& r) {
    r.bb |= (uint64_t)r.w0 << r.bc;
    r.bc += 32;
    tb_request(r, r.wi + TB_RING);                 // into the slot of word wi, which w0 has left
    asm volatile("cp.async.wait_group %0;" :: "n"(TB_RING - 1) : "memory");     // word wi + 1 has landed
    r.wi++;
    r.w0 = tb_slot(r, r.wi);
}
__device__ __forceinline__ void tb_seek(TBits& r, uint64_t stream_byte) {
    const uint64_t b = r.skip + stream_byte;
    const uint32_t w = (uint32_t)(b >> 2);
    const uint32_t sh = (uint32_t)(b & 3) * 8;
    asm volatile("cp.async.wait_all;" ::: "memory");
    #pragma unroll
    for (uint32_t k = 0; k < TB_RING; k++) tb_request(r, w + k);
    asm volatile("cp.async.wait_all;" ::: "memory");
    r.wi = w;
    r.w0 = tb_slot(r, w);
    r.bb = 0; r.bc = 0;
    tb_take(r);
    r.bb >>= sh; r.bc -= sh;
}
__device__ __forceinline__ void tb_refill(TBits& r) {      // afterwards bc >= 33
    if (r.bc < 33) tb_take(r);
}
__device__ __forceinline__ uint32_t tb_peek(const TBits& r, uint32_t n) { return (uint32_t)r.bb & ((1u << n) - 1u); }
__device__ __forceinline__ void tb_drop(TBits& r, uint32_t n) { r.bb >>= n; r.bc -= n; }
__device__ __forceinline__ uint32_t tb_get(TBits& r, uint32_t n) { const uint32_t v = tb_peek(r, n); tb_drop(r, n); return v; }
__device__ __forceinline__ uint64_t tb_bitpos(const TBits& r) {
    return (uint64_t)r.wi * 32 - r.bc - (uint64_t)r.skip * 8;
}

// ---- literal output: bytes are merged into aligned 32-bit words ---------------------------------------
struct TOut {
    uint8_t* ob;       // 4-byte aligned address at or below the unit's first output byte
    uint32_t al;       // bytes between ob and the unit's first output byte
    uint32_t v;        // virtual position: al + bytes produced
    uint32_t vlo;      // first virtual position this thread may write (al, or the start of its segment)
    uint32_t vlim;     // al + cap: first virtual position that must not be written
    uint32_t wlo, wspan;   // words [wlo, wlo + wspan) lie entirely inside [vlo, vlim): whole-word stores allowed
    uint32_t lw;       // bytes of the current word gathered so far
};
__device__ __forceinline__ void to_window(TOut& o) {
    o.wlo = (o.vlo + 3) & ~3u;
    const uint32_t whi = o.vlim & ~3u;
    o.wspan = whi > o.wlo ? whi - o.wlo : 0;
}
// the careful way to store (part of) word [ws, ws+4): only bytes inside [vlo, vlim) and below vend
__device__ __noinline__ void to_store_bytes(uint8_t* ob, uint32_t vlo, uint32_t vlim, uint32_t ws, uint32_t w, uint32_t vend) {
    #pragma unroll
    for (uint32_t k = 0; k < 4; k++) {
        const uint32_t p = ws + k;
        if (p >= vlo && p < vlim && p < vend) ob[p] = (uint8_t)(w >> (8 * k));
    }
}
// store word [ws, ws+4); bytes at or after vend are not part of this thread's output and may only be
// written when the whole word lies inside the thread's writable range
__device__ __forceinline__ void to_store(const TOut& o, uint32_t ws, uint32_t w, uint32_t vend) {
    if (ws - o.wlo < o.wspan && vend >= ws + 4) *reinterpret_cast<uint32_t*>(o.ob + ws) = w;
    else to_store_bytes(o.ob, o.vlo, o.vlim, ws, w, vend);
}
__device__ __forceinline__ void to_literal(TOut& o, uint32_t byte) {
    o.lw |= byte << ((o.v & 3) * 8);
    o.v++;
    if ((o.v & 3) == 0) {
        if (o.v - 4 - o.wlo < o.wspan) *reinterpret_cast<uint32_t*>(o.ob + (o.v - 4)) = o.lw;
        else to_store_bytes(o.ob, o.vlo, o.vlim, o.v - 4, o.lw, o.v);
        o.lw = 0;
    }
}
// the next n bytes belong to pass B (a copy).  A pending partial word is stored now; its upper bytes are
// garbage that pass B overwrites, which is only allowed when they all fall inside those n bytes.
__device__ __forceinline__ void to_skip(TOut& o, uint32_t n) {
    const uint32_t ws = o.v & ~3u;
    if (o.v & 3) to_store(o, ws, o.lw, n >= 4 - (o.v & 3) ? ws + 4 : o.v);
    const uint32_t nv = o.v + n;
    if ((nv & ~3u) != ws) o.lw = 0;
    o.v = nv;
}
// same for a back-reference: n >= 3, so the garbage always falls inside it and the word is left behind
__device__ __forceinline__ void to_skip_match(TOut& o, uint32_t n) {
    if (o.v & 3) {
        const uint32_t ws = o.v & ~3u;
        if (ws - o.wlo < o.wspan) *reinterpret_cast<uint32_t*>(o.ob + ws) = o.lw;
        else to_store_bytes(o.ob, o.vlo, o.vlim, ws, o.lw, ws + 4);
    }
    o.lw = 0;
    o.v += n;
}
__device__ __forceinline__ void to_flush(const TOut& o) {
    if (o.v & 3) to_store_bytes(o.ob, o.vlo, o.vlim, o.v & ~3u, o.lw, o.v);
}

struct TpUnit {
    const uint8_t* in; uint64_t in_len;
    uint8_t* out; uint64_t cap; uint64_t max_out;
    uint64_t* ops; uint32_t ops_cap;
    bool stop_at_sync;
};
// nops: op count of the unit's single list, or TP_SEGMENTED | number of per-segment lists (counts in segnops[])
struct TpResult { uint64_t in_end; uint64_t out_len; int32_t status; uint32_t end_flags; uint32_t nops; uint32_t pad; };
constexpr uint32_t TP_SEGMENTED = 0x80000000u;
constexpr uint32_t END_SEG = 16;                 // internal: stopped at the end of a segment
constexpr uint32_t TS_BLOCK = 0, TS_SYM = 1, TS_DONE = 2;

// what a decoding thread needs to know about its unit and its tables
struct TpCtx {
    uint16_t* lit; uint16_t* dst; uint32_t NT;      // tables: entry i at [i * NT]
    uint32_t lit_sa, dst_sa, ntb, lut_sa;           // the same as shared-memory addresses (ntb = bytes between entries)
    uint32_t lit_bits;                              // index width of the literal/length table
    TpTables* T;
    uint64_t in_len, in_bits;
    uint64_t* ops;
    uint32_t ops_cap, cap, max_out, seg_stop;       // seg_stop: stop once exactly this many bytes exist (segments)
    uint32_t stop_at;                               // min(seg_stop, max_out + 1): first byte count the hot loop looks at
    bool stop_at_sync, strict, allow_huffman;
};
struct TpState {
    TBits br; TOut o;
    uint32_t nops, state, empty_run, end_flags, bfinal;
    int st;
};
#define TP_FAIL(code) do { s.st = (code); s.state = TS_DONE; } while (0)

__device__ __forceinline__ void tp_ctx_tables(TpCtx& c, uint16_t* lit, uint16_t* dst, uint32_t NT, const uint32_t* s_lut, TpTables* T,
                                              uint32_t lit_bits) {
    c.lit = lit; c.dst = dst; c.NT = NT; c.T = T; c.lit_bits = lit_bits;
    c.lit_sa = (uint32_t)__cvta_generic_to_shared(lit);
    c.dst_sa = (uint32_t)__cvta_generic_to_shared(dst);
    c.lut_sa = (uint32_t)__cvta_generic_to_shared(s_lut);
    c.ntb = NT * 2;
}

// the block is over (end-of-block symbol, or the hot loop ran into one of its limits)
__device__ __forceinline__ void tp_end_block(const TpCtx& c, TpState& s, uint32_t wi_lim) {
    if (s.br.wi > wi_lim || tb_bitpos(s.br) > c.in_bits) TP_FAIL(ST_OVERRUN);
    else if (s.o.v - s.o.al > c.max_out) { s.end_flags |= END_TOO_BIG; TP_FAIL(ST_DATA); }
    else if (c.seg_stop != 0xFFFFFFFFu) TP_FAIL(ST_FALLBACK);      // end of block inside a non-final segment
    else if (s.bfinal) { s.end_flags |= END_FINAL; s.state = TS_DONE; }
    else s.state = TS_BLOCK;
}

// One symbol of the current Huffman block.  Mirrors the symbol loop of inflate_warp() in inflate.cuh decision
// for decision (error classes, reference quirks, truncation at cap); only the data movement differs.
// Written for a warp whose lanes decode different units: the refill is branch-free, the three outcomes
// (literal / back-reference / rare) are one if-else chain that reconverges before the next trip.
template <bool SEGM>
__device__ __forceinline__ void tp_step_symbol(const TpCtx& c, TpState& s) {
    // SEGM: tables are the chunk's SgChunk (entries contiguous, distance table right behind the literal table)
    const uint32_t ntb = SEGM ? 2u : c.ntb;
    const uint32_t dst_sa = SEGM ? c.lit_sa + (2u << SG_LIT_BITS) : c.dst_sa;
    constexpr uint32_t LB = SEGM ? SG_LIT_BITS : TP_LIT_BITS;
    TBits& br = s.br;
    TOut& o = s.o;
    const uint32_t produced = o.v - o.al;
    const uint32_t wi_lim = br.nw + 4;
    if (produced >= c.stop_at || br.wi > wi_lim) {
        // rare: end of the segment, output limit, or far past the end of the input
        if (br.wi <= wi_lim && produced <= c.max_out && produced >= c.seg_stop) {
            if (produced == c.seg_stop) { s.end_flags |= END_SEG; s.state = TS_DONE; }
            else TP_FAIL(ST_FALLBACK);           // a token crossed the segment boundary: the index is not ours
        } else {
            tp_end_block(c, s, wi_lim);
        }
    } else {
        // refill: a plain (reconverging) branch; the word requested here is waited for TB_RING - 1 refills later
        if (br.bc < 33) tb_take(br);
        uint32_t e = lds_u16(c.lit_sa + tb_peek(br, LB) * ntb);
        bool ok = true;
        if ((e & 15u) == 0) {
            const int r = tp_slow_symbol(*c.T, 0, (uint32_t)br.bb, LB + 1);
            if (r < 0) { TP_FAIL(ST_OVERRUN); ok = false; }
            else e = tp_lit_entry((uint32_t)r & 0xFFFFu, (uint32_t)r >> 16);
        }
        uint32_t l = e & 15u, p = e >> 4;
        bool is_match = ok && p >= 0x200u && !(p & 0x100u);
        if (ok && p < 256) {
            tb_drop(br, l);
            to_literal(o, p);
            if (SEGM) {
                // literals come in runs: take up to TP_LIT_RUN more in the same trip while the table has them, and
                // if the run ends at a back-reference, take that too -- every lane then does "literals, match" per
                // trip, the two halves of the trip are executed once for the whole warp
                #pragma unroll
                for (uint32_t k = 0; k < TP_LIT_RUN; k++) {
                    if (k) tb_refill(br);        // after the first two literals fewer than LB bits may be left
                    const uint32_t e2 = lds_u16(c.lit_sa + tb_peek(br, LB) * ntb);
                    const uint32_t l2 = e2 & 15u, p2 = e2 >> 4;
                    if (l2 == 0 || o.v - o.al >= c.stop_at) break;
                    if (p2 < 256) { tb_drop(br, l2); to_literal(o, p2); continue; }
                    if (p2 >= 0x200u && !(p2 & 0x100u)) { l = l2; p = p2; is_match = true; }
                    break;
                }
            }
        }
        if (is_match) {
            const uint32_t pos = o.v - o.al;
            tb_drop(br, l);
            const uint32_t lt = lds_u32(c.lut_sa + (p & 31u) * 4);
            const uint32_t length = (lt & 0xFFFFu) + tb_get(br, lt >> 16);
            tb_refill(br);
            const uint32_t de = lds_u16(dst_sa + tb_peek(br, TP_DST_BITS) * ntb);
            uint32_t dl = de & 15u, dsym = de >> 4;
            if (dl == 0) {
                const int r = tp_slow_symbol(*c.T, 1, (uint32_t)br.bb, TP_DST_BITS + 1);
                dsym = r < 0 ? 99u : ((uint32_t)r & 0xFFFFu);
                dl = r < 0 ? 0u : ((uint32_t)r >> 16);
            }
            if (dsym > 29) {
                TP_FAIL(tb_bitpos(br) + 16 > c.in_bits ? ST_OVERRUN : ST_DATA);
            } else {
                tb_drop(br, dl);
                const uint32_t dt = lds_u32(c.lut_sa + (32 + dsym) * 4);
                const uint32_t dist = (dt & 0xFFFFu) + tb_get(br, dt >> 16);
                if (dist > pos) {
                    if (c.stop_at_sync) { s.end_flags |= END_NEEDS_HISTORY; TP_FAIL(ST_DATA); }
                    else if (c.strict) TP_FAIL(ST_DATA);
                    // else reference: copies nothing (inflate.hpp:268-270)
                } else if (pos < c.cap && s.nops >= c.ops_cap) {
                    TP_FAIL(ST_FALLBACK);
                } else {
                    if (pos < c.cap) c.ops[s.nops++] = tp_op(pos, length, dist);
                    to_skip_match(o, length);
                }
            }
        } else if (ok && p >= 256) {
            if (p == 0x100u) { tb_drop(br, l); tp_end_block(c, s, wi_lim); }    // end of block
            else TP_FAIL(ST_DATA);                                              // 286 / 287
        }
    }
}

// One block header: stored blocks become ops, Huffman blocks get their tables built (single thread).
__device__ __noinline__ TpState tp_step_block(const TpCtx c, TpState s) {
    TBits& br = s.br;
    TOut& o = s.o;
    TpTables& T = *c.T;
    if (tb_bitpos(br) + 3 > c.in_bits) { TP_FAIL(ST_OVERRUN); return s; }
    tb_refill(br);
    const uint32_t hdr = tb_get(br, 3);
    s.bfinal = hdr & 1;
    const uint32_t btype = hdr >> 1;
    if (btype == 0) {
        const uint32_t padbits = tb_peek(br, br.bc & 7);
        tb_drop(br, br.bc & 7);
        tb_refill(br);
        const uint32_t len = tb_get(br, 16);
        const uint32_t nlen = tb_get(br, 16);
        if (c.strict && (len ^ nlen) != 0xFFFFu) { TP_FAIL(ST_DATA); return s; }
        const uint64_t bpos = tb_bitpos(br) >> 3;
        if (bpos + len > c.in_len) { TP_FAIL(ST_OVERRUN); return s; }
        if (len) {
            const uint32_t pos = o.v - o.al;
            if (pos < c.cap) {
                if ((s.nops & 31) == 31 && s.nops < c.ops_cap) c.ops[s.nops++] = 0;      // keep the pair inside one step
                if (s.nops + 2 > c.ops_cap) { TP_FAIL(ST_FALLBACK); return s; }
                c.ops[s.nops++] = tp_op(pos, len, 0);
                c.ops[s.nops++] = bpos;
            }
            to_skip(o, len);
        }
        tb_seek(br, bpos + len);
        if (len == 0 && !s.bfinal && padbits == 0) {
            // chunk separator = two empty stored blocks with all-zero padding (the segment index's have not)
            if (++s.empty_run == 2 && c.stop_at_sync) { s.end_flags |= END_SYNC; s.state = TS_DONE; return s; }
        } else {
            s.empty_run = 0;
        }
    } else if (btype == 3) {
        if (c.strict) { TP_FAIL(ST_DATA); return s; }
    } else {
        s.empty_run = 0;
        if (!c.allow_huffman) { TP_FAIL(ST_FALLBACK); return s; }
        uint32_t hlit = NLIT, hdist = NDIST;
        if (btype == 1) {
            for (uint32_t i = 0; i < NLIT; i++) T.lens[i] = (uint8_t)fixed_lit_len(i);
            for (uint32_t i = 0; i < NDIST; i++) T.lens[NLIT + i] = 5;
        } else {
            hlit = tb_get(br, 5) + 257;
            hdist = tb_get(br, 5) + 1;
            const uint32_t hclen = tb_get(br, 4) + 4;
            uint8_t pl[19];
            #pragma unroll 1
            for (uint32_t i = 0; i < 19; i++) pl[i] = 0;
            #pragma unroll 1
            for (uint32_t i = 0; i < hclen; i++) {
                tb_refill(br);
                pl[C_PRECODE_ORDER[i]] = (uint8_t)tb_get(br, 3);
            }
            // precode: 7-bit table in the (not yet built) distance table's space
            if (!tp_build(c.dst, c.NT, T, pl, 19, 1, 7, false)) { TP_FAIL(ST_DATA); return s; }
            const uint32_t total = hlit + hdist;
            uint32_t i = 0, prev = 0;
            #pragma unroll 1
            while (i < total) {
                tb_refill(br);
                const uint32_t e = c.dst[tb_peek(br, 7) * c.NT];
                const uint32_t l = e & 15u, sym = e >> 4;
                if (l == 0) { TP_FAIL(ST_OVERRUN); return s; }    // no code matches: the reference reads on until it overruns
                tb_drop(br, l);
                uint32_t rep = 1, val = sym;
                if (sym == 16) { rep = 3 + tb_get(br, 2); val = prev; }
                else if (sym == 17) { rep = 3 + tb_get(br, 3); val = 0; }
                else if (sym == 18) { rep = 11 + tb_get(br, 7); val = 0; }
                if (i + rep > total) { TP_FAIL(ST_DATA); return s; }
                #pragma unroll 1
                for (uint32_t j = 0; j < rep; j++) {
                    const uint32_t p = i + j;
                    T.lens[p < hlit ? p : NLIT + (p - hlit)] = (uint8_t)val;
                }
                i += rep;
                prev = val;
            }
            if (T.lens[256] == 0) { TP_FAIL(ST_DATA); return s; }
        }
        if (!tp_build(c.lit, c.NT, T, T.lens, hlit, 0, c.lit_bits, true)) { TP_FAIL(ST_DATA); return s; }
        if (!tp_build(c.dst, c.NT, T, T.lens + NLIT, hdist, 1, TP_DST_BITS, false)) { TP_FAIL(ST_DATA); return s; }
        s.state = TS_SYM;
        return s;
    }
    // stored or skipped block: the checks that follow every block
    if (o.v - o.al > c.max_out) { s.end_flags |= END_TOO_BIG; TP_FAIL(ST_DATA); }
    else if (s.bfinal) { s.end_flags |= END_FINAL; s.state = TS_DONE; }
    return s;
}

__device__ __forceinline__ void tp_ctx_init(TpCtx& c, const TpUnit& u, unsigned flags) {
    c.in_len = u.in_len; c.in_bits = u.in_len * 8;
    c.ops = u.ops; c.ops_cap = u.ops_cap;
    c.cap = (uint32_t)min(u.cap, (uint64_t)0xFFFFFF00u);
    c.max_out = (uint32_t)min(u.max_out, (uint64_t)0xFFFFFF00u);
    c.seg_stop = 0xFFFFFFFFu;
    c.stop_at = c.max_out + 1;
    c.stop_at_sync = u.stop_at_sync; c.strict = flags & 1u; c.allow_huffman = true;
}
__device__ __forceinline__ void tp_state_init(TpState& s, const TpCtx& c, const TpUnit& u, bool live, uint32_t* ring, uint32_t nthreads) {
    // ring: the CTA's [TB_RING][nthreads] word array in shared memory; this thread owns column threadIdx.x
    s.br.ring_sa = (uint32_t)__cvta_generic_to_shared(ring + threadIdx.x);
    s.br.ring_stride = nthreads * 4;
    s.st = ST_OK;
    s.state = live ? TS_BLOCK : TS_DONE;
    if (live && u.in_len >= (1ull << 33)) { s.st = ST_FALLBACK; s.state = TS_DONE; }     // word indices are 32-bit here
    s.br.skip = (uint32_t)(reinterpret_cast<uintptr_t>(u.in) & 3);
    s.br.wp = reinterpret_cast<const uint32_t*>(u.in - s.br.skip);
    s.br.nw = s.state == TS_DONE ? 0 : (uint32_t)((s.br.skip + u.in_len + 3) >> 2);
    if (s.state != TS_DONE) tb_seek(s.br, 0);                 // a thread without a unit never touches memory
    else { s.br.wi = 0; s.br.w0 = 0; s.br.bb = 0; s.br.bc = 0; }
    s.o.al = (uint32_t)(reinterpret_cast<uintptr_t>(u.out) & 3);
    s.o.ob = u.out - s.o.al;
    s.o.v = s.o.vlo = s.o.al;
    s.o.vlim = s.o.al + c.cap;
    to_window(s.o);
    s.o.lw = 0;
    s.nops = 0; s.empty_run = 0; s.end_flags = 0; s.bfinal = 0;
}
__device__ __forceinline__ void tp_finish(const TpCtx& c, TpState& s, TpResult& res) {
    asm volatile("cp.async.wait_all;" ::: "memory");        // nothing of this thread's ring traffic outlives its unit
    to_flush(s.o);
    if (s.st != ST_FALLBACK && tb_bitpos(s.br) > c.in_bits) s.st = ST_OVERRUN;
    res.in_end = (tb_bitpos(s.br) + 7) >> 3;
    res.out_len = s.o.v - s.o.al;
    res.status = s.st;
    res.end_flags = s.end_flags & ~END_SEG;
    res.nops = s.st == ST_FALLBACK ? 0 : s.nops;
    res.pad = 0;
}

__device__ __forceinline__ void tp_lut_init(uint32_t* s_lut, uint32_t tid) {
    for (uint32_t t = tid; t < TP_LUT_WORDS; t += blockDim.x) {
        if (t < 32) s_lut[t] = t < 29 ? ((uint32_t)C_LEN_BASE[t] | (len_extra_bits(t) << 16)) : 0;
        else { const uint32_t k = t - 32; s_lut[t] = k < 30 ? ((uint32_t)C_DIST_BASE[k] | (dist_extra_bits(k) << 16)) : 0; }
    }
}

// ---- unit descriptors -------------------------------------------------------------------------------
// chunk mode: unit i = candidate chunk i of one stream (see find_sync_kernel), decoded to out + i * CHUNK
struct ChunkUnits {
    const uint8_t* in; uint64_t n; const uint64_t* cand; uint64_t ncand; uint8_t* out; uint64_t cap; uint64_t* ops;
    const uint16_t* segnops;
    __device__ __forceinline__ uint64_t count() const { return ncand; }
    __device__ __forceinline__ TpUnit get(uint64_t i) const {
        TpUnit u;
        const uint64_t start = cand[i];
        const uint64_t end = cand[i + 1];                           // a well-formed chunk ends exactly there (cand[last + 1] = n)
        u.in = in + start; u.in_len = end - start;
        const uint64_t o0 = i * CHUNK;
        u.out = out + o0; u.cap = o0 >= cap ? 0 : min((uint64_t)CHUNK, cap - o0); u.max_out = CHUNK;
        u.ops = ops + i * OPS_PER_CHUNK; u.ops_cap = OPS_PER_CHUNK; u.stop_at_sync = true;
        return u;
    }
    __device__ __forceinline__ const uint16_t* seg_counts(uint64_t i) const { return segnops + i * NSEG; }
};
// batch mode: unit i = stream i; its op list lives at ops + ceil(out_off / 4), capacity cap / 4 - 1 (output
// regions are disjoint, so these are too)
struct BatchUnits {
    const uint8_t* in; const uint64_t* in_off; const uint64_t* in_len; uint8_t* out; const uint64_t* out_off;
    const uint64_t* out_cap; uint64_t nstreams; uint64_t* ops;
    __device__ __forceinline__ uint64_t count() const { return nstreams; }
    __device__ __forceinline__ TpUnit get(uint64_t i) const {
        TpUnit u;
        u.in = in + in_off[i]; u.in_len = in_len[i];
        const uint64_t oo = out_off[i], oc = out_cap[i];
        u.out = out + oo; u.cap = oc; u.max_out = ~0ull;
        const uint64_t o_lo = (oo + 3) >> 2, o_hi = (oo + oc) >> 2;
        u.ops = ops + o_lo; u.ops_cap = o_hi > o_lo ? (uint32_t)min(o_hi - o_lo, (uint64_t)0x7FFFFFFFu) : 0;
        u.stop_at_sync = false;
        return u;
    }
    __device__ __forceinline__ const uint16_t* seg_counts(uint64_t) const { return nullptr; }
};

// ---- pass A, generic: one thread per unit, private interleaved tables ---------------------------------
// Control flow is a per-thread state machine driven by ONE warp-wide loop: every trip all 32 lanes meet at
// the __any_sync vote, then each lane takes one step of its own unit (one symbol, or one block header).
// Without the vote the lanes of a warp drift apart after the first literal/match divergence and never
// reconverge -- the warp then issues every lane's instructions separately (measured: 5x slower).
template <class Units>
__global__ void __launch_bounds__(TP_THREADS)
inflate_symbols_kernel(Units U, TpResult* __restrict__ res, unsigned flags, unsigned long long* __restrict__ any_fallback) {
    extern __shared__ __align__(16) uint8_t tp_smem[];
    __shared__ uint32_t s_ring[TB_RING * TP_THREADS];
    uint32_t* s_lut = reinterpret_cast<uint32_t*>(tp_smem);
    uint16_t* tabs = reinterpret_cast<uint16_t*>(tp_smem + TP_LUT_WORDS * 4);
    tp_lut_init(s_lut, threadIdx.x);
    __syncthreads();
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < U.count();
    TpUnit u = {};
    if (live) u = U.get(i);
    TpTables T;
    TpCtx c;
    tp_ctx_init(c, u, flags);
    tp_ctx_tables(c, tabs + threadIdx.x, tabs + threadIdx.x + (size_t)(1u << TP_LIT_BITS) * blockDim.x, blockDim.x, s_lut, &T, TP_LIT_BITS);
    TpState s;
    tp_state_init(s, c, u, live, s_ring, blockDim.x);
    while (__any_sync(0xFFFFFFFFu, s.state != TS_DONE)) {
        if (s.state == TS_SYM) tp_step_symbol<false>(c, s);
        else if (s.state == TS_BLOCK) s = tp_step_block(c, s);
    }
    if (!live) return;
    TpResult r;
    tp_finish(c, s, r);
    res[i] = r;
    if (r.status == ST_FALLBACK) atomicOr(any_fallback, 1ull);
}

// ---- pass A, segmented: a CTA takes SG_CHUNKS indexed chunks, builds their tables, then its threads pull
// (chunk, segment) pairs from a shared queue until none is left -- segments differ a lot in symbol count
// (text vs image chunks), a fixed thread <-> segment mapping left half of the lanes idle on average.
constexpr uint32_t SG_CHUNKS = 4;                       // chunks per CTA
constexpr uint32_t SG_THREADS = 64;                     // one segment per thread (more chunks per CTA cost more in occupancy
                                                        // than the better balance wins back: measured)
constexpr uint32_t OPS_PER_SEG = OPS_PER_CHUNK / NSEG;  // 1024
struct __align__(16) SgChunk {
    uint16_t lit[1u << SG_LIT_BITS];
    uint16_t dst[1u << TP_DST_BITS];
    TpTables T;
    TpUnit u;
    TpResult r;                    // of the chunk's last segment (which also reads what follows the block)
    uint32_t seg_start[NSEG];      // bit offset of each segment's first symbol, from the chunk's first byte
    uint32_t seg_end[NSEG];        // where each segment's thread stopped
    uint16_t seg_nops[NSEG];
    uint8_t seg_ok[NSEG];
    uint32_t nseg;                 // 0: not indexed / not ours
    uint32_t bfinal;
    uint64_t unit;                 // global chunk index
};
constexpr uint32_t SG_SMEM_BYTES = SG_CHUNKS * sizeof(SgChunk) + TP_LUT_WORDS * 4 + 16;

// reads the segment index at the start of a chunk (common.cuh); returns the number of segments or 0
// (the index takes INDEX_BYTES_PER_SEG bytes per segment: word 0 says how many)
__device__ uint32_t sg_read_index(const uint8_t* in, uint64_t in_len, uint32_t* words /* [NSEG] */) {
    if (in_len < INDEX_BYTES_PER_SEG + 2) return 0;
    for (uint32_t w = 0; w < NSEG; w++) words[w] = 0;
    uint32_t ngroups = 4;
    for (uint32_t g = 0; g < ngroups; g++) {
        const uint8_t* b = in + 5 * g;
        const uint32_t b0 = b[0];
        if ((b0 & 0x87u) != 0x80u || b[1] != 0 || b[2] != 0 || b[3] != 0xFF || b[4] != 0xFF) return 0;
        words[g >> 2] |= ((b0 >> 3) & 15u) << (4 * (g & 3));
        if (g == 3) {
            if ((words[0] & 0x3FFu) != INDEX_MAGIC || (words[0] >> 14) != 0) return 0;
            ngroups = 4 * (((words[0] >> 10) & 15u) + 1);
            if (in_len < (uint64_t)5 * ngroups + 2) return 0;
        }
    }
    return ngroups / 4;
}

// Pass A0: one thread per chunk.  Indexed chunks are appended to `list` (the segment kernel then works on a
// dense list: no lanes wasted on stored chunks); a chunk without index is finished here if it consists of
// stored blocks (a few ops), and handed to the one-warp decoder if it is Huffman-coded.
constexpr uint32_t SG_BUCKETS = 32;      // indexed chunks are bucketed by compressed size (2 KiB steps)
__global__ void __launch_bounds__(256)
inflate_classify_kernel(ChunkUnits U, TpResult* __restrict__ res, uint32_t* __restrict__ blist, uint32_t* __restrict__ bcnt,
                        unsigned long long* __restrict__ nlist, unsigned flags, unsigned long long* __restrict__ any_fallback) {
    __shared__ uint32_t s_ring[TB_RING * 256];
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= U.count()) return;
    const TpUnit u = U.get(i);
    uint32_t words[NSEG];
    if (sg_read_index(u.in, u.in_len, words) >= 2) {
        // chunks of similar compressed size have similar symbol counts: the segment kernel takes them bucket by
        // bucket, largest first, so that the threads of a warp finish together and the longest chains start first
        const uint32_t b = (uint32_t)min(u.in_len >> 11, (uint64_t)SG_BUCKETS - 1);
        blist[(uint64_t)b * U.count() + atomicAdd(&bcnt[b], 1u)] = (uint32_t)i;
        atomicAdd(nlist, 1ull);
        return;
    }
    TpCtx c;
    tp_ctx_init(c, u, flags);
    c.lit = c.dst = nullptr; c.T = nullptr; c.NT = 1; c.lit_sa = c.dst_sa = c.lut_sa = 0; c.ntb = 2; c.lit_bits = TP_LIT_BITS;
    c.allow_huffman = false;
    TpState s;
    tp_state_init(s, c, u, true, s_ring, 256);
    while (s.state != TS_DONE) s = tp_step_block(c, s);
    TpResult r;
    tp_finish(c, s, r);
    res[i] = r;
    if (r.status == ST_FALLBACK) atomicOr(any_fallback, 1ull);
}

template <int OCC>
__global__ void __launch_bounds__(SG_THREADS, OCC)
inflate_segments_kernel(ChunkUnits U, const uint32_t* __restrict__ blist, const uint32_t* __restrict__ bcnt,
                        const unsigned long long* __restrict__ nlist, TpResult* __restrict__ res, uint16_t* __restrict__ segnops, unsigned flags,
                        unsigned long long* __restrict__ any_fallback) {
    const uint64_t nl = *nlist;
    if ((uint64_t)blockIdx.x * SG_CHUNKS >= nl) return;
    extern __shared__ __align__(16) uint8_t tp_smem[];
    __shared__ uint32_t s_ring[TB_RING * SG_THREADS];
    uint32_t* s_lut = reinterpret_cast<uint32_t*>(tp_smem);
    uint32_t* s_queue = s_lut + TP_LUT_WORDS;
    SgChunk* SCs = reinterpret_cast<SgChunk*>(tp_smem + TP_LUT_WORDS * 4 + 16);
    tp_lut_init(s_lut, threadIdx.x);
    if (threadIdx.x == 0) *s_queue = 0;
    __syncthreads();
    TpCtx c;
    TpState s;

    // ---- phase 1: every (SG_THREADS / SG_CHUNKS)-th thread prepares one chunk: index, block header, tables
    constexpr uint32_t BUILD_STRIDE = SG_THREADS / SG_CHUNKS;
    if (threadIdx.x % BUILD_STRIDE == 0) {
        SgChunk* SC = SCs + threadIdx.x / BUILD_STRIDE;
        const uint64_t li = (uint64_t)blockIdx.x * SG_CHUNKS + threadIdx.x / BUILD_STRIDE;
        SC->nseg = 0;
        if (li < nl) {
            // position li of the bucket-ordered list (largest bucket first)
            uint64_t i = 0, off = 0;
            for (int b = SG_BUCKETS - 1; b >= 0; b--) {
                const uint32_t nb = bcnt[b];
                if (li < off + nb) { i = blist[(uint64_t)b * U.count() + (li - off)]; break; }
                off += nb;
            }
            const TpUnit u = U.get(i);
            SC->u = u; SC->unit = i;
            tp_ctx_init(c, u, flags);
            tp_ctx_tables(c, SC->lit, SC->dst, 1, s_lut, &SC->T, SG_LIT_BITS);
            uint32_t words[NSEG];
            const uint32_t nseg = sg_read_index(u.in, u.in_len, words);
            tp_state_init(s, c, u, true, s_ring, SG_THREADS);
            if (nseg >= 2) {
                tb_seek(s.br, nseg * INDEX_BYTES_PER_SEG);
                s = tp_step_block(c, s);
            }
            if (nseg >= 2 && s.state == TS_SYM) {
                uint32_t bit = (uint32_t)tb_bitpos(s.br);
                SC->seg_start[0] = bit;
                for (uint32_t k = 1; k < nseg; k++) { bit += words[k]; SC->seg_start[k] = bit; }
                SC->nseg = nseg;
                SC->bfinal = s.bfinal;
            } else {
                // not a Huffman block after the index: let the one-warp decoder sort it out
                TpResult r;
                s.st = ST_FALLBACK; s.state = TS_DONE;
                tp_finish(c, s, r);
                res[i] = r;
                atomicOr(any_fallback, 1ull);
            }
        }
    }
    __syncthreads();

    // ---- phase 2: threads pull segments from the queue; every trip of the loop all lanes of a warp meet at
    // the vote, lanes with a segment decode one symbol, lanes without one fetch the next
    SgChunk* SC = SCs;
    uint32_t seg = 0;
    bool have = false, more = true;
    s.state = TS_DONE;
    while (__any_sync(0xFFFFFFFFu, more)) {
        if (s.state == TS_SYM) {
            tp_step_symbol<true>(c, s);
        } else if (more) {
            if (have) {
                // finish the segment: the chunk's last one also consumes what follows the block (the chunk
                // separator, or nothing after a final block; a second Huffman block -> one-warp decoder)
                while (s.state == TS_BLOCK) s = tp_step_block(c, s);
                if (seg + 1 < SC->nseg) {
                    asm volatile("cp.async.wait_all;" ::: "memory");
                    to_flush(s.o);
                    SC->seg_end[seg] = (uint32_t)tb_bitpos(s.br);
                    SC->seg_ok[seg] = s.st == ST_OK && (s.end_flags & END_SEG);
                } else {
                    TpResult r;
                    tp_finish(c, s, r);
                    SC->r = r;
                }
                SC->seg_nops[seg] = (uint16_t)s.nops;
                have = false;
            }
            const uint32_t q = atomicAdd(s_queue, 1u);
            if (q >= SG_CHUNKS * NSEG) {
                more = false;
            } else {
                SC = SCs + q / NSEG;
                seg = q % NSEG;
                const uint32_t nseg = SC->nseg;
                if (seg < nseg) {
                    const TpUnit u = SC->u;
                    tp_ctx_init(c, u, flags);
                    tp_ctx_tables(c, SC->lit, SC->dst, 1, s_lut, &SC->T, SG_LIT_BITS);
                    tp_state_init(s, c, u, true, s_ring, SG_THREADS);
                    const uint32_t bit = SC->seg_start[seg];
                    tb_seek(s.br, bit >> 3);
                    tb_drop(s.br, bit & 7);
                    s.o.v = s.o.vlo = s.o.al + seg * SEG;
                    to_window(s.o);
                    s.state = TS_SYM;
                    s.bfinal = SC->bfinal;
                    c.ops = u.ops + seg * OPS_PER_SEG;
                    c.ops_cap = OPS_PER_SEG;
                    c.seg_stop = seg + 1 < nseg ? (seg + 1) * SEG : 0xFFFFFFFFu;
                    c.stop_at = min(c.seg_stop, c.max_out + 1);
                    c.allow_huffman = false;              // one Huffman block per indexed chunk
                    have = true;
                }
            }
        }
    }
    __syncthreads();

    // ---- phase 3: the segments of a chunk must chain exactly, else the index was not ours
    if (threadIdx.x % BUILD_STRIDE == 0) {
        SgChunk* SK = SCs + threadIdx.x / BUILD_STRIDE;
        const uint32_t nseg = SK->nseg;
        if (nseg) {
            const uint64_t i = SK->unit;
            TpResult r = SK->r;
            bool ok = true;
            for (uint32_t k = 0; k + 1 < nseg; k++) ok = ok && SK->seg_ok[k] && SK->seg_end[k] == SK->seg_start[k + 1];
            if (!ok) { r.status = ST_FALLBACK; r.nops = 0; }
            else if (r.status != ST_FALLBACK) {
                r.nops = TP_SEGMENTED | nseg;
                for (uint32_t k = 0; k < nseg; k++) segnops[i * NSEG + k] = SK->seg_nops[k];
            }
            res[i] = r;
            if (r.status == ST_FALLBACK) atomicOr(any_fallback, 1ull);
        }
    }
}
#undef TP_FAIL

// ---- fallback: units pass A gave up on, one warp each (the decoder of inflate.cuh) -----------------------
template <class Units>
__global__ void __launch_bounds__(INF_THREADS)
inflate_fallback_kernel(Units U, TpResult* __restrict__ res, unsigned flags, const unsigned long long* __restrict__ any_fallback,
                        unsigned long long* __restrict__ counter) {
    if (*any_fallback == 0) return;
    __shared__ InfWarp S[INF_WARPS];
    __shared__ ModLut ML;
    modlut_init(&ML, threadIdx.x, INF_THREADS);
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t n = U.count();
    for (;;) {
        unsigned long long i = 0;
        if (lane == 0) i = atomicAdd(counter, 32ull);
        i = __shfl_sync(0xFFFFFFFFu, i, 0);
        if (i >= n) break;
        // 32 units per grab: lanes look at one status each
        const bool need = i + lane < n && res[i + lane].status == ST_FALLBACK;
        uint32_t m = __ballot_sync(0xFFFFFFFFu, need);
        while (m) {
            const uint32_t j = __ffs(m) - 1;
            m &= m - 1;
            const TpUnit u = U.get(i + j);
            uint64_t ol = 0, used = 0;
            uint32_t ef = 0;
            const int st = inflate_warp(&S[warp], &ML, u.in, u.in_len, u.out, u.cap, u.stop_at_sync, u.max_out, flags, lane, ol, used, ef);
            if (lane == 0) {
                TpResult r;
                r.in_end = used; r.out_len = ol; r.status = st; r.end_flags = ef; r.nops = 0; r.pad = 0;
                res[i + j] = r;
            }
        }
    }
}

// ---- pass B -------------------------------------------------------------------------------------------
// cooperative copy of one back-reference by all 32 lanes (same scheme as inflate_warp's)
__device__ __forceinline__ void tp_warp_match(uint8_t* dp, uint32_t dist, uint32_t length, uint32_t ncopy, const ModLut* ML, uint32_t lane) {
    const uint8_t* sp = dp - dist;
    if (dist >= length || dist >= 32) {
        if (length <= 32) {
            if (lane < ncopy) dp[lane] = sp[lane];
        } else {
            #pragma unroll 1
            for (uint32_t b = 0; b < length; b += 32) {
                const uint32_t i = b + lane;
                if (i < ncopy) dp[i] = sp[i];
                if (dist < length) __syncwarp();
            }
        }
    } else if (dist == 1) {
        const uint8_t v = ncopy ? sp[0] : 0;
        #pragma unroll 1
        for (uint32_t i = lane; i < ncopy; i += 32) dp[i] = v;
    } else {
        uint32_t r = ML->m[dist][lane];
        const uint32_t step = ML->m[dist][32];
        #pragma unroll 1
        for (uint32_t i = lane; i < ncopy; i += 32) {
            dp[i] = sp[r];
            r += step;
            if (r >= dist) r -= dist;
        }
    }
}
// stored block: input bytes -> output.  Destination-aligned 16-byte stores; the source is read as aligned
// 32-bit words and realigned with funnel shifts (five words give one 16-byte vector).
__device__ __forceinline__ void tp_warp_stored(uint8_t* dp, const uint8_t* sp, uint32_t n, uint32_t lane) {
    const uint32_t head = min(n, (uint32_t)((16 - (reinterpret_cast<uintptr_t>(dp) & 15)) & 15));
    if (lane < head) dp[lane] = sp[lane];
    const uint32_t vecs = (n - head) >> 4;
    const uint8_t* s = sp + head;
    const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(s) & 3) * 8;
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(s - (sh >> 3));
    uint4* dv = reinterpret_cast<uint4*>(dp + head);
    // sw[4 i + 4] of the last vector stays inside the aligned word that holds the block's last byte (sh != 0)
    #pragma unroll 2
    for (uint32_t i = lane; i < vecs; i += 32) {
        const uint32_t* w = sw + 4 * i;
        const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2), w3 = __ldg(w + 3);
        uint4 v;
        if (sh == 0) { v = make_uint4(w0, w1, w2, w3); }
        else {
            const uint32_t w4 = __ldg(w + 4);
            v = make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh));
        }
        dv[i] = v;
    }
    const uint32_t done = head + vecs * 16;
    if (done + lane < n) dp[done + lane] = sp[done + lane];
}

constexpr uint32_t TP_FREE_MAX = 16;     // longest op a single lane copies on its own

// Lane-local copy of n <= 16 bytes that do not overlap (dist >= n): 16 predicated byte loads with immediate
// offsets, then 16 predicated byte stores -- one memory round trip, no address arithmetic per byte.
#define TP_R16(M) M(0) M(1) M(2) M(3) M(4) M(5) M(6) M(7) M(8) M(9) M(10) M(11) M(12) M(13) M(14) M(15)
#define TP_SETP(i) "setp.gt.u32 p" #i ", %2, " #i ";\n"
#define TP_LD(i) "@p" #i " ld.global.u8 t" #i ", [%0+" #i "];\n"
#define TP_ST(i) "@p" #i " st.global.u8 [%1+" #i "], t" #i ";\n"
__device__ __forceinline__ void tp_lane_copy16(uint8_t* dp, const uint8_t* sp, uint32_t n) {
    asm volatile("{\n.reg .pred p<16>;\n.reg .b32 t<16>;\n" TP_R16(TP_SETP) TP_R16(TP_LD) TP_R16(TP_ST) "}\n"
                 :: "l"(sp), "l"(dp), "r"(n) : "memory");
}
// Lane-local copy of n <= 16 bytes of a pattern of period dist <= 8 that starts at sp (= dp - dist): the 8
// pattern bytes are packed into two registers, output byte i = pattern byte i % dist (byte permute with the
// selector nibbles of ML->nib[dist]).
#define TP_LDP(i) "@p" #i " ld.global.u8 t" #i ", [%2+" #i "];\n"
__device__ __forceinline__ void tp_lane_pattern16(uint8_t* dp, const uint8_t* sp, uint32_t dist, uint32_t n, uint64_t nib) {
    uint32_t lo, hi;
    asm volatile("{\n.reg .pred p<8>;\n.reg .b32 t<8>;\n"
                 "mov.b32 t0, 0; mov.b32 t1, 0; mov.b32 t2, 0; mov.b32 t3, 0; mov.b32 t4, 0; mov.b32 t5, 0; mov.b32 t6, 0; mov.b32 t7, 0;\n"
                 "setp.gt.u32 p0, %3, 0; setp.gt.u32 p1, %3, 1; setp.gt.u32 p2, %3, 2; setp.gt.u32 p3, %3, 3;\n"
                 "setp.gt.u32 p4, %3, 4; setp.gt.u32 p5, %3, 5; setp.gt.u32 p6, %3, 6; setp.gt.u32 p7, %3, 7;\n"
                 TP_LDP(0) TP_LDP(1) TP_LDP(2) TP_LDP(3) TP_LDP(4) TP_LDP(5) TP_LDP(6) TP_LDP(7)
                 "prmt.b32 t0, t0, t1, 0x0040; prmt.b32 t2, t2, t3, 0x0040; prmt.b32 %0, t0, t2, 0x5410;\n"
                 "prmt.b32 t4, t4, t5, 0x0040; prmt.b32 t6, t6, t7, 0x0040; prmt.b32 %1, t4, t6, 0x5410;\n"
                 "}\n" : "=r"(lo), "=r"(hi) : "l"(sp), "r"(dist) : "memory");
    #pragma unroll
    for (uint32_t i = 0; i < TP_FREE_MAX; i++) {
        const uint32_t v = __byte_perm(lo, hi, (uint32_t)(nib >> (4 * i)) & 7u);
        if (i < n) dp[i] = (uint8_t)v;
    }
}

// Applies one op list in order.  32 ops per step.  An op is READY when every byte it reads is final before
// the step starts: its source ends at or below the step's first destination, or begins at or after the end
// of the previous op (then it only reads literals, which pass A wrote), or it is the step's first op.  Ready
// ops of at most TP_FREE_MAX bytes are copied one per lane, all at once (one memory round trip for the
// whole step); the others go through the cooperative copy in order.
__device__ void tp_apply_ops(const uint64_t* __restrict__ ops, uint32_t nops, uint8_t* out, uint32_t cap,
                             const uint8_t* in, const ModLut* ML, uint32_t lane, unsigned tune) {
    const uint32_t FULL = 0xFFFFFFFFu;
    // the op words of the next two steps are always in flight (a step can be shorter than a DRAM round trip)
    uint64_t o_next = lane < nops ? ops[lane] : 0ull;
    uint64_t o_next2 = 32 + lane < nops ? ops[32 + lane] : 0ull;
    for (uint32_t b = 0; b < nops; b += 32) {
        const uint64_t o = o_next;
        o_next = o_next2;
        if (b + 64 < nops) o_next2 = b + 64 + lane < nops ? ops[b + 64 + lane] : 0ull;
        if (tune & 1u) {
            // the step after this one will read around these addresses: start pulling the lines in now
            const uint32_t npos = (uint32_t)o_next, ndist = (uint32_t)(o_next >> 48);
            if (ndist) asm volatile("prefetch.global.L2 [%0];" :: "l"(out + (npos - ndist)));
        }
        const uint32_t pos = (uint32_t)o, len = (uint32_t)(o >> 32) & 0xFFFFu, dist = (uint32_t)(o >> 48);
        // stored pairs: a header slot (dist 0, len > 0) is followed by a raw data slot whose bits mean nothing
        uint32_t hdrs = __ballot_sync(FULL, len != 0 && dist == 0), data = 0;
        for (uint32_t h = hdrs; h;) {
            const uint32_t j = __ffs(h) - 1;
            if (j < 31) data |= 2u << j;
            h &= ~(3u << j);
        }
        hdrs &= ~data;
        const bool is_data = (data >> lane) & 1u;
        const bool is_match = !is_data && len != 0 && dist != 0;
        const uint32_t matches = __ballot_sync(FULL, is_match);
        const uint32_t live = matches | hdrs;
        if (!live) continue;
        const uint32_t first = __ffs(live) - 1;
        const uint32_t p0 = __shfl_sync(FULL, pos, first);
        const uint32_t ncopy = pos < cap ? min(len, cap - pos) : 0;
        const uint32_t prev_end = __shfl_up_sync(FULL, pos + len, 1);
        const bool prev_live = lane != 0 && ((live >> (lane - 1)) & 1u);
        const uint32_t a = pos - dist, m = min(len, dist);    // the op reads [a, a + m)
        const bool local = is_match && len <= TP_FREE_MAX && (dist >= len || dist <= 8);
        const bool ready = local && (a + m <= p0 || lane == first || (prev_live && a >= prev_end));
        const bool plain = ready && dist >= len;
        if (__any_sync(FULL, plain)) {
            if (plain) tp_lane_copy16(out + pos, out + a, ncopy);
        }
        if (__any_sync(FULL, ready && !plain)) {
            // overlapping (dist < len): the dist bytes below the destination repeat; they are final as well
            if (ready && !plain) tp_lane_pattern16(out + pos, out + a, dist, ncopy, ML->nib[dist]);
        }
        __syncwarp();
        uint32_t seq = live & ~__ballot_sync(FULL, ready);
        while (seq) {
            const uint32_t j = __ffs(seq) - 1;
            seq &= seq - 1;
            const uint32_t jpos = __shfl_sync(FULL, pos, j);
            const uint32_t jlen = __shfl_sync(FULL, len, j);
            const uint32_t jdist = __shfl_sync(FULL, dist, j);
            const uint32_t jn = __shfl_sync(FULL, ncopy, j);
            if (jdist) {
                tp_warp_match(out + jpos, jdist, jlen, jn, ML, lane);
            } else {
                const uint32_t lo = __shfl_sync(FULL, (uint32_t)o, (j + 1) & 31), hi = __shfl_sync(FULL, (uint32_t)(o >> 32), (j + 1) & 31);
                tp_warp_stored(out + jpos, in + (((uint64_t)hi << 32) | lo), jn, lane);
            }
            __syncwarp();
        }
    }
}

template <class Units, int OCC>
__global__ void __launch_bounds__(INF_THREADS, OCC)
inflate_copy_kernel(Units U, const TpResult* __restrict__ res, unsigned long long* __restrict__ counter, unsigned tune) {
    __shared__ ModLut ML;
    modlut_init(&ML, threadIdx.x, INF_THREADS);
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t n = U.count();
    for (;;) {
        unsigned long long i = 0;
        if (lane == 0) i = atomicAdd(counter, 1ull);
        i = __shfl_sync(0xFFFFFFFFu, i, 0);
        if (i >= n) break;
        const uint32_t nops = res[i].nops;
        if (!nops) continue;
        const TpUnit u = U.get(i);
        const uint32_t cap = (uint32_t)min(u.cap, (uint64_t)0xFFFFFF00u);
        if (nops & TP_SEGMENTED) {
            const uint16_t* sn = U.seg_counts(i);
            const uint32_t nseg = nops & 0xFFu;
            for (uint32_t k = 0; k < nseg; k++) tp_apply_ops(u.ops + k * OPS_PER_SEG, sn[k], u.out, cap, u.in, &ML, lane, tune);
        } else {
            tp_apply_ops(u.ops, nops, u.out, cap, u.in, &ML, lane, tune);
        }
    }
}

// chunk mode verdict (same rule as validate_chunks_kernel): result[0] = 1 if the optimistic layout is right,
// result[1] = total decoded bytes.  last_is_final = 0: these chunks are a prefix of a longer stream, every one of
// them must be a full chunk that ends at the next candidate (cand[ncand] is valid).
__global__ void validate_units_kernel(const uint64_t* __restrict__ cand, uint64_t ncand, uint64_t n,
                                      const TpResult* __restrict__ res, unsigned long long* __restrict__ result,
                                      int last_is_final) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncand) return;
    const TpResult r = res[i];
    bool ok = r.status == ST_OK;
    if (i + 1 < ncand || !last_is_final) {
        ok = ok && cand[i] + r.in_end == cand[i + 1] && r.out_len == CHUNK && (r.end_flags & END_SYNC);
        if (ok && i + 1 == ncand) result[1] = ncand * CHUNK;
    } else {
        ok = ok && (r.end_flags & END_FINAL);
        if (ok) result[1] = i * CHUNK + r.out_len;
    }
    if (!ok) atomicAnd(&result[0], 0ull);
}

// batch mode epilogue: TpResult -> the API's out_len / status arrays
__global__ void batch_results_kernel(const TpResult* __restrict__ res, uint64_t n, uint64_t* __restrict__ out_len,
                                     int32_t* __restrict__ status) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out_len[i] = res[i].out_len;
    status[i] = res[i].status;
}

// max over streams of out_off + out_cap (sizes the op-list scratch)
__global__ void batch_span_kernel(const uint64_t* __restrict__ out_off, const uint64_t* __restrict__ out_cap, uint64_t n,
                                  unsigned long long* __restrict__ span) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long v = i < n ? out_off[i] + out_cap[i] : 0ull;
    for (int o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
    if ((threadIdx.x & 31) == 0 && v) atomicMax(span, v);
}

}  // namespace b200
