// inflate_foreign.cuh -- block-parallel INFLATE of ONE stream this library did not write (zlib, the reference's
// own compressor, ...): SURVEY.md 8(f) rank 1.
//
// Replaces, for long foreign streams, the serial block loop of the reference's realDecompress
// (include/inflate.hpp:277-322): one growing output that every block may reference up to 32 KiB back
// (decompressHuffmanBlock, :262-272) and blocks that start at arbitrary BIT offsets.  Nothing in such a stream says
// where a block starts or what the 32 KiB before it will hold, so the work is cut in two the way pugz / rapidgzip do
// on CPUs, with the pieces laid out for a GPU:
//
//   F1 foreign_find_blocks_kernel   one warp per 16 KiB piece of the input tests EVERY bit offset of its piece as the
//        start of a dynamic-Huffman block: 17 header bits (BFINAL 0, BTYPE 10, HLIT <= 29, HDIST <= 29) pass 1 offset
//        in 9; survivors are queued in shared memory and checked 32 at a time for a COMPLETE precode (1-3 % pass) and
//        then by decoding the code-length lists (exact count, complete literal/length code with an end-of-block
//        symbol, complete / single / empty distance code).  What passes all three by chance is ~1e-9 per offset; the
//        chain check below catches it.  The piece's first hit is a candidate unit start.
//   F2 foreign_decode_kernel<COUNT>  one THREAD per candidate decodes from its bit offset, block after block, until a
//        block ends at or behind the next candidate: end bit, output bytes, number of back-references.  The host
//        walks the chain from bit 0 (unit k + 1 must start exactly where unit k ended; candidates the chain jumps
//        over were false, a missing one is added and counted in a second round) and prefix-sums the sizes.
//   F3 foreign_decode_kernel<EMIT>   the same decode with known offsets: literals go to a u16 SYMBOL image S of the
//        output (one u16 per output byte), back-references and stored blocks to the unit's op list.
//   F4 foreign_copy_kernel           one warp per unit applies the ops in order INSIDE S: a source below the unit's
//        first byte is not known yet and becomes a marker 0x8000 | index into the 32 KiB window before the unit;
//        markers are copied like bytes, so afterwards every symbol of a unit is a literal or a window marker.
//   F5 foreign_window_kernel         window propagation.  The last 32 KiB of every unit are resolved against the
//        window before it, unit after unit, in a ring in shared memory.  Sequential over units but two-level: groups
//        of units compose their windows symbolically in parallel (the ring starts as identity markers), one CTA
//        chains the ~150 group maps, then the groups run again with their real start window and write the tails.
//   F6 foreign_resolve_kernel        everything that is not a tail: out[p] = literal or out[unit_start - 32768 + idx].
//
// Results are bit-identical to the one-warp decoder in inflate.cuh (same decisions, reference quirks included: the
// only position-dependent one -- a distance that reaches before the start of the output copies nothing,
// inflate.hpp:268-270 -- can only happen in the first 32 KiB and is decided by unit 0, which knows its position).
// Any error, a chain that does not close in a few rounds, or an output buffer smaller than the stream hands the
// stream back to the sequential decoder, which also produces the reference's error codes.
#pragma once
#include "inflate_tp.cuh"

namespace b200 {

constexpr uint32_t FB_PIECE = 16384;                 // input bytes whose bit offsets one warp tests
constexpr uint32_t FB_WARPS = 4;
constexpr uint32_t FB_THREADS = FB_WARPS * 32;
constexpr uint32_t FB_QUEUE1 = 1024 + 64;            // stage-1 survivors of one sweep of 32 words (all 1024 offsets in the worst case)
constexpr uint32_t FB_QUEUE = 96;                    // stage-2 survivors (flushed when >= 32 are waiting)
constexpr uint32_t FB_HDR_MAX_BITS = 2400;           // 17 + 19 * 3 + 316 * (7 + 7) rounded up: no dynamic header is longer
struct __align__(16) FbWarp {
    uint32_t q1[FB_QUEUE1], q2[FB_QUEUE];            // bit offsets relative to the piece
    uint8_t pre[128 * 32];                           // per-lane 7-bit precode table: entry e of lane l at [e * 32 + l]
};

// The piece is read straight from global memory: the 32 lanes of a warp look at 32 consecutive bit offsets, i.e. at the
// same two or three words -- one L1 line, broadcast.  (A shared-memory copy of the piece costs 17 KB per warp, which
// leaves 8 warps per SM; the scan is a chain of dependent loads and wants the occupancy more than the lower latency:
// measured 23.5 ms -> see DESIGN.md.)  w = the stream as aligned 32-bit words, word indices are clamped to the last
// word of the stream (offsets that close to the end cannot start a block that matters).
struct FbSrc { const uint32_t* w; uint32_t last; };  // last: highest word index that may be read, relative to w
__device__ __forceinline__ uint32_t fb_word(const FbSrc& s, uint32_t i) { return __ldg(s.w + min(i, s.last)); }
__device__ __forceinline__ uint64_t fb_bits64(const FbSrc& s, uint32_t q) {
    const uint32_t i = q >> 5, sh = q & 31;
    const uint32_t a = fb_word(s, i), b = fb_word(s, i + 1), c = fb_word(s, i + 2);
    return (uint64_t)__funnelshift_r(a, b, sh) | ((uint64_t)__funnelshift_r(b, c, sh) << 32);
}
__device__ __forceinline__ uint32_t fb_bits32(const FbSrc& s, uint32_t q) {
    const uint32_t i = q >> 5;
    return __funnelshift_r(fb_word(s, i), fb_word(s, i + 1), q & 31);
}

// stage 2: the precode (HCLEN x 3 bits from bit 17) must be a complete prefix code: sum of 2^-len == 1.  The sum comes
// from a table indexed by four lengths at a time (klut[i] = sum over the four 3-bit fields of i of 128 >> len, 0 for len 0).
__device__ __forceinline__ bool fb_precode_complete(const FbSrc& w, uint32_t q, const uint16_t* __restrict__ klut) {
    const uint32_t hclen = (fb_bits32(w, q + 13) & 15u) + 4;
    uint64_t pb = fb_bits64(w, q + 17);
    const uint32_t nb = 3 * hclen;                                 // 12 .. 57
    pb &= (1ull << nb) - 1ull;
    const uint32_t kraft = klut[(uint32_t)pb & 4095u] + klut[(uint32_t)(pb >> 12) & 4095u] + klut[(uint32_t)(pb >> 24) & 4095u] +
                           klut[(uint32_t)(pb >> 36) & 4095u] + klut[(uint32_t)(pb >> 48) & 4095u];
    return kraft == 128u;
}

// stage 3: decode HLIT + HDIST code lengths through the precode and check what a valid block guarantees
__device__ bool fb_header_valid(const FbSrc& w, uint32_t q, uint8_t* pre /* this lane's column */) {
    const uint64_t h = fb_bits64(w, q);
    const uint32_t hlit = ((uint32_t)(h >> 3) & 31u) + 257, hdist = ((uint32_t)(h >> 8) & 31u) + 1;
    const uint32_t hclen = ((uint32_t)(h >> 13) & 15u) + 4;
    uint32_t pl[19];
    #pragma unroll
    for (uint32_t i = 0; i < 19; i++) pl[i] = 0;
    #pragma unroll
    for (uint32_t i = 0; i < 19; i++) {
        const uint32_t v = i < hclen ? (fb_bits32(w, q + 17 + 3 * i) & 7u) : 0u;
        // C_PRECODE_ORDER[i] with a compile-time index after unrolling
        constexpr uint8_t ORD[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
        pl[ORD[i]] = v;
    }
    // canonical precode -> 128-entry table (sym | len << 5); the code is complete (stage 2), so every entry is set
    uint32_t cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    #pragma unroll
    for (uint32_t s = 0; s < 19; s++) cnt[pl[s]]++;
    uint32_t next[8];
    {
        uint32_t code = 0;
        cnt[0] = 0;
        #pragma unroll
        for (uint32_t l = 1; l <= 7; l++) { code = (code + cnt[l - 1]) << 1; next[l] = code; }
    }
    #pragma unroll
    for (uint32_t s = 0; s < 19; s++) {
        const uint32_t l = pl[s];
        if (!l) continue;
        const uint32_t rev = __brev(next[l]++) >> (32 - l);
        for (uint32_t k = rev; k < 128; k += 1u << l) pre[k * 32] = (uint8_t)(s | (l << 5));
    }
    uint32_t pos = q + 17 + 3 * hclen;
    const uint32_t pos_max = q + FB_HDR_MAX_BITS;
    const uint32_t total = hlit + hdist;
    uint32_t i = 0, prev = 0, klit = 0, kdist = 0, ndist = 0, eob = 0;
    while (i < total) {
        const uint32_t b = fb_bits32(w, pos);
        const uint32_t e = pre[(b & 127u) * 32];
        const uint32_t l = e >> 5, sym = e & 31u;
        pos += l;
        uint32_t rep = 1, val = sym;
        if (sym >= 16) {
            const uint32_t x = b >> l;
            if (sym == 16) { if (i == 0) return false; rep = 3 + (x & 3u); val = prev; pos += 2; }
            else if (sym == 17) { rep = 3 + (x & 7u); val = 0; pos += 3; }
            else { rep = 11 + (x & 127u); val = 0; pos += 7; }
        }
        if (i + rep > total) return false;
        if (val) {
            // a run may straddle the literal/length -> distance boundary (RFC 1951 3.2.7)
            const uint32_t in_lit = i < hlit ? min(rep, hlit - i) : 0;
            klit += in_lit * (32768u >> val);
            kdist += (rep - in_lit) * (32768u >> val);
            ndist += rep - in_lit;
            if (i <= 256 && 256 < i + rep) eob = 1;
        }
        i += rep;
        prev = val;
        // over-subscribed already: no need to read the rest (random bits get here after a few dozen lengths, a tenth of a header)
        if (klit > 32768u || kdist > 32768u || pos > pos_max) return false;
    }
    if (!eob || klit != 32768u) return false;
    // distance code: complete, or a single code of length 1 (zlib accepts exactly that), or none at all
    return kdist == 32768u || (ndist == 1 && kdist == 16384u) || ndist == 0;
}

// cand[p] = absolute bit offset of the first valid-looking dynamic block header that STARTS in piece p, or ~0
__global__ void __launch_bounds__(FB_THREADS)
foreign_find_blocks_kernel(const uint8_t* __restrict__ in, uint64_t n, uint64_t npieces, unsigned long long* __restrict__ cand) {
    __shared__ FbWarp fb_warps[FB_WARPS];
    __shared__ uint16_t klut[4096];
    for (uint32_t i = threadIdx.x; i < 4096; i += FB_THREADS) {
        uint32_t k = 0;
        #pragma unroll
        for (uint32_t f = 0; f < 4; f++) { const uint32_t l = (i >> (3 * f)) & 7u; k += l ? (128u >> l) : 0u; }
        klut[i] = (uint16_t)k;
    }
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t p = (uint64_t)blockIdx.x * FB_WARPS + warp;
    if (p >= npieces) return;
    FbWarp* W = &fb_warps[warp];
    const uint32_t FULL = 0xFFFFFFFFu;
    const uint64_t byte0 = p * FB_PIECE;
    // the stream as aligned words: the piece starts `qoff` bits into word `word0`
    const uint32_t skip = (uint32_t)(reinterpret_cast<uintptr_t>(in) & 3);
    const uint32_t* wall = reinterpret_cast<const uint32_t*>(in - skip);
    const uint64_t nwords = (skip + n + 3) >> 2;
    const uint64_t word0 = (skip + byte0) >> 2;
    const uint32_t qoff = (uint32_t)((skip + byte0) & 3) * 8;
    FbSrc w;
    w.w = wall + word0;
    w.last = (uint32_t)min(nwords - 1 - word0, (uint64_t)0x7FFFFFFFu);
    const uint64_t avail = n - byte0;
    // offsets tested: [0, nbits) of the piece; the last bytes of the stream cannot hold a header plus a block worth finding
    uint32_t nbits = (uint32_t)min((uint64_t)FB_PIECE, avail) * 8;
    if (avail < FB_PIECE + 16) nbits = avail > 16 ? (uint32_t)(avail - 16) * 8 : 0;
    uint32_t best = 0xFFFFFFFFu;
    uint32_t n1 = 0, n2 = 0;
    auto drain2 = [&](uint32_t count) {          // stage 3 on the first `count` entries of q2
        const bool have = lane < count;
        const uint32_t q = have ? W->q2[lane] : 0;
        bool ok = false;
        if (have && q < best) ok = fb_header_valid(w, q + qoff, W->pre + lane);
        uint32_t m = __ballot_sync(FULL, ok);
        while (m) {
            const uint32_t j = __ffs(m) - 1;
            m &= m - 1;
            best = min(best, __shfl_sync(FULL, q, j));
        }
        __syncwarp();
        // compact the queue
        const uint32_t rest = n2 - count;
        uint32_t v = 0;
        if (lane < rest) v = W->q2[count + lane];
        uint32_t v2 = 0;
        if (lane + 32 < rest) v2 = W->q2[count + lane + 32];
        __syncwarp();
        if (lane < rest) W->q2[lane] = v;
        if (lane + 32 < rest) W->q2[lane + 32] = v2;
        n2 = rest;
        __syncwarp();
    };
    // Stage 1 is bit-parallel: a lane takes ONE word of the piece and tests its 32 bit offsets at once on a 64-bit window
    // x (bit j of every term below speaks for the header that starts at bit j):
    //   BFINAL = 0, BTYPE = 10      ~x & ~(x >> 1) & (x >> 2)
    //   HLIT  <= 29  (bits 3..7)    not all of bits 4..7 set
    //   HDIST <= 29  (bits 8..12)   not all of bits 9..12 set
    // 1 offset in 9 survives; the survivors of 32 words are queued in offset order and go through stage 2, 32 at a time.
    const uint32_t nwords_piece = (qoff + nbits + 31) >> 5;
    for (uint32_t w0 = 0; w0 < nwords_piece; w0 += 32) {
        if (best != 0xFFFFFFFFu && w0 * 32 > best + qoff) break;       // only the FIRST hit of the piece matters
        const uint32_t wi = w0 + lane;
        const uint64_t x = (uint64_t)fb_word(w, wi) | ((uint64_t)fb_word(w, wi + 1) << 32);
        uint64_t m64 = ~x & ~(x >> 1) & (x >> 2);
        m64 &= ~((x >> 4) & (x >> 5) & (x >> 6) & (x >> 7));
        m64 &= ~((x >> 9) & (x >> 10) & (x >> 11) & (x >> 12));
        uint32_t m = (uint32_t)m64;
        // offsets of this word that lie inside [qoff, qoff + nbits)
        const uint32_t lo = wi * 32, hi = lo + 32;
        if (wi >= nwords_piece) m = 0;
        else {
            if (lo < qoff) m &= 0xFFFFFFFFu << (qoff - lo);
            if (hi > qoff + nbits) m &= (qoff + nbits > lo) ? (0xFFFFFFFFu >> (hi - (qoff + nbits))) : 0u;
        }
        const uint32_t cnt = __popc(m);
        uint32_t incl = cnt;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(FULL, incl, o);
            if ((int)lane >= o) incl += u;
        }
        uint32_t slot = n1 + incl - cnt;
        while (m) {
            const uint32_t j = __ffs(m) - 1;
            m &= m - 1;
            W->q1[slot++] = lo + j - qoff;
        }
        n1 += __shfl_sync(FULL, incl, 31);
        __syncwarp();
        // stage 2 over everything queued (the queue is in offset order, so is what survives)
        for (uint32_t r0 = 0; r0 < n1; r0 += 32) {
            const bool have = r0 + lane < n1;
            const uint32_t q = have ? W->q1[r0 + lane] : 0;
            const bool ok = have && q < best && fb_precode_complete(w, q + qoff, klut);
            const uint32_t mk = __ballot_sync(FULL, ok);
            if (ok) W->q2[n2 + __popc(mk & ((1u << lane) - 1u))] = q;
            n2 += __popc(mk);
            __syncwarp();
            if (n2 >= 32) drain2(32);
        }
        n1 = 0;
        __syncwarp();
    }
    while (n2) drain2(min(n2, 32u));
    if (lane == 0) cand[p] = best == 0xFFFFFFFFu ? ~0ull : (unsigned long long)(byte0 * 8 + best);
}

// ---- F2 / F3: one thread per unit ---------------------------------------------------------------------------
// Threads per CTA x table widths: every thread owns private, interleaved u16 tables in shared memory, so the table size
// decides how many threads (= units in flight) an SM holds.  The decode is a chain of dependent shared-memory lookups:
// with 4 warps per SM every one of its latencies is exposed.  Smaller tables send more (rarer) codes through the
// canonical search in tp_slow_symbol, more threads hide more latency; variants are measured in DESIGN.md.
constexpr uint32_t fd_smem_bytes(uint32_t lb, uint32_t db, uint32_t nth) { return ((1u << lb) + (1u << db)) * 2u * nth + TP_LUT_WORDS * 4u; }
constexpr uint32_t FU_FINAL = 1;                                  // the unit ended with a BFINAL block
constexpr uint32_t FU_REACHED = 2;                                // ... at a block end at or behind its stop offset
struct FUnitRes { uint64_t end_bit; uint64_t out_len; uint32_t nops; int32_t status; uint32_t flags; uint32_t pad; };
constexpr uint32_t F_MARK = 0x8000u;                              // symbol image: 0x8000 | index into the 32 KiB window before the unit
constexpr uint32_t F_WINDOW = 32768;
constexpr uint32_t FD_MAX_OUT = 0xF0000000u;                      // positions inside a unit are 32-bit

struct FdState {
    TBits br;
    uint32_t pos;            // bytes produced by this unit
    uint32_t nops;
    uint32_t state, bfinal, flags;
    uint32_t btype;          // of the block whose tables are about to be built (TS_TABLES)
    int st;
};
constexpr uint32_t TS_TABLES = 3;        // a Huffman block's 3 header bits are read, its tables are not built yet
// Table builds are ~16 000 instructions of single-thread work per block (code lengths through the precode, two table
// fills); run by whichever lane happens to reach one they serialise: 31 lanes wait while one works, 32 times per round
// of blocks -- a third of the decode kernels' instruction stream.  A lane that reaches a Huffman block therefore WAITS
// (the others keep decoding) until FD_TAB_BATCH lanes are waiting, nobody is decoding any more, or FD_TAB_WAIT trips of
// the symbol loop have passed; the waiting lanes then build their tables together, the same instructions side by side.
constexpr uint32_t FD_TAB_BATCH = 4;
constexpr uint32_t FD_TAB_WAIT = 1024;

// Codes longer than the direct table's index.  tp_slow_symbol() walks the lengths one by one with two local-memory loads
// per step (~300 cycles), and in a warp whose 32 lanes decode 32 different units the other lanes wait for it: with a
// 9-bit table one trip in three has such a lane.  The canonical code makes a search without a loop possible: left-aligned
// to 15 bits, the codes of length l fill [first[l] << (15 - l), (first[l] + count[l]) << (15 - l)), and these ranges follow one
// another in order of length -- so the length of a code is the smallest l whose upper bound lies above it.  Per thread and
// per length above the table width one word in shared memory ([slot][thread], conflict-free):
//   [15:0] upper bound (left-aligned, <= 0x8000), [31:16] offs[l] - first[l] (mod 2^16): index into sorted[] = that + code.
template <uint32_t TB>
__device__ __forceinline__ void fd_slow_fill(uint32_t* slow /* this thread's slot 0 */, uint32_t NT, const TpTables& T, uint32_t which) {
    #pragma unroll 1
    for (uint32_t l = TB + 1; l <= 15; l++) {
        const uint32_t lim = ((uint32_t)T.first[which][l] + T.count[which][l]) << (15 - l);
        const uint32_t base = ((uint32_t)T.offs[which][l] - T.first[which][l]) & 0xFFFFu;
        slow[(l - TB - 1) * NT] = lim | (base << 16);
    }
}
// returns symbol | length << 16, or -1 (no code starts with these bits); sorted: the alphabet's part of TpTables::sorted
template <uint32_t TB>
__device__ __forceinline__ int fd_slow_symbol(uint32_t slow_sa, uint32_t stride, const uint16_t* sorted, uint32_t bits32) {
    const uint32_t code15 = __brev(bits32) >> 17;
    uint32_t sel = 0, len = 0;
    // ascending with an early exit: most long codes are one bit longer than the table (measured: the unrolled, branch-free
    // form -- always all words -- made the 9-bit variant 15 % slower)
    #pragma unroll 1
    for (uint32_t k = 0; k < 15 - TB; k++) {
        const uint32_t w = lds_u32(slow_sa + k * stride);
        if (code15 < (w & 0xFFFFu)) { sel = w; len = TB + 1 + k; break; }
    }
    if (!len) return -1;
    const uint32_t idx = ((sel >> 16) + (code15 >> (15 - len))) & 0xFFFFu;
    return (int)(sorted[idx] | (len << 16));
}

// One block header (single thread): stored blocks become ops, Huffman blocks get their tables built.  The decoder state
// travels BY VALUE (as in tp_step_block): passed by reference to this non-inlined function it has an address, lives in local
// memory for the whole kernel, and every bit-buffer operation of the symbol loop becomes a local load and store (the first
// version did exactly that: 955 LDL / STL instructions in the kernel, F2 + F3 = 37.6 ms).
// LB / DB: index widths of the thread's private literal/length and distance tables (the 7-bit precode table borrows the
// literal table's space, which is not built yet when it is needed).
template <bool EMIT>
__device__ __noinline__ FdState fd_block(FdState s, uint64_t in_len, uint64_t in_bits, bool strict, uint64_t stop_bit, uint64_t* ops) {
    TBits& br = s.br;
    if (tb_bitpos(br) + 3 > in_bits) { s.st = ST_OVERRUN; s.state = TS_DONE; return s; }
    tb_refill(br);
    const uint32_t hdr = tb_get(br, 3);
    s.bfinal = hdr & 1;
    const uint32_t btype = hdr >> 1;
    if (btype == 0) {
        tb_drop(br, br.bc & 7);
        tb_refill(br);
        const uint32_t len = tb_get(br, 16);
        const uint32_t nlen = tb_get(br, 16);
        if (strict && (len ^ nlen) != 0xFFFFu) { s.st = ST_DATA; s.state = TS_DONE; return s; }
        const uint64_t bpos = tb_bitpos(br) >> 3;
        if (bpos + len > in_len) { s.st = ST_OVERRUN; s.state = TS_DONE; return s; }
        if (len) {
            if (EMIT) { ops[s.nops] = tp_op(s.pos, len, 0); ops[s.nops + 1] = bpos; }
            s.nops += 2;
            s.pos += len;
        }
        tb_seek(br, bpos + len);
    } else if (btype == 3) {
        if (strict) { s.st = ST_DATA; s.state = TS_DONE; return s; }           // the reference's switch has no case 3: skipped
    } else {
        s.btype = btype;
        s.state = TS_TABLES;
        return s;
    }
    // stored or skipped block: what follows every block
    if (s.pos > FD_MAX_OUT) { s.st = ST_FALLBACK; s.state = TS_DONE; }
    else if (s.bfinal) { s.flags |= FU_FINAL; s.state = TS_DONE; }
    else if (tb_bitpos(br) >= stop_bit) { s.flags |= FU_REACHED; s.state = TS_DONE; }
    return s;
}

// The tables of a fixed (btype 1) or dynamic (btype 2) block whose 3 header bits have been read.
template <uint32_t LB, uint32_t DB>
__device__ __noinline__ FdState fd_tables(FdState s, uint16_t* lit, uint16_t* dst, uint32_t* slow, uint32_t NT, TpTables& T) {
    TBits& br = s.br;
    uint32_t hlit = NLIT, hdist = NDIST;
    if (s.btype == 1) {
        for (uint32_t i = 0; i < NLIT; i++) T.lens[i] = (uint8_t)fixed_lit_len(i);
        for (uint32_t i = 0; i < NDIST; i++) T.lens[NLIT + i] = 5;
    } else {
        tb_refill(br);
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
        static_assert(LB >= 7, "the precode table needs 128 entries");
        if (!tp_build(lit, NT, T, pl, 19, 1, 7, false)) { s.st = ST_DATA; s.state = TS_DONE; return s; }
        const uint32_t total = hlit + hdist;
        uint32_t i = 0, prev = 0;
        #pragma unroll 1
        while (i < total) {
            tb_refill(br);
            const uint32_t e = lit[tb_peek(br, 7) * NT];
            const uint32_t l = e & 15u, sym = e >> 4;
            if (l == 0) { s.st = ST_OVERRUN; s.state = TS_DONE; return s; }
            tb_drop(br, l);
            uint32_t rep = 1, val = sym;
            if (sym == 16) { rep = 3 + tb_get(br, 2); val = prev; }
            else if (sym == 17) { rep = 3 + tb_get(br, 3); val = 0; }
            else if (sym == 18) { rep = 11 + tb_get(br, 7); val = 0; }
            if (i + rep > total) { s.st = ST_DATA; s.state = TS_DONE; return s; }
            #pragma unroll 1
            for (uint32_t j = 0; j < rep; j++) {
                const uint32_t p = i + j;
                T.lens[p < hlit ? p : NLIT + (p - hlit)] = (uint8_t)val;
            }
            i += rep;
            prev = val;
        }
        if (T.lens[256] == 0) { s.st = ST_DATA; s.state = TS_DONE; return s; }
    }
    if (!tp_build(lit, NT, T, T.lens, hlit, 0, LB, true)) { s.st = ST_DATA; s.state = TS_DONE; return s; }
    fd_slow_fill<LB>(slow, NT, T, 0);
    if (!tp_build(dst, NT, T, T.lens + NLIT, hdist, 1, DB, false)) { s.st = ST_DATA; s.state = TS_DONE; return s; }
    fd_slow_fill<DB>(slow + (15 - LB) * NT, NT, T, 1);
    s.state = TS_SYM;
    return s;
}

// Persistent: the grid is one CTA per SM (the private tables fill the shared memory) and every THREAD pulls the next
// unit from a global counter when it has finished one -- units differ a lot in length (a wave of fixed assignments takes
// as long as its longest unit; measured: 2 waves x the longest = 18 ms, pulled from a queue in order of decreasing
// compressed size = see DESIGN.md).  order[q] = unit taken by the q-th pull (NULL: q itself).
template <bool EMIT, uint32_t LB, uint32_t DB, uint32_t NTH>
__global__ void __launch_bounds__(NTH)
foreign_decode_kernel(const uint8_t* __restrict__ in, uint64_t n, const uint64_t* __restrict__ starts, const uint64_t* __restrict__ stops,
                      uint64_t nunits, FUnitRes* __restrict__ res, const uint64_t* __restrict__ out_base, const uint64_t* __restrict__ ops_base,
                      uint16_t* __restrict__ S, uint64_t* __restrict__ ops_all, unsigned flags, const uint32_t* __restrict__ order,
                      unsigned long long* __restrict__ queue) {
    extern __shared__ __align__(16) uint8_t tp_smem[];
    __shared__ uint32_t s_ring[TB_RING * NTH];
    __shared__ uint32_t s_slow[((15 - LB) + (15 - DB)) * NTH];          // long-code search words, [slot][thread]
    uint32_t* s_lut = reinterpret_cast<uint32_t*>(tp_smem);
    uint16_t* tabs = reinterpret_cast<uint16_t*>(tp_smem + TP_LUT_WORDS * 4);
    tp_lut_init(s_lut, threadIdx.x);
    __syncthreads();
    const uint32_t NT = blockDim.x;
    uint16_t* lit = tabs + threadIdx.x;
    uint16_t* dst = tabs + threadIdx.x + (size_t)(1u << LB) * NT;
    const uint32_t lit_sa = (uint32_t)__cvta_generic_to_shared(lit);
    const uint32_t dst_sa = (uint32_t)__cvta_generic_to_shared(dst);
    const uint32_t lut_sa = (uint32_t)__cvta_generic_to_shared(s_lut);
    const uint32_t ntb = NT * 2;
    const bool strict = flags & 1u;
    const uint64_t in_bits = n * 8;
    TpTables T;
    uint32_t* slow = s_slow + threadIdx.x;
    const uint32_t slow_lit_sa = (uint32_t)__cvta_generic_to_shared(slow);
    const uint32_t slow_dst_sa = slow_lit_sa + (15 - LB) * NTH * 4;
    FdState s;
    s.br.ring_sa = (uint32_t)__cvta_generic_to_shared(s_ring + threadIdx.x);
    s.br.ring_stride = NT * 4;
    s.br.skip = (uint32_t)(reinterpret_cast<uintptr_t>(in) & 3);
    s.br.wp = reinterpret_cast<const uint32_t*>(in - s.br.skip);
    s.br.nw = (uint32_t)((s.br.skip + n + 3) >> 2);
    s.br.wi = 0; s.br.w0 = 0; s.br.bb = 0; s.br.bc = 0;
    s.pos = 0; s.nops = 0; s.bfinal = 0; s.flags = 0; s.st = ST_OK; s.btype = 0;
    s.state = TS_DONE;
    uint64_t stop_bit = ~0ull, u = 0;
    uint16_t* Su = nullptr;
    uint64_t* ops = nullptr;
    bool first_unit = false;                                  // the unit at bit 0 knows its absolute position: the too-far quirk is decided there
    bool have = false, more = true;
    // A unit that started on a false candidate decodes garbage; it normally dies within a few thousand symbols (an
    // end-of-block turns up, what follows is no header), but nothing guarantees that: it may not read more than 1 MiB
    // past its stop offset (no real producer writes blocks that long; if one did, the stream goes to the sequential decoder)
    const uint32_t wi_end = s.br.nw + 4;
    uint32_t wi_lim = wi_end;

    uint32_t waited = 0;                                      // trips since a lane began to wait for its tables (warp-uniform)
    while (__any_sync(0xFFFFFFFFu, more)) {
        const uint32_t m_tab = __ballot_sync(0xFFFFFFFFu, s.state == TS_TABLES);
        const uint32_t m_sym = __ballot_sync(0xFFFFFFFFu, s.state == TS_SYM);
        waited = m_tab ? waited + 1 : 0;
        const bool build_now = m_tab && (__popc(m_tab) >= FD_TAB_BATCH || m_sym == 0 || waited >= FD_TAB_WAIT);
        if (build_now) waited = 0;
        if (s.state == TS_DONE) {
            if (!more) continue;
            if (have) {                                       // publish the unit that has just ended
                asm volatile("cp.async.wait_all;" ::: "memory");
                if (s.st == ST_OK && tb_bitpos(s.br) > in_bits) s.st = ST_OVERRUN;
                FUnitRes r;
                r.end_bit = tb_bitpos(s.br); r.out_len = s.pos; r.nops = s.nops; r.status = s.st; r.flags = s.flags; r.pad = 0;
                res[u] = r;
                have = false;
            }
            const unsigned long long q = atomicAdd(queue, 1ull);
            if (q >= nunits) { more = false; continue; }
            u = order ? order[q] : q;
            const uint64_t start_bit = starts[u];
            stop_bit = stops[u];
            tb_seek(s.br, start_bit >> 3);
            tb_drop(s.br, (uint32_t)(start_bit & 7));
            if (EMIT) { Su = S + out_base[u]; ops = ops_all + ops_base[u]; }
            first_unit = start_bit == 0;
            const uint64_t stop_word = stop_bit == ~0ull ? (uint64_t)wi_end : ((stop_bit + (uint64_t)s.br.skip * 8) >> 5) + (1u << 18);
            wi_lim = (uint32_t)min((uint64_t)wi_end, stop_word);
            s.pos = 0; s.nops = 0; s.bfinal = 0; s.flags = 0; s.st = ST_OK;
            s.state = TS_BLOCK;
            have = true;
            continue;
        }
        if (s.state == TS_SYM) {
            TBits& br = s.br;
            if (s.pos > FD_MAX_OUT || br.wi > wi_lim) {
                s.st = br.wi > wi_end ? ST_OVERRUN : ST_FALLBACK; s.state = TS_DONE;
            } else {
                if (br.bc < 33) tb_take(br);
                uint32_t e = lds_u16(lit_sa + tb_peek(br, LB) * ntb);
                bool ok = true;
                if ((e & 15u) == 0) {
                    const int r = fd_slow_symbol<LB>(slow_lit_sa, NTH * 4, T.sorted, (uint32_t)br.bb);
                    if (r < 0) { s.st = ST_OVERRUN; s.state = TS_DONE; ok = false; }
                    else e = tp_lit_entry((uint32_t)r & 0xFFFFu, (uint32_t)r >> 16);
                }
                uint32_t l = e & 15u, p = e >> 4;
                bool is_match = ok && p >= 0x200u && !(p & 0x100u);
                if (ok && p < 256) {
                    tb_drop(br, l);
                    if (EMIT) Su[s.pos] = (uint16_t)p;
                    s.pos++;
                    #pragma unroll
                    for (uint32_t k = 0; k < TP_LIT_RUN; k++) {
                        if (k) tb_refill(br);
                        const uint32_t e2 = lds_u16(lit_sa + tb_peek(br, LB) * ntb);
                        const uint32_t l2 = e2 & 15u, p2 = e2 >> 4;
                        if (l2 == 0) break;
                        if (p2 < 256) { tb_drop(br, l2); if (EMIT) Su[s.pos] = (uint16_t)p2; s.pos++; continue; }
                        if (p2 >= 0x200u && !(p2 & 0x100u)) { l = l2; p = p2; is_match = true; }
                        break;
                    }
                }
                if (is_match) {
                    tb_drop(br, l);
                    const uint32_t lt = lds_u32(lut_sa + (p & 31u) * 4);
                    const uint32_t length = (lt & 0xFFFFu) + tb_get(br, lt >> 16);
                    tb_refill(br);
                    const uint32_t de = lds_u16(dst_sa + tb_peek(br, DB) * ntb);
                    uint32_t dl = de & 15u, dsym = de >> 4;
                    if (dl == 0) {
                        const int r = fd_slow_symbol<DB>(slow_dst_sa, NTH * 4, T.sorted + NLIT, (uint32_t)br.bb);
                        dsym = r < 0 ? 99u : ((uint32_t)r & 0xFFFFu);
                        dl = r < 0 ? 0u : ((uint32_t)r >> 16);
                    }
                    if (dsym > 29) {
                        s.st = tb_bitpos(br) + 16 > in_bits ? ST_OVERRUN : ST_DATA; s.state = TS_DONE;
                    } else {
                        tb_drop(br, dl);
                        const uint32_t dt = lds_u32(lut_sa + (32 + dsym) * 4);
                        const uint32_t dist = (dt & 0xFFFFu) + tb_get(br, dt >> 16);
                        if (first_unit && dist > s.pos) {
                            if (strict) { s.st = ST_DATA; s.state = TS_DONE; }
                            // else the reference copies nothing (inflate.hpp:268-270)
                        } else {
                            if (EMIT) ops[s.nops] = tp_op(s.pos, length, dist);
                            s.nops++;
                            s.pos += length;
                        }
                    }
                } else if (ok && p >= 256) {
                    if (p == 0x100u) {                                           // end of block
                        tb_drop(br, l);
                        if (br.wi > wi_end || tb_bitpos(br) > in_bits) { s.st = ST_OVERRUN; s.state = TS_DONE; }
                        else if (s.bfinal) { s.flags |= FU_FINAL; s.state = TS_DONE; }
                        else if (tb_bitpos(br) >= stop_bit) { s.flags |= FU_REACHED; s.state = TS_DONE; }
                        else s.state = TS_BLOCK;
                    } else { s.st = ST_DATA; s.state = TS_DONE; }                // 286 / 287
                }
            }
        } else if (s.state == TS_BLOCK) {
            s = fd_block<EMIT>(s, n, in_bits, strict, stop_bit, ops);
        } else if (s.state == TS_TABLES && build_now) {
            s = fd_tables<LB, DB>(s, lit, dst, slow, NT, T);
        }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
}

// ---- F4: ops inside the symbol image, one warp per unit ----------------------------------------------------------
// err: set to 1 when something that must not happen happened (the stream then goes to the sequential decoder)
//
// 32 ops per step, as in pass B of the chunk path (tp_apply_ops): an op is READY when every symbol it reads is final
// before the step starts -- its source ends at or below the step's first destination (a source below the unit's first
// symbol is a window marker and depends on nothing), or it is the step's first op, or its source begins at or after the
// end of the op before it (then it reads only literals, which the emit pass wrote).  Ready ops of at most 16 symbols that
// do not overlap themselves are copied one per LANE, all at once: one memory round trip for the step instead of one per
// op (the first version applied the ops one after the other, every one a load -> store round trip through the L2: 9.1 ms
// per GiB of output).  The others go through the cooperative copy in order.
constexpr uint32_t FC_FREE_MAX = 16;
__global__ void __launch_bounds__(INF_THREADS)
foreign_copy_kernel(const uint8_t* __restrict__ in, const FUnitRes* __restrict__ res, const uint64_t* __restrict__ out_base,
                    const uint64_t* __restrict__ ops_base, uint64_t nunits, uint16_t* __restrict__ S, const uint64_t* __restrict__ ops_all,
                    unsigned long long* __restrict__ counter, unsigned int* __restrict__ err) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t FULL = 0xFFFFFFFFu;
    for (;;) {
        unsigned long long u = 0;
        if (lane == 0) u = atomicAdd(counter, 1ull);
        u = __shfl_sync(FULL, u, 0);
        if (u >= nunits) break;
        const uint32_t nops = res[u].nops;
        uint16_t* Su = S + out_base[u];
        const uint64_t* ops = ops_all + ops_base[u];
        uint64_t o_next = lane < nops ? ops[lane] : 0ull;
        uint32_t carry = 0;                                                      // 1: slot 0 of this step is a stored op's source slot
        for (uint32_t b = 0; b < nops; b += 32) {
            const uint64_t o = o_next;
            o_next = 0ull;
            if (b + 32 < nops) o_next = b + 32 + lane < nops ? ops[b + 32 + lane] : 0ull;
            const uint32_t cnt = min(32u, nops - b);
            const uint32_t pos = (uint32_t)o, len = (uint32_t)(o >> 32) & 0xFFFFu, dist = (uint32_t)(o >> 48);
            const bool valid = lane < cnt;
            // stored pairs: a header slot (dist 0) is followed by a raw source-offset slot whose bits mean nothing
            uint32_t data = carry, hdrs = 0;
            carry = 0;
            for (uint32_t h = __ballot_sync(FULL, valid && dist == 0); h;) {
                const uint32_t j = __ffs(h) - 1;
                h &= h - 1;
                if ((data >> j) & 1u) continue;
                hdrs |= 1u << j;
                if (j < 31) { data |= 2u << j; h &= ~(2u << j); } else carry = 1;
            }
            const bool is_data = (data >> lane) & 1u;
            const bool is_match = valid && !is_data && dist != 0;
            const uint32_t matches = __ballot_sync(FULL, is_match);
            const uint32_t live = matches | hdrs;
            if (!live) continue;
            const uint32_t first = __ffs(live) - 1;
            const int64_t p0 = (int64_t)__shfl_sync(FULL, pos, first);
            const uint32_t prev_end = __shfl_up_sync(FULL, pos + len, 1);
            const bool prev_live = lane != 0 && ((live >> (lane - 1)) & 1u);
            const int64_t a = (int64_t)pos - (int64_t)dist;                      // the op reads symbols [a, a + len) of the unit (a < 0: window)
            const bool local = is_match && len <= FC_FREE_MAX && dist >= len;
            const bool ready = local && (a + (int64_t)len <= p0 || lane == first || (prev_live && a >= (int64_t)prev_end));
            if (__any_sync(FULL, ready)) {
                uint16_t v[FC_FREE_MAX];
                #pragma unroll
                for (uint32_t i = 0; i < FC_FREE_MAX; i++) {
                    v[i] = 0;
                    if (ready && i < len) {
                        const int64_t sidx = a + i;
                        v[i] = sidx < 0 ? (uint16_t)(F_MARK | (uint32_t)(sidx + F_WINDOW)) : Su[sidx];
                    }
                }
                #pragma unroll
                for (uint32_t i = 0; i < FC_FREE_MAX; i++)
                    if (ready && i < len) Su[pos + i] = v[i];
            }
            __syncwarp();
            uint32_t seq = live & ~__ballot_sync(FULL, ready);
            while (seq) {
                const uint32_t j = __ffs(seq) - 1;
                seq &= seq - 1;
                const uint32_t jpos = __shfl_sync(FULL, pos, j), jlen = __shfl_sync(FULL, len, j), jdist = __shfl_sync(FULL, dist, j);
                if (jdist == 0) {
                    // stored block: the next slot is the source byte offset in the stream (it may sit in the next step)
                    const uint64_t in_step = __shfl_sync(FULL, o, (j + 1) & 31);
                    const uint64_t in_next = __shfl_sync(FULL, o_next, 0);
                    const uint8_t* sp = in + (j + 1 < 32 ? in_step : in_next);
                    for (uint32_t i = lane; i < jlen; i += 32) Su[jpos + i] = sp[i];
                } else if (jdist >= jlen || jdist >= 32) {
                    // back-reference: sources below the unit's first symbol are window markers
                    for (uint32_t c0 = 0; c0 < jlen; c0 += 32) {
                        const uint32_t i = c0 + lane;
                        if (i < jlen) {
                            const int64_t sidx = (int64_t)jpos - jdist + i;
                            Su[jpos + i] = sidx < 0 ? (uint16_t)(F_MARK | (uint32_t)(sidx + F_WINDOW)) : Su[sidx];
                        }
                        if (jdist < jlen) __syncwarp();                          // the next 32 may read what these wrote
                    }
                } else {
                    // period < 32 and overlapping: every output symbol is one of the `dist` symbols below pos
                    uint32_t v = 0;
                    if (lane < jdist) {
                        const int64_t sidx = (int64_t)jpos - jdist + lane;
                        v = sidx < 0 ? (F_MARK | (uint32_t)(sidx + F_WINDOW)) : Su[sidx];
                    }
                    uint32_t r = lane % jdist;
                    const uint32_t step = 32 % jdist;
                    for (uint32_t c0 = 0; c0 < jlen; c0 += 32) {                 // all lanes take part in the shuffle
                        const uint32_t x = __shfl_sync(FULL, v, r);
                        if (c0 + lane < jlen) Su[jpos + c0 + lane] = (uint16_t)x;
                        r += step;
                        if (r >= jdist) r -= jdist;
                    }
                }
                __syncwarp();
            }
        }
        if (lane == 0 && res[u].status != ST_OK) atomicExch(err, 1u);
    }
}

// ---- F5: window propagation -----------------------------------------------------------------------------------------
// Units [g * G, (g + 1) * G) of group g are walked in order by one CTA with the 32 KiB before the current unit in a ring
// of u16 symbols in shared memory (ring slot of absolute output position a: a & 32767).
//   MODE 0 (compose): the ring starts as identity markers for the 32 KiB before the GROUP; at the end the ring is the
//                     group's map: gmap[g][i] = symbol of absolute position (group end - 32768 + i) as a literal or a
//                     marker into the window before the group.
//   MODE 1 (apply):   the ring starts as the real window before the group (gwin[g], bytes; group 0 has none), every
//                     unit's resolved tail is written to out[].
constexpr uint32_t FW_THREADS = 1024;
constexpr uint32_t FW_PER = F_WINDOW / FW_THREADS;      // 32 symbols per thread per unit
constexpr uint32_t FW_SMEM_BYTES = F_WINDOW * 2;
template <int MODE>
__global__ void __launch_bounds__(FW_THREADS)
foreign_window_kernel(const uint16_t* __restrict__ S, const uint64_t* __restrict__ out_base, uint64_t nunits, uint64_t total,
                      uint32_t G, uint16_t* __restrict__ gmap, const uint8_t* __restrict__ gwin, uint8_t* __restrict__ out,
                      unsigned int* __restrict__ err) {
    extern __shared__ __align__(16) uint8_t fw_smem[];
    uint16_t* ring = reinterpret_cast<uint16_t*>(fw_smem);
    const uint32_t tid = threadIdx.x;
    const uint64_t g = blockIdx.x;
    const uint64_t u0 = g * G, u1 = min(nunits, u0 + G);
    if (u0 >= nunits) return;
    const uint64_t gstart = out_base[u0];
    for (uint32_t i = tid; i < F_WINDOW; i += FW_THREADS) {
        // ring slot of absolute position gstart - 32768 + i
        const uint32_t slot = (uint32_t)((gstart + i) & (F_WINDOW - 1));
        if (MODE == 0) ring[slot] = (uint16_t)(F_MARK | i);
        else ring[slot] = (g == 0) ? (uint16_t)(F_MARK | i) : (uint16_t)gwin[g * F_WINDOW + i];
    }
    __syncthreads();
    bool bad = false;
    for (uint64_t u = u0; u < u1; u++) {
        const uint64_t start = out_base[u];
        const uint64_t end = u + 1 < nunits ? out_base[u + 1] : total;
        const uint64_t len = end - start;
        const uint32_t T = (uint32_t)min(len, (uint64_t)F_WINDOW);
        const uint64_t t0 = end - T;                                   // first tail position
        uint16_t v[FW_PER];
        #pragma unroll
        for (uint32_t k = 0; k < FW_PER; k++) {
            const uint32_t j = tid + k * FW_THREADS;
            uint16_t s = 0;
            if (j < T) {
                s = S[t0 + j];
                if (s & F_MARK) s = ring[(uint32_t)((start + (s & (F_MARK - 1))) & (F_WINDOW - 1))];   // position start - 32768 + idx
            }
            v[k] = s;
        }
        __syncthreads();                                               // every read of the old window is done
        #pragma unroll
        for (uint32_t k = 0; k < FW_PER; k++) {
            const uint32_t j = tid + k * FW_THREADS;
            if (j < T) {
                ring[(uint32_t)((t0 + j) & (F_WINDOW - 1))] = v[k];
                if (MODE == 1) {
                    if (v[k] & F_MARK) bad = true;                     // reaches before the start of the stream
                    out[t0 + j] = (uint8_t)v[k];
                }
            }
        }
        __syncthreads();
    }
    if (MODE == 0) {
        const uint64_t gend = u1 < nunits ? out_base[u1] : total;
        for (uint32_t i = tid; i < F_WINDOW; i += FW_THREADS) {
            const uint16_t s = ring[(uint32_t)((gend + i) & (F_WINDOW - 1))];
            // group 0 has nothing before it: a marker at a position that exists reaches before the start of the stream
            if (g == 0 && (s & F_MARK) && gend + i >= F_WINDOW) bad = true;
            gmap[g * F_WINDOW + i] = s;
        }
    }
    if (bad) atomicExch(err, 1u);
}

// chain the group maps: gwin[g + 1] = gmap[g] applied to gwin[g] (one CTA, sequential over groups)
__global__ void __launch_bounds__(FW_THREADS)
foreign_chain_kernel(const uint16_t* __restrict__ gmap, uint8_t* __restrict__ gwin, uint64_t ngroups, unsigned int* __restrict__ err) {
    extern __shared__ __align__(16) uint8_t fw_smem[];
    uint8_t (*win)[F_WINDOW] = reinterpret_cast<uint8_t (*)[F_WINDOW]>(fw_smem);
    const uint32_t tid = threadIdx.x;
    for (uint32_t i = tid; i < F_WINDOW; i += FW_THREADS) win[0][i] = 0;
    __syncthreads();
    for (uint64_t g = 0; g + 1 < ngroups; g++) {
        const uint32_t a = g & 1, b = a ^ 1;
        for (uint32_t i = tid; i < F_WINDOW; i += FW_THREADS) {
            const uint16_t s = gmap[g * F_WINDOW + i];
            const uint8_t x = (s & F_MARK) ? win[a][s & (F_MARK - 1)] : (uint8_t)s;
            win[b][i] = x;
            gwin[(g + 1) * F_WINDOW + i] = x;
        }
        __syncthreads();
    }
    (void)err;
}

// ---- F6: everything that is not a tail -----------------------------------------------------------------------------------
// grid = units; the CTA walks its unit's positions [start, end - 32768)
constexpr uint32_t FR_THREADS = 256;
constexpr uint32_t FR_SLAB = 32768;                   // positions per CTA
struct FSlab { uint64_t unit; uint64_t lo; };         // slab list built on the host: unit index, first position
__global__ void __launch_bounds__(FR_THREADS)
foreign_resolve_kernel(const uint16_t* __restrict__ S, const uint64_t* __restrict__ out_base, uint64_t nunits, uint64_t total,
                       const FSlab* __restrict__ slabs, uint8_t* __restrict__ out, unsigned int* __restrict__ err) {
    const FSlab sl = slabs[blockIdx.x];
    const uint64_t start = out_base[sl.unit];
    const uint64_t end = sl.unit + 1 < nunits ? out_base[sl.unit + 1] : total;
    const uint64_t body_end = end - start > F_WINDOW ? end - F_WINDOW : start;        // tails were written by the window pass
    const uint64_t hi = min(body_end, sl.lo + FR_SLAB);
    bool bad = false;
    for (uint64_t p = sl.lo + threadIdx.x; p < hi; p += FR_THREADS) {
        const uint16_t s = S[p];
        uint8_t x = (uint8_t)s;
        if (s & F_MARK) {
            const uint64_t idx = s & (F_MARK - 1);
            if (start + idx < F_WINDOW) bad = true;                                    // before the start of the stream
            else x = out[start - F_WINDOW + idx];
        }
        out[p] = x;
    }
    if (bad) atomicExch(err, 1u);
}

}  // namespace b200
