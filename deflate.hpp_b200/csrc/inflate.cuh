// inflate.cuh -- K6: table-driven INFLATE, one warp per independent stream or chunk.
//
// Replaces the reference's realDecompress (include/inflate.hpp:277-322), decompressHuffmanBlock
// (:226-275: one readBits(1) + a tree walk from the root per BIT), decodeTree/readCodeLengthTree/
// readDynamicTreeCodes (:136-224) and Bitwrapper (:29-134).  Not a port:
//   * all 32 lanes run the bit-serial symbol decode redundantly (warp-uniform control flow, table
//     reads are shared-memory broadcasts), so every lane knows every symbol without shuffles;
//   * symbols come from a 10-bit (lit/len) / 8-bit (dist) direct lookup table built per block by the
//     whole warp, with a canonical first-code search for the rare longer codes;
//   * literals are gathered one per lane and flushed as 32-byte coalesced stores, back-references
//     are copied by all lanes at once (period-aware when distance < 32);
//   * input reaches the bit buffer through a 1 KiB per-warp shared-memory ring filled with 16-byte
//     coalesced loads.
// Behaviour on malformed input follows the reference unless B200_F_STRICT is set (see
// include/b200_deflate.h): BTYPE 3 is skipped, NLEN is not verified, a distance that reaches before
// the start of the output copies nothing.
#pragma once
#include "common.cuh"

namespace b200 {

constexpr uint32_t INF_WARPS = 4;
constexpr uint32_t INF_THREADS = INF_WARPS * 32;
constexpr uint32_t LIT_BITS = 10;
constexpr uint32_t DST_BITS = 8;
constexpr uint32_t RING_WORDS = 256;        // 2 halves x 128 words

// status codes mirror include/b200_deflate.h
constexpr int ST_OK = 0, ST_OVERRUN = 1, ST_DATA = 2;
// extra (internal) result of a chunk decode
constexpr uint32_t END_FINAL = 1;           // stopped after a BFINAL block
constexpr uint32_t END_SYNC = 2;            // stopped after an empty stored block (chunk mode)
constexpr uint32_t END_NEEDS_HISTORY = 4;   // a distance reached before this chunk's first byte
constexpr uint32_t END_TOO_BIG = 8;         // chunk mode: produced more than the chunk capacity

// lit/len entry (u32):
//   literal(s)  : bit 31 = 0; [4:0] bits to consume (one or two codes), [15:8] first byte, [23:16] second
//                 byte, [27:24] length of the first code alone, bit 30 = entry carries two literals
//   non-literal : bit 31 = 1; [4:0] code length, [7:5] kind (1 length, 2 EOB, 3 long-or-invalid),
//                 length: [16:8] base, [23:20] extra-bit count
// dist entry   : [3:0] code length, [5:4] kind (0 ok, 3 long-or-invalid), [11:8] extra count, [31:16] base
constexpr uint32_t E_NONLIT = 0x80000000u, E_TWO = 0x40000000u;
constexpr uint32_t K_LEN = 1, K_EOB = 2, K_LONG = 3;

__constant__ uint16_t C_LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31,
                                        35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__constant__ uint16_t C_DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193,
                                         257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145,
                                         8193, 12289, 16385, 24577};
__constant__ uint8_t C_PRECODE_ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

struct __align__(16) InfWarp {
    uint32_t lit[1u << LIT_BITS];
    uint32_t dst[1u << DST_BITS];
    uint32_t ring[RING_WORDS];
    uint16_t sorted[NSYM];          // symbols ordered by (length, value): lit/len [0,288), dist [288,320)
    uint16_t first[2][16];          // first canonical code of each length
    uint16_t count[2][16];
    uint16_t offs[2][16];           // index into sorted[] of each length's first symbol
    uint16_t next[2][16];
    uint16_t pre[128];              // precode table: [3:0] len, [12:8] symbol
    uint8_t lens[NSYM];
};

constexpr uint32_t LIT_INVALID = E_NONLIT | (K_LONG << 5);

__device__ __forceinline__ uint32_t lit_entry(uint32_t sym, uint32_t len) {
    if (sym < 256) return len | (sym << 8) | (len << 24);
    if (sym == 256) return E_NONLIT | len | (K_EOB << 5);
    if (sym > 285) return LIT_INVALID;     // 286/287 exist in the fixed code but never in data
    const uint32_t idx = sym - 257;
    return E_NONLIT | len | (K_LEN << 5) | ((uint32_t)C_LEN_BASE[idx] << 8) | (len_extra_bits(idx) << 20);
}
__device__ __forceinline__ uint32_t dst_entry(uint32_t sym, uint32_t len) {
    if (sym > 29) return K_LONG << 4;
    return len | (dist_extra_bits(sym) << 8) | ((uint32_t)C_DIST_BASE[sym] << 16);
}

// ---- bit reader (warp-uniform state; every lane holds the same copy) -----------------------------
struct BitReader {
    const uint8_t* base16;   // 16-byte aligned address at or below the stream start
    uint64_t limit;          // stream end, in bytes from base16
    uint64_t bb;             // bit buffer (LSB first)
    uint32_t bc;             // valid bits in bb
    uint32_t wi;             // next ring word to pull (absolute word index from base16)
    uint32_t skip;           // bytes between base16 and the stream start
};

// Load ring half `k & 1` with words [k*128, (k+1)*128); zeros past the end of the stream.
__device__ __forceinline__ void ring_fill(InfWarp* S, const BitReader& br, uint32_t k, uint32_t lane) {
    const uint64_t byte0 = (uint64_t)k * 512 + lane * 16;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (byte0 < br.limit) v = *reinterpret_cast<const uint4*>(br.base16 + byte0);
    __syncwarp();   // every lane is done reading the half that is about to be overwritten
    *reinterpret_cast<uint4*>(&S->ring[(k & 1) * 128 + lane * 4]) = v;
    __syncwarp();
}

__device__ __forceinline__ void br_seek(InfWarp* S, BitReader& br, uint64_t stream_byte, uint32_t lane) {
    const uint64_t b = br.skip + stream_byte;
    br.wi = (uint32_t)(b >> 2);
    const uint32_t k = br.wi >> 7;
    __syncwarp();
    ring_fill(S, br, k, lane);
    ring_fill(S, br, k + 1, lane);
    const uint32_t sh = (uint32_t)(b & 3) * 8;
    br.bb = (uint64_t)(S->ring[br.wi & (RING_WORDS - 1)] >> sh);
    br.bc = 32 - sh;
    br.wi++;
    if ((br.wi & 127) == 0) ring_fill(S, br, (br.wi >> 7) + 1, lane);
}

// make sure at least 32 bits are buffered
__device__ __forceinline__ void br_need32(InfWarp* S, BitReader& br, uint32_t lane) {
    if (br.bc < 32) {
        br.bb |= (uint64_t)S->ring[br.wi & (RING_WORDS - 1)] << br.bc;
        br.bc += 32;
        br.wi++;
        if ((br.wi & 127) == 0) ring_fill(S, br, (br.wi >> 7) + 1, lane);
    }
}
__device__ __forceinline__ uint32_t br_peek(const BitReader& br, uint32_t n) { return (uint32_t)br.bb & ((1u << n) - 1u); }
__device__ __forceinline__ void br_drop(BitReader& br, uint32_t n) { br.bb >>= n; br.bc -= n; }
__device__ __forceinline__ uint32_t br_get(InfWarp* S, BitReader& br, uint32_t n, uint32_t lane) {   // n <= 16
    br_need32(S, br, lane);
    uint32_t v = br_peek(br, n);
    br_drop(br, n);
    return v;
}
// bits consumed so far, relative to the stream start
__device__ __forceinline__ uint64_t br_bitpos(const BitReader& br) {
    return (uint64_t)br.wi * 32 - br.bc - (uint64_t)br.skip * 8;
}

// ---- table construction (whole warp) --------------------------------------------------------------
// which = 0: lit/len alphabet lens[0..n) -> S->lit;  which = 1: distance alphabet -> S->dst.
// Returns false on an over-subscribed code.
__device__ bool build_table(InfWarp* S, const uint8_t* lens, uint32_t n, uint32_t which, uint32_t lane) {
    const uint32_t FULL = 0xFFFFFFFFu;
    uint16_t* count = S->count[which];
    uint16_t* first = S->first[which];
    uint16_t* offs = S->offs[which];
    uint16_t* next = S->next[which];
    uint16_t* sorted = S->sorted + (which ? NLIT : 0);
    uint32_t* table = which ? S->dst : S->lit;
    const uint32_t tbits = which ? DST_BITS : LIT_BITS;

    // per-length counts via ballots (16 lengths x ceil(n/32) batches is small)
    uint32_t mycount = 0;   // lane l (1..15) accumulates count[l]
    for (uint32_t b = 0; b < n; b += 32) {
        const uint32_t i = b + lane;
        const uint32_t l = i < n ? lens[i] : 0;
        for (uint32_t q = 1; q <= 15; q++) {
            const uint32_t m = __ballot_sync(FULL, l == q);
            if (lane == q) mycount += __popc(m);
        }
    }
    // exclusive scans across lanes 1..15: canonical first code and sorted[] offset
    int left = 1;
    uint32_t code = 0, off = 0, myfirst = 0, myoff = 0;
    bool over = false;
    for (uint32_t q = 1; q <= 15; q++) {
        const uint32_t c = __shfl_sync(FULL, mycount, q);
        code <<= 1;
        if (lane == q) { myfirst = code; myoff = off; }
        code += c; off += c;
        left = (left << 1) - (int)c;
        if (left < 0) over = true;
    }
    if (over) return false;
    if (lane < 16) { count[lane] = (uint16_t)(lane ? mycount : 0); first[lane] = (uint16_t)myfirst; offs[lane] = (uint16_t)myoff; next[lane] = 0; }
    for (uint32_t i = lane; i < (1u << tbits); i += 32) table[i] = which ? (K_LONG << 4) : LIT_INVALID;
    __syncwarp();

    for (uint32_t b = 0; b < n; b += 32) {
        const uint32_t i = b + lane;
        const uint32_t l = i < n ? lens[i] : 0;
        const uint32_t peers = __match_any_sync(FULL, l);
        const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
        uint32_t r = 0;
        if (l) r = next[l] + rank;
        __syncwarp();
        if (l && rank == 0) next[l] += (uint16_t)__popc(peers);
        __syncwarp();
        if (l) {
            sorted[offs[l] + r] = (uint16_t)i;
            if (l <= tbits) {
                const uint32_t rev = bitrev(first[l] + r, l);
                const uint32_t e = which ? dst_entry(i, l) : lit_entry(i, l);
                for (uint32_t k = rev; k < (1u << tbits); k += (1u << l)) table[k] = e;
            }
        }
    }
    __syncwarp();
    if (which == 0) {
        // second pass: where a literal's code leaves room in the index for another complete literal code,
        // let one lookup deliver both.  Entry j = i >> len1 holds the symbol whose code starts right
        // after the first one; it is usable iff its own code fits in the remaining index bits.  The
        // single-literal fields ([15:8], [27:24]) survive the rewrite, so in-place is race-free.
        for (uint32_t i = lane; i < (1u << LIT_BITS); i += 32) {
            const uint32_t e = table[i];
            if (!(e & E_NONLIT)) {
                const uint32_t l1 = (e >> 24) & 15u;
                const uint32_t e2 = table[i >> l1];
                const uint32_t l2 = (e2 >> 24) & 15u;
                if (!(e2 & E_NONLIT) && l1 + l2 <= LIT_BITS)
                    table[i] = (l1 + l2) | (e & 0x0F00FF00u) | (((e2 >> 8) & 0xFFu) << 16) | E_TWO;
            }
        }
        __syncwarp();
    }
    return true;
}

// Canonical search for codes the direct table does not resolve (longer than the table index, or not
// a code at all).  Returns the symbol and its length, or -1.
__device__ __forceinline__ int slow_symbol(const InfWarp* S, uint32_t which, uint32_t bits32, uint32_t& len_out) {
    const uint32_t rb = __brev(bits32);
    for (uint32_t l = 1; l <= 15; l++) {
        const uint32_t code = rb >> (32 - l);
        const uint32_t rel = code - S->first[which][l];
        if (code >= S->first[which][l] && rel < S->count[which][l]) {
            len_out = l;
            return S->sorted[(which ? NLIT : 0) + S->offs[which][l] + rel];
        }
    }
    return -1;
}

// ---- dynamic block header ------------------------------------------------------------------------
__device__ int read_dynamic_header(InfWarp* S, BitReader& br, uint32_t lane) {
    const uint32_t hlit = br_get(S, br, 5, lane) + 257;
    const uint32_t hdist = br_get(S, br, 5, lane) + 1;
    const uint32_t hclen = br_get(S, br, 4, lane) + 4;
    // precode lengths (3 bits each) -> 7-bit direct table, built redundantly per lane pair-free:
    uint32_t mylen = 0;        // lane s (< 19) holds the length of precode symbol s
    for (uint32_t i = 0; i < hclen; i++) {
        const uint32_t v = br_get(S, br, 3, lane);
        if (lane == C_PRECODE_ORDER[i]) mylen = v;
    }
    const uint32_t FULL = 0xFFFFFFFFu;
    {
        uint32_t code = 0;
        int left = 1;
        uint32_t mycode = 0;
        for (uint32_t q = 1; q <= 7; q++) {
            const uint32_t m = __ballot_sync(FULL, mylen == q);
            code <<= 1;
            if (mylen == q) mycode = code + __popc(m & ((1u << lane) - 1u));
            code += __popc(m);
            left = (left << 1) - (int)__popc(m);
        }
        if (left < 0) return ST_DATA;
        for (uint32_t i = lane; i < 128; i += 32) S->pre[i] = 0;
        __syncwarp();
        if (mylen) {
            const uint32_t rev = bitrev(mycode, mylen);
            for (uint32_t k = rev; k < 128; k += (1u << mylen)) S->pre[k] = (uint16_t)(mylen | (lane << 8));
        }
        __syncwarp();
    }
    // code lengths: lit/len list then distance list, one run-length state across both (RFC 1951;
    // the reference parses them with two independent calls, inflate.hpp:216-220)
    const uint32_t total = hlit + hdist;
    uint32_t i = 0, prev = 0;
    while (i < total) {
        br_need32(S, br, lane);
        const uint32_t e = S->pre[br_peek(br, 7)];
        const uint32_t l = e & 15u, sym = e >> 8;
        if (l == 0) return ST_OVERRUN;   // no code matches: the reference reads on until it overruns (inflate.hpp:171-175)
        br_drop(br, l);
        uint32_t rep = 1, val = sym;
        if (sym == 16) { rep = 3 + br_peek(br, 2); br_drop(br, 2); val = prev; }   // prev starts at 0 (inflate.hpp:170)
        else if (sym == 17) { rep = 3 + br_peek(br, 3); br_drop(br, 3); val = 0; }
        else if (sym == 18) { rep = 11 + br_peek(br, 7); br_drop(br, 7); val = 0; }
        if (i + rep > total) return ST_DATA;
        // lanes write the run in parallel; position j of the combined list maps to lens[]
        for (uint32_t j = lane; j < rep; j += 32) {
            const uint32_t p = i + j;
            S->lens[p < hlit ? p : NLIT + (p - hlit)] = (uint8_t)val;
        }
        i += rep;
        prev = val;
    }
    __syncwarp();
    if (S->lens[256] == 0) return ST_DATA;   // no end-of-block code: the block could never terminate
    if (!build_table(S, S->lens, hlit, 0, lane)) return ST_DATA;
    if (!build_table(S, S->lens + NLIT, hdist, 1, lane)) return ST_DATA;
    return ST_OK;
}

__device__ void fixed_tables(InfWarp* S, uint32_t lane) {
    for (uint32_t i = lane; i < NLIT; i += 32) S->lens[i] = (uint8_t)fixed_lit_len(i);
    for (uint32_t i = lane; i < NDIST; i += 32) S->lens[NLIT + i] = 5;
    __syncwarp();
    build_table(S, S->lens, NLIT, 0, lane);
    build_table(S, S->lens + NLIT, NDIST, 1, lane);
}

// ---- one stream ------------------------------------------------------------------------------------
// s_mod[d][c] = c % d for d in 1..31, c in 0..32 (period-aware copies without integer division)
struct ModLut { uint8_t m[32][36]; uint64_t nib[9]; };   // nib[d]: nibble i = i % d (d = 1..8)
__device__ __forceinline__ void modlut_init(ModLut* L, uint32_t tid, uint32_t nthreads) {
    for (uint32_t i = tid; i < 32 * 33; i += nthreads) {
        const uint32_t d = i / 33, c = i % 33;
        L->m[d][c] = (uint8_t)(d ? c % d : 0);
    }
    if (tid < 9) {
        uint64_t v = 0;
        for (uint32_t i = 0; i < 16; i++) v |= (uint64_t)(tid ? i % tid : 0) << (4 * i);
        L->nib[tid] = v;
    }
}

// Decodes blocks until BFINAL (or, if stop_at_sync, until an empty stored block).  Warp-uniform: all
// 32 lanes execute the same symbol decode; lanes differ only in which output bytes they store.
// `cap` bounds what may be written or read
// back; decoding continues past it so the full size is still reported (the reference truncates the
// same way, inflate.hpp:345).  max_out: stop (END_TOO_BIG) once more than this was produced.
__device__ int inflate_warp(InfWarp* S, const ModLut* ML, const uint8_t* in, uint64_t in_len, uint8_t* out,
                            uint64_t cap64, bool stop_at_sync, uint64_t max_out64, unsigned flags, uint32_t lane,
                            uint64_t& out_len, uint64_t& in_used, uint32_t& end_flags) {
    BitReader br;
    br.skip = (uint32_t)(reinterpret_cast<uintptr_t>(in) & 15);
    br.base16 = in - br.skip;
    br.limit = br.skip + in_len;
    br_seek(S, br, 0, lane);
    // Output positions are 64-bit end to end: `op` counts bytes since the last REBASE (the hot loop keeps 32-bit
    // arithmetic); once op passes 2 GiB the output pointers move up by 2 GiB and `base` remembers it.  Between two
    // rebase checks (every bit-buffer refill, every stored block) op grows by at most 32 symbols x 258 bytes or one
    // stored block of 65535, so it cannot wrap.
    constexpr uint32_t REBASE = 0x80000000u;
    uint64_t base = 0;      // bytes rebased away so far
    uint32_t cap = (uint32_t)min(cap64, (uint64_t)0xFFFFFFF0u);
    uint32_t max_out = (uint32_t)min(max_out64, (uint64_t)0xFFFFFF00u);
    uint32_t op = 0;        // bytes produced since `base`
    const bool strict = flags & 1u;
    const uint64_t in_bits = in_len * 8;
    const uint32_t wi_limit = (uint32_t)(br.limit >> 2) + 4;
    // Literals are stored straight away: lane 0 writes an entry's first literal, lane 1 its second
    // (entries carry one or two).  No pending state, nothing to flush before a back-reference.
    // The per-lane constants are laundered through an empty asm so that the compiler keeps them in
    // registers instead of re-deriving them from kernel parameters on every symbol (the integer ALU
    // pipe is the bottleneck of this loop -- see profiles/).
    uint8_t* outl = out + lane;                                  // this lane's byte column
    uint8_t* outb = out;
    uint32_t lit_shift = lane ? 16 : 8;
    uint32_t capl = cap > lane ? cap - lane : 0;                 // op < capl  <=>  op + lane < cap
    asm volatile("" : "+l"(outl), "+l"(outb), "+r"(lit_shift), "+r"(capl));
    auto rebase = [&]() {
        outl += REBASE; outb += REBASE; out += REBASE; op -= REBASE; base += REBASE;
        cap = (uint32_t)min(cap64 > base ? cap64 - base : (uint64_t)0, (uint64_t)0xFFFFFFF0u);
        max_out = (uint32_t)min(max_out64 > base ? max_out64 - base : (uint64_t)0, (uint64_t)0xFFFFFF00u);
        capl = cap > lane ? cap - lane : 0;
    };
    int st = ST_OK;
    uint32_t empty_run = 0;
    end_flags = 0;

    for (;;) {
        if (br_bitpos(br) + 3 > in_bits) { st = ST_OVERRUN; break; }
        const uint32_t hdr = br_get(S, br, 3, lane);
        const uint32_t bfinal = hdr & 1, btype = hdr >> 1;
        if (btype == 0) {
            const uint32_t padbits = br_peek(br, br.bc & 7);   // only an all-zero padding counts towards a separator
            br_drop(br, br.bc & 7);                       // to the byte boundary (bc == 0 mod 8 <=> aligned)
            br_need32(S, br, lane);
            const uint32_t len = br_peek(br, 16);
            br_drop(br, 16);
            const uint32_t nlen = br_peek(br, 16);
            br_drop(br, 16);
            if (strict && (len ^ nlen) != 0xFFFFu) { st = ST_DATA; break; }
            const uint64_t bpos = br_bitpos(br) >> 3;      // byte-aligned here
            if (bpos + len > in_len) { st = ST_OVERRUN; break; }
            __syncwarp();
            const uint8_t* sp = in + bpos;
            const uint32_t ncopy = op < cap ? min(len, cap - op) : 0;
            uint8_t* dp = out + op;
            if (((reinterpret_cast<uintptr_t>(sp) ^ reinterpret_cast<uintptr_t>(dp)) & 3) == 0) {
                // source and destination agree modulo 4: bytes up to alignment, then one u32 per lane
                const uint32_t head = min(ncopy, (uint32_t)((4 - (reinterpret_cast<uintptr_t>(dp) & 3)) & 3));
                if (lane < head) dp[lane] = sp[lane];
                const uint32_t words = (ncopy - head) >> 2;
                for (uint32_t i = lane; i < words; i += 32)
                    reinterpret_cast<uint32_t*>(dp + head)[i] = reinterpret_cast<const uint32_t*>(sp + head)[i];
                const uint32_t done = head + words * 4;
                if (done + lane < ncopy) dp[done + lane] = sp[done + lane];
            } else {
                for (uint32_t i = lane; i < ncopy; i += 32) dp[i] = sp[i];
            }
            op += len;
            if (op >= REBASE) rebase();
            __syncwarp();
            br_seek(S, br, bpos + len, lane);
            if (len == 0 && !bfinal && padbits == 0) {
                // chunk separator = two empty stored blocks in a row (common.cuh); the segment index's
                // stored blocks have a non-zero padding and never count
                if (++empty_run == 2 && stop_at_sync) { end_flags |= END_SYNC; break; }
            } else {
                empty_run = 0;
            }
        } else if (btype == 3) {
            if (strict) { st = ST_DATA; break; }           // the reference's switch has no case 3
        } else {
            empty_run = 0;
            if (btype == 1) fixed_tables(S, lane);
            else { st = read_dynamic_header(S, br, lane); if (st) break; }
            // ---- symbol loop ----
            for (;;) {
                if (br.bc < 32) {
                    br.bb |= (uint64_t)S->ring[br.wi & (RING_WORDS - 1)] << br.bc;
                    br.bc += 32;
                    br.wi++;
                    if ((br.wi & 127) == 0) ring_fill(S, br, (br.wi >> 7) + 1, lane);
                    if (br.wi > wi_limit || op > max_out) break;        // resolved after the loop
                    if (op >= REBASE) rebase();
                }
                uint32_t e = S->lit[br_peek(br, LIT_BITS)];
                if ((int32_t)e >= 0) {
                    br_drop(br, e & 31u);
                    const uint32_t two = e >> 30;                       // E_TWO is bit 30, bit 31 is clear
                    const uint8_t v = (uint8_t)(e >> lit_shift);
                    uint8_t* a = outl + op;
                    if (lane <= two && op < capl) *a = v;
                    op += 1 + two;
                    continue;
                }
                uint32_t kind = (e >> 5) & 7u;
                if (kind == K_LONG) {
                    uint32_t sl;
                    const int sym = slow_symbol(S, 0, (uint32_t)br.bb, sl);
                    // no literal/length code matches: the reference keeps extending the code until it
                    // runs off the input (inflate.hpp:231-235), i.e. it throws the overrun error
                    if (sym < 0) { st = ST_OVERRUN; break; }
                    e = lit_entry((uint32_t)sym, sl);
                    if (e == LIT_INVALID) { st = ST_DATA; break; }
                    if ((int32_t)e >= 0) {
                        br_drop(br, sl);
                        if (lane == 0 && op < cap) outb[op] = (uint8_t)(e >> 8);
                        op++;
                        continue;
                    }
                    kind = (e >> 5) & 7u;
                }
                br_drop(br, e & 31u);
                if (kind == K_EOB) break;
                const uint32_t ne = (e >> 20) & 15u;
                const uint32_t length = ((e >> 8) & 0x1FFu) + br_peek(br, ne);
                br_drop(br, ne);
                if (br.bc < 32) {
                    // the same refill AND the same limit checks as at the top of the loop: a run of back-references with
                    // short codes (e.g. the zero bits behind a truncated stream) refills only here, never up there
                    br.bb |= (uint64_t)S->ring[br.wi & (RING_WORDS - 1)] << br.bc;
                    br.bc += 32;
                    br.wi++;
                    if ((br.wi & 127) == 0) ring_fill(S, br, (br.wi >> 7) + 1, lane);
                    if (br.wi > wi_limit || op > max_out) break;        // resolved after the loop
                    if (op >= REBASE) rebase();
                }
                uint32_t de = S->dst[br_peek(br, DST_BITS)];
                uint32_t dl = de & 15u;
                if (((de >> 4) & 3u) == K_LONG) {
                    uint32_t sl;
                    const int sym = slow_symbol(S, 1, (uint32_t)br.bb, sl);
                    if (sym < 0 || sym > 29) { st = br_bitpos(br) + 16 > in_bits ? ST_OVERRUN : ST_DATA; break; }
                    de = dst_entry((uint32_t)sym, sl);
                    dl = sl;
                }
                br_drop(br, dl);
                const uint32_t dne = (de >> 8) & 15u;
                const uint32_t dist = (de >> 16) + br_peek(br, dne);
                br_drop(br, dne);
                if (dist > op && base == 0) {
                    // reaches before the first byte this warp produced
                    if (stop_at_sync) { end_flags |= END_NEEDS_HISTORY; st = ST_DATA; break; }
                    if (strict) { st = ST_DATA; break; }
                    continue;                              // reference: copies nothing (inflate.hpp:268-270)
                }
                // ---- back-reference copy, all lanes ----
                __syncwarp();                              // literal stores of lanes 0/1 are visible to every lane
                const uint32_t room = op < cap ? cap - op : 0;          // bytes that may still be written
                const uint32_t ncopy = min(length, room);
                uint8_t* dp = outb + op;
                const uint8_t* sp = dp - dist;
                if (dist >= length || dist >= 32) {
                    if (length <= 32) {
                        // the common case: one predicated load + store, nothing read that this step writes.  (Measured and
                        // rejected: deferring the store until the next back-reference so that decoding goes on under the
                        // load's L2 latency -- 19 % of the stall samples sit here -- made the batch 9 % slower.)
                        if (lane < ncopy) dp[lane] = sp[lane];
                    } else {
                        #pragma unroll 1
                        for (uint32_t b = 0; b < length; b += 32) {
                            // a 32-byte step only reads bytes stored before the step began (dist >= 32)
                            const uint32_t i = b + lane;
                            if (i < ncopy) dp[i] = sp[i];
                            if (dist < length) __syncwarp();
                        }
                    }
                } else if (dist == 1) {
                    const uint8_t v = room ? sp[0] : 0;
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
                op += length;
                // no barrier here: the next reader of these bytes is a later back-reference, and every
                // back-reference starts with the __syncwarp() above
            }
            if (st) break;
            if (br.wi > wi_limit || br_bitpos(br) > in_bits) { st = ST_OVERRUN; break; }
            if (op > max_out) { end_flags |= END_TOO_BIG; st = ST_DATA; break; }
        }
        if (op > max_out) { end_flags |= END_TOO_BIG; st = ST_DATA; break; }
        if (bfinal) { end_flags |= END_FINAL; break; }
    }
    __syncwarp();
    if (br_bitpos(br) > in_bits) st = ST_OVERRUN;   // whatever else went wrong, the reference would have thrown first
    out_len = base + op;
    in_used = (br_bitpos(br) + 7) >> 3;
    return st;
}

// ---- kernels -----------------------------------------------------------------------------------
// Both kernels are persistent: the grid is sized to the machine and every warp pulls the next stream
// (or chunk) index from a global counter, because the cost of a unit varies a lot (a stored chunk is
// a memcpy, a text chunk is ~25k symbols).  *counter must be zero at launch.
constexpr uint32_t INF_MAX_CTAS_PER_SM = 7;

// Batch: stream i = in + in_off[i] (in_len[i] bytes) -> out + out_off[i] (<= out_cap[i] bytes).
__global__ void __launch_bounds__(INF_THREADS)
inflate_batch_kernel(const uint8_t* __restrict__ in, const uint64_t* __restrict__ in_off,
                     const uint64_t* __restrict__ in_len, uint8_t* __restrict__ out,
                     const uint64_t* __restrict__ out_off, const uint64_t* __restrict__ out_cap,
                     uint64_t* __restrict__ out_len, int32_t* __restrict__ status, uint64_t n_streams,
                     unsigned flags, unsigned long long* __restrict__ counter) {
    __shared__ InfWarp S[INF_WARPS];
    __shared__ ModLut ML;
    modlut_init(&ML, threadIdx.x, INF_THREADS);
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (;;) {
        unsigned long long i = 0;
        if (lane == 0) i = atomicAdd(counter, 1ull);
        i = __shfl_sync(0xFFFFFFFFu, i, 0);
        if (i >= n_streams) break;
        uint64_t ol = 0, used = 0;
        uint32_t ef = 0;
        const int st = inflate_warp(&S[warp], &ML, in + in_off[i], in_len[i], out + out_off[i], out_cap[i], false,
                                    ~0ull, flags, lane, ol, used, ef);
        if (lane == 0) { out_len[i] = ol; status[i] = st; }
    }
}

// ---- one small stream: the output window lives in shared memory ------------------------------------
// A short stream is one warp's serial chain, and in global memory every back-reference is a store -> load round trip
// through the L2 (~600 cycles: test.bmp, 21 KB, took 0.7 ms -- slower than the reference on a CPU core).  Here the same
// decoder writes into a shared-memory image of the output (the decoder only sees a generic pointer; loads of bytes just
// stored cost ~30 cycles), and all threads of the CTA copy the image out afterwards.  cap <= SMALL_OUT_BYTES.
// result[0] = decoded length (the whole stream's, also beyond cap), result[1] = status; host_result, if given, is pinned
// mapped host memory that receives the same two words (no copy-engine round trip for 16 bytes).
constexpr uint32_t SMALL_THREADS = 256;
constexpr uint32_t SMALL_OUT_BYTES = 160u << 10;
constexpr uint32_t SMALL_IN_BYTES = 64u << 10;
constexpr unsigned long long SMALL_ST_INDEXED = 0x100;      // result[1]: not decoded, the stream carries a segment index
__global__ void __launch_bounds__(SMALL_THREADS)
inflate_small_kernel(const uint8_t* __restrict__ in, uint64_t n, uint8_t* __restrict__ out, uint32_t cap, unsigned flags,
                     unsigned long long* __restrict__ result, volatile unsigned long long* __restrict__ host_result) {
    extern __shared__ __align__(16) uint8_t sm_out[];
    __shared__ InfWarp S;
    __shared__ ModLut ML;
    __shared__ unsigned long long s_res[2];
    // a stream that opens with a segment index group (common.cuh) is this library's: the segment-parallel path decodes
    // it several times faster than one warp can -- say so and leave
    if (n >= 7 && (in[0] & 0x87u) == 0x80u && in[1] == 0 && in[2] == 0 && in[3] == 0xFF && in[4] == 0xFF) {
        if (threadIdx.x == 0) {
            result[0] = 0; result[1] = SMALL_ST_INDEXED;
            if (host_result) { host_result[0] = 0; host_result[1] = SMALL_ST_INDEXED; __threadfence_system(); }
        }
        return;
    }
    modlut_init(&ML, threadIdx.x, SMALL_THREADS);
    __syncthreads();
    if (threadIdx.x < 32) {
        uint64_t ol = 0, used = 0;
        uint32_t ef = 0;
        const int st = inflate_warp(&S, &ML, in, n, sm_out, cap, false, ~0ull, flags, threadIdx.x, ol, used, ef);
        if (threadIdx.x == 0) { s_res[0] = ol; s_res[1] = (unsigned long long)(uint32_t)st; }
    }
    __syncthreads();
    const uint32_t w = (uint32_t)(s_res[0] < cap ? s_res[0] : cap);
    if ((reinterpret_cast<uintptr_t>(out) & 15) == 0) {
        for (uint32_t i = threadIdx.x * 16; i + 16 <= w; i += SMALL_THREADS * 16)
            *reinterpret_cast<uint4*>(out + i) = *reinterpret_cast<const uint4*>(sm_out + i);
        for (uint32_t i = (w & ~15u) + threadIdx.x; i < w; i += SMALL_THREADS) out[i] = sm_out[i];
    } else {
        for (uint32_t i = threadIdx.x; i < w; i += SMALL_THREADS) out[i] = sm_out[i];
    }
    if (threadIdx.x == 0) {
        result[0] = s_res[0]; result[1] = s_res[1];
        if (host_result) { host_result[0] = s_res[0]; host_result[1] = s_res[1]; __threadfence_system(); }
    }
}

// ---- single stream, chunk-parallel ---------------------------------------------------------------
// Pass 1/2: find every chunk separator tail (00 00 FF FF 00 00 00 FF FF, see common.cuh); the byte after
// it is a candidate chunk start.  One warp scans a contiguous 16 KiB region with coalesced 16-byte loads, so
// candidates come out in stream order: pass A counts per warp, an exclusive scan gives each warp its
// slot, pass B (WRITE) stores the candidate offsets.  cand[0] = 0 is written by the host side.
constexpr uint32_t SYNC_REGION = 16384;
constexpr uint32_t SYNC_CACHE = 4;          // candidates per region the count pass remembers (a region holds one on average
                                            // at most: chunks compress to >= 20 KB or are stored); the write pass re-reads
                                            // the input only for regions with more
template <bool WRITE>
__global__ void __launch_bounds__(256)
find_sync_kernel(const uint8_t* __restrict__ in, uint64_t n, uint32_t* __restrict__ counts, uint32_t* __restrict__ cache,
                 const uint64_t* __restrict__ offsets, uint64_t* __restrict__ cand, uint64_t cand_cap, uint64_t min_cand) {
    // min_cand: candidates at or below this offset are not reported (a scan that restarts at a known chunk start)
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t lo = w * SYNC_REGION;
    if (lo >= n) return;
    const uint64_t hi = min(n, lo + SYNC_REGION);
    const uint32_t FULL = 0xFFFFFFFFu;
    const bool aligned = (reinterpret_cast<uintptr_t>(in) & 15) == 0;
    uint64_t run = WRITE ? offsets[w] : 0;
    if (WRITE) {
        const uint32_t c = counts[w];
        if (c <= SYNC_CACHE) {
            if (lane < c && run + lane < cand_cap) cand[run + lane] = lo + cache[w * SYNC_CACHE + lane] + SYNC_PATTERN_BYTES;
            return;
        }
    }
    uint32_t total = 0;
    #pragma unroll 2
    for (uint64_t b = lo; b < hi; b += 512) {
        const uint64_t p = b + lane * 16;
        uint32_t wd[7] = {0, 0, 0, 0, 0, 0, 0};
        const bool fast = aligned && p + 28 <= n;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (fast) v = *reinterpret_cast<const uint4*>(in + p);
        // the 12 bytes after a lane's 16 are the next lane's first three words: one load per lane, not four
        const uint32_t nx = __shfl_down_sync(FULL, v.x, 1), ny = __shfl_down_sync(FULL, v.y, 1), nz = __shfl_down_sync(FULL, v.z, 1);
        if (fast) {
            wd[0] = v.x; wd[1] = v.y; wd[2] = v.z; wd[3] = v.w;
            if (lane < 31 && p + 44 <= n) { wd[4] = nx; wd[5] = ny; wd[6] = nz; }
            else {
                wd[4] = *reinterpret_cast<const uint32_t*>(in + p + 16);
                wd[5] = *reinterpret_cast<const uint32_t*>(in + p + 20);
                wd[6] = *reinterpret_cast<const uint32_t*>(in + p + 24);
            }
        } else {
            for (uint32_t j = 0; j < 28; j++)
                if (p + j < n) wd[j >> 2] |= (uint32_t)in[p + j] << (8 * (j & 3));
        }
        uint32_t hits = 0;   // bit j: the 9-byte separator tail 00 00 FF FF 00 00 00 FF FF starts at byte p + j
        #pragma unroll
        for (uint32_t j = 0; j < 16; j++) {
            const uint32_t sh = (j & 3) * 8;
            const uint32_t x0 = __funnelshift_r(wd[j >> 2], wd[(j >> 2) + 1], sh);
            const uint32_t x1 = __funnelshift_r(wd[(j >> 2) + 1], wd[(j >> 2) + 2], sh);
            const uint32_t x2 = __funnelshift_r(wd[(j >> 2) + 2], wd[(j >> 2) + 3], sh);
            if (x0 == 0xFFFF0000u && x1 == 0xFF000000u && (x2 & 0xFFu) == 0xFFu && p + j + SYNC_PATTERN_BYTES < n && p + j < hi &&
                p + j + SYNC_PATTERN_BYTES > min_cand)
                hits |= 1u << j;
        }
        if (__ballot_sync(FULL, hits != 0)) {
            const uint32_t c = __popc(hits);
            uint32_t incl = c;
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t u = __shfl_up_sync(FULL, incl, o);
                if ((int)lane >= o) incl += u;
            }
            uint64_t slot = (WRITE ? run : (uint64_t)total) + (incl - c);
            uint32_t h = hits;
            while (h) {
                const uint32_t j = __ffs(h) - 1;
                h &= h - 1;
                if (WRITE) { if (slot < cand_cap) cand[slot] = p + j + SYNC_PATTERN_BYTES; }
                else if (slot < SYNC_CACHE) cache[w * SYNC_CACHE + slot] = (uint32_t)(p + j - lo);
                slot++;
            }
            const uint32_t tot = __shfl_sync(FULL, incl, 31);
            run += tot; total += tot;
        }
    }
    if (!WRITE && lane == 0) counts[w] = total;
}

// Pass 3: one warp per candidate chunk, decoded optimistically into out + i * CHUNK.
struct ChunkResult { uint64_t in_end; uint32_t out_len; uint16_t status; uint16_t end_flags; };

__global__ void __launch_bounds__(INF_THREADS)
inflate_chunks_kernel(const uint8_t* __restrict__ in, uint64_t n, const uint64_t* __restrict__ cand,
                      uint64_t ncand, uint8_t* __restrict__ out, uint64_t cap, ChunkResult* __restrict__ res,
                      unsigned flags, unsigned long long* __restrict__ counter) {
    __shared__ InfWarp S[INF_WARPS];
    __shared__ ModLut ML;
    modlut_init(&ML, threadIdx.x, INF_THREADS);
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (;;) {
        unsigned long long i = 0;
        if (lane == 0) i = atomicAdd(counter, 1ull);
        i = __shfl_sync(0xFFFFFFFFu, i, 0);
        if (i >= ncand) break;
        const uint64_t start = cand[i];
        const uint64_t o0 = i * CHUNK;
        const uint64_t mycap = o0 >= cap ? 0 : min((uint64_t)CHUNK, cap - o0);
        uint64_t ol = 0, used = 0;
        uint32_t ef = 0;
        const int st = inflate_warp(&S[warp], &ML, in + start, n - start, out + o0, mycap, true, CHUNK, flags, lane,
                                    ol, used, ef);
        if (lane == 0) {
            ChunkResult r;
            r.in_end = start + used; r.out_len = (uint32_t)ol; r.status = (uint16_t)st; r.end_flags = (uint16_t)ef;
            res[i] = r;
        }
    }
}

// Pass 4: the optimistic layout is right iff every chunk decoded cleanly, ended exactly where the next
// candidate starts, produced exactly CHUNK bytes (all but the last) and only the last one is final.
// result[0] = 1 if valid, result[1] = total decoded bytes.
__global__ void validate_chunks_kernel(const uint64_t* __restrict__ cand, uint64_t ncand,
                                       const ChunkResult* __restrict__ res, unsigned long long* __restrict__ result) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncand) return;
    const ChunkResult r = res[i];
    bool ok = r.status == ST_OK;
    if (i + 1 < ncand) ok = ok && r.in_end == cand[i + 1] && r.out_len == CHUNK && (r.end_flags & END_SYNC);
    else { ok = ok && (r.end_flags & END_FINAL); if (ok) result[1] = i * CHUNK + r.out_len; }
    if (!ok) atomicAnd(&result[0], 0ull);
}

}  // namespace b200
