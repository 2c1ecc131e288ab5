// huffman.cuh -- K2: per-chunk histogram reduce, length-limited Huffman code construction,
// dynamic-header generation, exact block-type selection and segment bit offsets.  One warp per chunk.
//
// Replaces (not ports) the reference's
//   CodeMap / constructDynamicHuffmanTree          include/deflate.hpp:35-79, 402-418
//   FlatHuffmanTree::generateCodeLengths           include/common.hpp:322-404   (priority queue + DFS)
//   FlatHuffmanTree::construct (canonical codes)   include/common.hpp:104-145
//   writeDynamicHuffmanTree & friends              include/deflate.hpp:419-626
//   the "encode twice, keep the smaller" choice    include/deflate.hpp:728-746
// Differences that matter: symbols are sorted once (warp bitonic sort) and code lengths come from the
// in-place Moffat-Katajainen pass instead of a heap + per-symbol tree search; over-long codes are
// repaired to an exact Kraft sum instead of throwing (common.hpp:398-402 falls back to fixed codes);
// block sizes are computed exactly from the histograms so nothing is encoded twice; the literal
// histogram counts emitted tokens only (the reference also counts bytes hidden under matches,
// deflate.hpp:406-414).  The literal/length and distance code-length lists are run-length coded
// SEPARATELY (the reference inflater cannot parse a run that crosses the two lists).
#pragma once
#include "common.cuh"

namespace b200 {

constexpr uint32_t HUF_WARPS = 4;
constexpr uint32_t HUF_THREADS = HUF_WARPS * 32;

__constant__ uint8_t PRECODE_ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

struct HufScratch {
    uint32_t freq[NSYM];     // reduced histogram (lit/len incl. EOB, dist)
    uint32_t keys[512];      // sort keys (freq << 9 | sym); reused per alphabet
    uint32_t A[NLIT];        // Moffat-Katajainen work array
    uint32_t next_code[16];
    uint32_t count[16];
    uint32_t pfreq[32];      // precode histogram
    uint16_t rle[NSYM + 8];  // precode symbol | extra << 8
    uint8_t lens[NSYM];      // code lengths, lit/len then dist
    uint8_t plens[32];       // precode lengths
    uint16_t pcodes[32];     // precode (bit-reversed) codes
    uint32_t misc[8];
};

// Bitonic sort of N (power of two, >= 32) keys in shared memory by one warp.
__device__ __forceinline__ void warp_sort(uint32_t* keys, uint32_t N, uint32_t lane) {
    for (uint32_t k = 2; k <= N; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = lane; i < N; i += 32) {
                uint32_t ixj = i ^ j;
                if (ixj > i) {
                    uint32_t a = keys[i], b = keys[ixj];
                    bool up = (i & k) == 0;
                    if ((a > b) == up) { keys[i] = b; keys[ixj] = a; }
                }
            }
            __syncwarp();
        }
    }
}

// Code lengths for one alphabet.  freq[0..nsym) -> lens[0..nsym), every length <= maxbits, Kraft sum
// exactly 1 (at least two codes are always emitted, as zlib does, so single-symbol alphabets stay
// decodable by strict decoders).  Returns false only if the repair could not reach Kraft == 1.
// Warp-cooperative: sort and scatter are parallel, the O(n) tree pass runs on lane 0.
__device__ bool build_lengths(HufScratch* s, const uint32_t* freq, uint8_t* lens, uint32_t nsym,
                              uint32_t maxbits, uint32_t lane) {
    // The symbols in use are packed to the front (keys are unique, so the sorted order does not depend on where they start)
    // and only the next power of two of them is sorted: a text chunk uses ~100 of the 288 literal/length symbols, and the
    // bitonic network over 128 keys is a sixth of the one over 512.
    const uint32_t NP = nsym > 32 ? 512 : 32;
    uint32_t used = 0;
    for (uint32_t i0 = 0; i0 < NP; i0 += 32) {
        const uint32_t i = i0 + lane;
        const uint32_t f = i < nsym ? freq[i] : 0;
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, f != 0);
        if (f) s->keys[used + __popc(m & ((1u << lane) - 1u))] = (f << 9) | i;
        used += __popc(m);
    }
    uint32_t N = 32;
    while (N < used) N <<= 1;
    for (uint32_t i = used + lane; i < N; i += 32) s->keys[i] = 0xFFFFFFFFu;
    for (uint32_t i = lane; i < nsym; i += 32) lens[i] = 0;
    __syncwarp();
    // force at least two symbols (zlib build_tree does the same): add the lowest unused symbols, freq 1
    if (used < 2) {
        if (lane == 0) {
            uint32_t have = 0xFFFFu, hk = 0;
            for (uint32_t i = 0; i < N; i++)
                if (s->keys[i] != 0xFFFFFFFFu) { hk = s->keys[i]; have = hk & 511u; }
            for (uint32_t i = 0; i < N; i++) s->keys[i] = 0xFFFFFFFFu;
            uint32_t k = 0;
            if (have != 0xFFFFu) s->keys[k++] = hk;
            for (uint32_t c = 0; k < 2; c++)
                if (c != have) s->keys[k++] = (1u << 9) | c;
        }
        used = 2;
        __syncwarp();
    }
    warp_sort(s->keys, N, lane);
    const uint32_t n = used;
    bool ok = true;
    if (lane == 0) {
        uint32_t* A = s->A;
        for (uint32_t i = 0; i < n; i++) A[i] = s->keys[i] >> 9;
        if (n == 2) {
            A[0] = A[1] = 1;
        } else {
            // Moffat & Katajainen, "In-place calculation of minimum-redundancy codes" (1995)
            A[0] += A[1];
            uint32_t root = 0, leaf = 2;
            for (uint32_t next = 1; next < n - 1; next++) {
                if (leaf >= n || A[root] < A[leaf]) { A[next] = A[root]; A[root++] = next; }
                else A[next] = A[leaf++];
                if (leaf >= n || (root < next && A[root] < A[leaf])) { A[next] += A[root]; A[root++] = next; }
                else A[next] += A[leaf++];
            }
            A[n - 2] = 0;
            for (int next = (int)n - 3; next >= 0; next--) A[next] = A[A[next]] + 1;
            int avbl = 1, usedn = 0, dpth = 0, rt = (int)n - 2, nx = (int)n - 1;
            while (avbl > 0) {
                while (rt >= 0 && (int)A[rt] == dpth) { usedn++; rt--; }
                while (avbl > usedn) { A[nx--] = dpth; avbl--; }
                avbl = 2 * usedn; dpth++; usedn = 0;
            }
        }
        // A[i] = length of the i-th least frequent symbol (non-increasing in i).  Limit to maxbits.
        if (A[0] > maxbits) {
            uint32_t* cnt = s->count;
            for (uint32_t l = 0; l < 16; l++) cnt[l] = 0;
            for (uint32_t i = 0; i < n; i++) cnt[min(A[i], maxbits)]++;
            const uint32_t cap = 1u << maxbits;
            uint32_t K = 0;
            for (uint32_t l = 1; l <= maxbits; l++) K += cnt[l] << (maxbits - l);
            while (K > cap) {   // demote the deepest code that is still above the limit
                uint32_t l = maxbits - 1;
                while (l > 0 && cnt[l] == 0) l--;
                if (l == 0) break;
                cnt[l]--; cnt[l + 1]++;
                K -= 1u << (maxbits - l - 1);
            }
            for (uint32_t l = maxbits; l >= 2 && K < cap; l--) {   // refill exactly, cheapest moves first
                const uint32_t unit = 1u << (maxbits - l);
                while (cnt[l] > 0 && cap - K >= unit) { cnt[l]--; cnt[l - 1]++; K += unit; }
            }
            ok = (K == cap);
            uint32_t i = 0;
            for (uint32_t l = maxbits; l >= 1; l--)
                for (uint32_t c = 0; c < cnt[l]; c++) A[i++] = l;
        }
    }
    __syncwarp();
    ok = __shfl_sync(0xFFFFFFFFu, (int)ok, 0);
    for (uint32_t i = lane; i < n; i += 32) lens[s->keys[i] & 511u] = (uint8_t)s->A[i];
    __syncwarp();
    return ok;
}

// Canonical codes (RFC 1951 3.2.2) for lens[0..nsym): out[i] = len | bit-reversed code << 8.
__device__ void assign_codes(HufScratch* s, const uint8_t* lens, uint32_t nsym, uint32_t* out, uint32_t lane) {
    if (lane < 16) s->count[lane] = 0;
    __syncwarp();
    for (uint32_t i = lane; i < nsym; i += 32)
        if (lens[i]) atomicAdd(&s->count[lens[i]], 1u);
    __syncwarp();
    if (lane == 0) {
        uint32_t code = 0;
        s->next_code[0] = 0;
        for (uint32_t l = 1; l <= 15; l++) {
            code = (code + (l > 1 ? s->count[l - 1] : 0)) << 1;
            s->next_code[l] = code;
        }
    }
    __syncwarp();
    for (uint32_t b = 0; b < nsym; b += 32) {
        const uint32_t i = b + lane;
        const uint32_t l = i < nsym ? lens[i] : 0;
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, l);
        const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
        uint32_t code = l ? s->next_code[l] + rank : 0;
        __syncwarp();
        if (l && rank == 0) s->next_code[l] += __popc(peers);
        __syncwarp();
        if (i < nsym) out[i] = l ? (l | (bitrev(code, l) << 8)) : 0;
    }
}

struct BitSink {   // lane-0 serial bit writer into a global word array
    uint32_t* dst;
    uint64_t acc;
    uint32_t nacc, nwords;
    __device__ void put(uint32_t v, uint32_t nb) {
        acc |= (uint64_t)v << nacc;
        nacc += nb;
        if (nacc >= 32) { dst[nwords++] = (uint32_t)acc; acc >>= 32; nacc -= 32; }
    }
    __device__ void flush() { if (nacc) { dst[nwords++] = (uint32_t)acc; acc = 0; nacc = 0; } }
};

// Run-length code one code-length list (RFC 1951 3.2.7) into s->rle[pos..]; returns new pos.
__device__ uint32_t rle_lengths(HufScratch* s, const uint8_t* lens, uint32_t n, uint32_t pos) {
    uint32_t i = 0;
    while (i < n) {
        const uint32_t v = lens[i];
        uint32_t run = 1;
        while (i + run < n && lens[i + run] == v) run++;
        i += run;
        if (v == 0) {
            while (run >= 11) { uint32_t r = min(run, 138u); s->rle[pos++] = (uint16_t)(18 | ((r - 11) << 8)); run -= r; }
            if (run >= 3) { s->rle[pos++] = (uint16_t)(17 | ((run - 3) << 8)); run = 0; }
            while (run--) s->rle[pos++] = 0;
        } else {
            s->rle[pos++] = (uint16_t)v; run--;
            while (run >= 3) { uint32_t r = min(run, 6u); s->rle[pos++] = (uint16_t)(16 | ((r - 3) << 8)); run -= r; }
            while (run--) s->rle[pos++] = (uint16_t)v;
        }
    }
    return pos;
}

__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

// What one candidate block would cost, and what the emitter needs to write it (the code lengths, the run-length coded
// header and the precode stay in the warp's HufScratch: a block is emitted right after it was planned).
struct BlockPlan {
    uint32_t use_dyn;        // 1: dynamic codes are smaller than the fixed ones
    uint32_t bits;           // whole block: 3-bit header (+ dynamic tables) + symbols + end of block
    uint32_t hdr_bits;       // 3 + dynamic tables (3 for a fixed block)
    uint32_t hlit, hdist, hclen, nr;
    uint32_t ok;             // code construction succeeded (always, in practice)
};

// Plans the block that covers segments [seg_lo, seg_hi) of the chunk whose per-segment histograms are h.
__device__ __noinline__ BlockPlan plan_block(HufScratch* s, const uint16_t* __restrict__ h, uint32_t seg_lo, uint32_t seg_hi, bool have_tokens,
                                uint32_t lane) {
    BlockPlan P;
    // ---- reduce the per-segment histograms ------------------------------------------------
    {
        uint32_t f[NSYM / 32];                 // this lane's ten symbols; ten loads in flight per segment
        #pragma unroll
        for (uint32_t k = 0; k < NSYM / 32; k++) f[k] = 0;
        if (have_tokens)                       // (level 0 runs no tokeniser)
            for (uint32_t sgm = seg_lo; sgm < seg_hi; sgm++) {
                #pragma unroll
                for (uint32_t k = 0; k < NSYM / 32; k++) f[k] += h[sgm * NSYM + lane + 32 * k];
            }
        #pragma unroll
        for (uint32_t k = 0; k < NSYM / 32; k++) s->freq[lane + 32 * k] = f[k] + (lane + 32 * k == 256 ? 1u : 0u);
    }
    __syncwarp();
    // ---- code lengths ------------------------------------------------------------------------
    bool ok = build_lengths(s, s->freq, s->lens, 286, 15, lane);
    ok &= build_lengths(s, s->freq + NLIT, s->lens + NLIT, 30, 15, lane);
    if (lane < 2) s->lens[286 + lane] = 0;
    if (lane < 2) s->lens[NLIT + 30 + lane] = 0;
    __syncwarp();
    // ---- dynamic header: HLIT/HDIST, RLE of both lists (separately), precode ---------------
    if (lane == 0) {
        uint32_t hlit = 286, hdist = 30;
        while (hlit > 257 && s->lens[hlit - 1] == 0) hlit--;
        while (hdist > 1 && s->lens[NLIT + hdist - 1] == 0) hdist--;
        uint32_t nr = rle_lengths(s, s->lens, hlit, 0);
        nr = rle_lengths(s, s->lens + NLIT, hdist, nr);
        for (uint32_t i = 0; i < 32; i++) s->pfreq[i] = 0;
        for (uint32_t i = 0; i < nr; i++) s->pfreq[s->rle[i] & 31u]++;
        s->misc[0] = hlit; s->misc[1] = hdist; s->misc[2] = nr;
    }
    __syncwarp();
    P.hlit = s->misc[0]; P.hdist = s->misc[1]; P.nr = s->misc[2];
    ok &= build_lengths(s, s->pfreq, s->plens, 19, 7, lane);
    {   // precode canonical codes: 19 symbols, one batch
        uint32_t tmp = 0;
        if (lane < 16) s->count[lane] = 0;
        __syncwarp();
        const uint32_t l = lane < 19 ? s->plens[lane] : 0;
        if (l) atomicAdd(&s->count[l], 1u);
        __syncwarp();
        if (lane == 0) {
            uint32_t code = 0;
            for (uint32_t b = 1; b <= 7; b++) { code = (code + (b > 1 ? s->count[b - 1] : 0)) << 1; s->next_code[b] = code; }
        }
        __syncwarp();
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, l);
        const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
        if (l) tmp = bitrev(s->next_code[l] + rank, l);
        if (lane < 32) s->pcodes[lane] = (uint16_t)tmp;
        __syncwarp();
    }
    // ---- exact sizes -----------------------------------------------------------------------
    uint32_t dyn = 0, fix = 0;
    for (uint32_t i = lane; i < NSYM; i += 32) {
        const uint32_t f = s->freq[i];
        if (!f) continue;
        if (i < NLIT) {
            const uint32_t ex = i > 256 ? len_extra_bits(i - 257) : 0;
            dyn += f * (s->lens[i] + ex);
            fix += f * (fixed_lit_len(i) + ex);
        } else {
            const uint32_t ex = dist_extra_bits(i - NLIT);
            dyn += f * (s->lens[i] + ex);
            fix += f * (5 + ex);
        }
    }
    dyn = warp_sum(dyn);
    fix = warp_sum(fix);
    uint32_t hclen = 19;
    while (hclen > 4 && s->plens[PRECODE_ORDER[hclen - 1]] == 0) hclen--;
    uint32_t hdr_bits = 3 + 5 + 5 + 4 + 3 * hclen;
    {
        uint32_t hb = 0;
        for (uint32_t i = lane; i < P.nr; i += 32) {
            const uint32_t r = s->rle[i], sym = r & 31u;
            hb += s->plens[sym] + (sym == 16 ? 2u : sym == 17 ? 3u : sym == 18 ? 7u : 0u);
        }
        hdr_bits += warp_sum(hb);
    }
    const uint32_t dyn_bits = ok ? hdr_bits + dyn : 0xFFFFFFFFu;
    const uint32_t fix_bits = 3 + fix;
    P.use_dyn = dyn_bits < fix_bits;
    P.bits = P.use_dyn ? dyn_bits : fix_bits;
    P.hdr_bits = P.use_dyn ? hdr_bits : 3u;
    P.hclen = hclen;
    P.ok = ok;
    return P;
}

// Writes the planned block's codes (what the encoder indexes by symbol) and its header bits.  btype 2 dynamic / 1 fixed.
__device__ __noinline__ void emit_block(HufScratch* s, const BlockPlan& P, uint32_t btype, bool last, uint32_t* __restrict__ mycodes,
                           uint32_t* __restrict__ myhdr, uint32_t lane) {
    if (btype == 2) {
        assign_codes(s, s->lens, NLIT, mycodes, lane);
        assign_codes(s, s->lens + NLIT, NDIST, mycodes + NLIT, lane);
    } else {
        for (uint32_t i = lane; i < NSYM; i += 32) {
            uint32_t v;
            if (i < NLIT) {   // common.hpp:442-482 (RFC 1951 3.2.6)
                uint32_t l = fixed_lit_len(i);
                uint32_t c = i < 144 ? 0x30 + i : i < 256 ? 0x190 + (i - 144) : i < 280 ? (i - 256) : 0xC0 + (i - 280);
                v = l | (bitrev(c, l) << 8);
            } else {
                v = 5u | (bitrev(i - NLIT, 5) << 8);
            }
            mycodes[i] = v;
        }
    }
    __syncwarp();
    if (lane == 0) {
        if (btype == 2) {
            BitSink bs{myhdr, 0, 0, 0};
            bs.put((last ? 1u : 0u) | (2u << 1), 3);
            bs.put(P.hlit - 257, 5);
            bs.put(P.hdist - 1, 5);
            bs.put(P.hclen - 4, 4);
            for (uint32_t i = 0; i < P.hclen; i++) bs.put(s->plens[PRECODE_ORDER[i]], 3);
            for (uint32_t i = 0; i < P.nr; i++) {
                const uint32_t r = s->rle[i], sym = r & 31u;
                bs.put(s->pcodes[sym], s->plens[sym]);
                if (sym == 16) bs.put(r >> 8, 2);
                else if (sym == 17) bs.put(r >> 8, 3);
                else if (sym == 18) bs.put(r >> 8, 7);
            }
            bs.flush();
        } else {
            myhdr[0] = (last ? 1u : 0u) | (1u << 1);
        }
    }
    __syncwarp();
}

// payload bits of segments [seg_lo, seg_hi) under the code lengths currently in the scratch; offsets go to d->seg_bitoff
__device__ __noinline__ uint32_t segment_offsets(HufScratch* s, const uint16_t* __restrict__ h, uint32_t btype, uint32_t seg_lo, uint32_t seg_hi,
                                    bool have_tokens, uint32_t off, BlockDesc* d, uint32_t lane) {
    // bits per symbol of this lane's ten symbols: the same for every segment
    uint32_t cost[NSYM / 32];
    #pragma unroll
    for (uint32_t k = 0; k < NSYM / 32; k++) {
        const uint32_t i = lane + 32 * k;
        const uint32_t cl = btype == 2 ? s->lens[i] : (i < NLIT ? fixed_lit_len(i) : 5u);
        const uint32_t ex = i < NLIT ? (i > 256 ? len_extra_bits(i - 257) : 0u) : dist_extra_bits(i - NLIT);
        cost[k] = cl + ex;
    }
    for (uint32_t sgm = seg_lo; sgm < seg_hi; sgm++) {
        uint32_t sb = 0;
        if (have_tokens) {
            uint32_t f[NSYM / 32];            // ten loads in flight instead of a load -> test -> multiply chain per symbol
            #pragma unroll
            for (uint32_t k = 0; k < NSYM / 32; k++) f[k] = h[sgm * NSYM + lane + 32 * k];
            #pragma unroll
            for (uint32_t k = 0; k < NSYM / 32; k++) sb += f[k] * cost[k];
        }
        sb = warp_sum(sb);
        if (lane == 0) d->seg_bitoff[sgm] = off;
        off += sb;
    }
    return off;
}

// Block splitting (the reference emits a block per 32 KB, deflate.hpp:692-749).  A full chunk may
// be cut at a segment boundary into TWO blocks with their own code tables when the statistics change inside it.  Cost
// model: the zero-order entropy of the literal/length + distance histograms (already in HBM per 4 KiB segment) of the
// two sides against the whole, for the cuts at 16, 32 and 48 KiB; the best cut is taken to the exact stage -- both
// blocks are planned for real -- only if the estimate beats the price of a second header plus SPLIT_MARGIN of the chunk,
// and is used only if the exact bit count is smaller by that margin.  The margin is what the segment index is worth: a
// split chunk carries none (two code tables per chunk do not fit pass A's per-chunk table slot) and is inflated by the
// one-warp decoder, ~10x slower than an indexed chunk -- so drifting statistics (a few hundred bytes to gain) do not
// split, a boundary between two kinds of content (tar members, the halves in tests/test_gpu_compress.py) does.
constexpr uint32_t SPLIT_MARGIN_SHIFT = 5;       // 1/32 of the one-block size (3 %)
__device__ __noinline__ uint32_t split_candidate(const uint16_t* __restrict__ h, uint32_t hdr_bits, uint32_t one_block_bits, uint32_t lane) {
    float gain[3] = {0.f, 0.f, 0.f};
    uint32_t nA[3][2] = {{0, 0}, {0, 0}, {0, 0}}, nS[2] = {0, 0};
    // per-lane symbols: counts left of each cut and in total
    uint32_t fA[3][NSYM / 32], fS[NSYM / 32];
    #pragma unroll
    for (uint32_t k = 0; k < NSYM / 32; k++) {
        const uint32_t i = lane + 32 * k;
        uint32_t run = 0;
        for (uint32_t sgm = 0; sgm < NSEG; sgm++) {
            if (sgm == 4) fA[0][k] = run;
            if (sgm == 8) fA[1][k] = run;
            if (sgm == 12) fA[2][k] = run;
            run += h[sgm * NSYM + i];
        }
        fS[k] = run;
        const uint32_t alpha = i < NLIT ? 0 : 1;
        nS[alpha] += run;
        #pragma unroll
        for (uint32_t c = 0; c < 3; c++) nA[c][alpha] += fA[c][k];
    }
    #pragma unroll
    for (uint32_t a = 0; a < 2; a++) {
        nS[a] = warp_sum(nS[a]);
        #pragma unroll
        for (uint32_t c = 0; c < 3; c++) nA[c][a] = warp_sum(nA[c][a]);
    }
    // sum f log2(N / f): bits of the side under its own ideal code
    #pragma unroll
    for (uint32_t k = 0; k < NSYM / 32; k++) {
        const uint32_t i = lane + 32 * k;
        const uint32_t a = i < NLIT ? 0 : 1;
        const float s_bits = fS[k] ? (float)fS[k] * (__log2f((float)nS[a]) - __log2f((float)fS[k])) : 0.f;
        #pragma unroll
        for (uint32_t c = 0; c < 3; c++) {
            const uint32_t fa = fA[c][k], fb = fS[k] - fa, na = nA[c][a], nb = nS[a] - na;
            const float a_bits = fa ? (float)fa * (__log2f((float)na) - __log2f((float)fa)) : 0.f;
            const float b_bits = fb ? (float)fb * (__log2f((float)nb) - __log2f((float)fb)) : 0.f;
            gain[c] += s_bits - a_bits - b_bits;
        }
    }
    uint32_t best = 0;
    float bestg = 0.f;
    #pragma unroll
    for (uint32_t c = 0; c < 3; c++) {
        float g = gain[c];
        for (int o = 16; o > 0; o >>= 1) g += __shfl_xor_sync(0xFFFFFFFFu, g, o);
        if (g > bestg) { bestg = g; best = 4 * (c + 1); }
    }
    // worth the exact stage only if the ideal-code saving pays a second header and the margin
    return bestg > (float)hdr_bits + 256.f + (float)(one_block_bits >> SPLIT_MARGIN_SHIFT) ? best : 0u;
}

// grid = ceil(nchunks / HUF_WARPS), block = 128.
//   hist   : [nchunks][NSEG][NSYM] u16 from K1 (a segment has at most 4096 tokens)
//   codes  : [nchunks][2][NSYM]   len | reversed code << 8   (what the encoder indexes by symbol; [1] = second block of a split chunk)
//   hdr    : [nchunks][2][HDR_WORDS] block header bits (BFINAL/BTYPE + dynamic tables)
//   desc   : [nchunks] BlockDesc;  sizes: [nchunks] bytes per chunk (input of the offset scan)
// level 0 forces stored blocks.  last_is_final: the final chunk of this buffer carries BFINAL.
__global__ void __launch_bounds__(HUF_THREADS)
huffman_kernel(const uint16_t* __restrict__ hist, uint64_t n, uint32_t nchunks, int level, int last_is_final,
               int with_index, int with_split, const ChunkSrc* __restrict__ srcs, uint32_t* __restrict__ codes, uint32_t* __restrict__ hdr,
               BlockDesc* __restrict__ desc, uint32_t* __restrict__ sizes) {
    __shared__ HufScratch scratch[HUF_WARPS];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t chunk = blockIdx.x * HUF_WARPS + warp;
    if (chunk >= nchunks) return;
    HufScratch* s = &scratch[warp];
    const uint32_t clen = srcs ? srcs[chunk].clen : (uint32_t)min((uint64_t)CHUNK, n - (uint64_t)chunk * CHUNK);
    const bool last = srcs ? (srcs[chunk].last & 1u) != 0 : (last_is_final && chunk == nchunks - 1);
    const uint16_t* h = hist + (size_t)chunk * NSEG * NSYM;
    uint32_t* mycodes = codes + (size_t)chunk * 2 * NSYM;
    uint32_t* myhdr = hdr + (size_t)chunk * 2 * HDR_WORDS;
    BlockDesc* d = desc + chunk;

    const uint32_t nstored = (clen + 65534u) / 65535u;
    // an empty input (batch compression only) still needs a block: stored is not an option, the fixed block wins (03 00)
    const uint32_t stored_bytes = clen ? clen + 5u * nstored + (last ? 0u : SYNC_BYTES_ALIGNED) : 0xFFFFFFFFu;

    if (level == 0 && clen) {
        if (lane == 0) {
            d->btype = 0; d->hdr_bits = 0; d->total_bits = 0; d->nbytes = stored_bytes; d->clen = clen;
            d->last = last; d->eob = 0; d->index_bytes = 0; d->split_seg = 0;
            sizes[chunk] = stored_bytes;
        }
        return;
    }
    const bool have_tokens = clen != 0;

    // ---- one block for the whole chunk ---------------------------------------------------------------------
    BlockPlan S = plan_block(s, h, 0, NSEG, have_tokens, lane);
    // ---- or two?  (full chunks only: the cut is a segment boundary) -------------------------------------------
    uint32_t split = 0;
    bool scratch_is_S = true;                     // the scratch still holds the whole-chunk plan
    BlockPlan A = S, B = S;
    if (with_split && clen == CHUNK) {
        const uint32_t cut = split_candidate(h, S.hdr_bits, S.bits, lane);
        if (cut) {
            A = plan_block(s, h, 0, cut, true, lane);
            B = plan_block(s, h, cut, NSEG, true, lane);
            scratch_is_S = false;
            if (A.bits + B.bits + (S.bits >> SPLIT_MARGIN_SHIFT) < S.bits) split = cut;
        }
    }
    const uint32_t bits = split ? A.bits + B.bits : S.bits;
    // chunks carry the segment index (common.cuh) so that they can be inflated by one thread per segment -- the short last
    // chunk of a stream too if it has INDEX_MIN_SEGS segments or more.  (Measured, tools/probe_small.py: a lone thread takes
    // 0.4-0.9 ms for its 4 KiB segment and the copy pass 0.2 ms more, whatever the chunk's size; one warp with the output
    // window in shared memory -- inflate_small_kernel -- needs 26-45 us per KiB: it wins below ~30 KiB, e.g. test.bmp,
    // 21 KB: 0.56 against 0.73 ms.)  Only where it costs little: the chunk must save at least four times the index's size;
    // never for a split chunk.
    const uint32_t huff_plain = last ? (bits + 7) / 8 : (bits + 3 + 7) / 8 + (SYNC_BYTES_ALIGNED - 1);
    const uint32_t nseg_used = (clen + SEG - 1) / SEG;
    const uint32_t index_want = nseg_used >= INDEX_MIN_SEGS ? nseg_used * INDEX_BYTES_PER_SEG : 0u;
    const uint32_t index_bytes = (with_index && !split && index_want && huff_plain + 4 * index_want <= stored_bytes) ? index_want : 0u;
    const uint32_t huff_bytes = index_bytes + huff_plain;
    const bool huffman = huff_bytes < stored_bytes;

    if (!huffman) {
        if (lane == 0) {
            d->btype = 0; d->hdr_bits = 0; d->total_bits = 0; d->nbytes = stored_bytes; d->clen = clen;
            d->last = last; d->eob = 0; d->index_bytes = 0; d->split_seg = 0;
            sizes[chunk] = stored_bytes;
        }
        return;
    }
    // ---- emit: the scratch holds the LAST planned block, so each block is planned again right before it is written ----
    uint32_t btype, off;
    if (!split) {
        if (!scratch_is_S) S = plan_block(s, h, 0, NSEG, have_tokens, lane);
        btype = S.use_dyn ? 2u : 1u;
        emit_block(s, S, btype, last, mycodes, myhdr, lane);
        off = segment_offsets(s, h, btype, 0, NSEG, have_tokens, S.hdr_bits, d, lane);
        if (lane == 0) {
            const uint32_t eob_len = btype == 2 ? s->lens[256] : 7u;
            d->btype = btype; d->hdr_bits = S.hdr_bits; d->total_bits = off + eob_len;
            d->split_seg = 0; d->btype2 = 0; d->hdr_bits2 = 0; d->block2_bit = 0;
        }
    } else {
        A = plan_block(s, h, 0, split, true, lane);
        btype = A.use_dyn ? 2u : 1u;
        emit_block(s, A, btype, false, mycodes, myhdr, lane);
        off = segment_offsets(s, h, btype, 0, split, true, A.hdr_bits, d, lane);
        const uint32_t eobA = btype == 2 ? s->lens[256] : 7u;
        const uint32_t block2 = off + eobA;                    // first bit of the second block's header
        B = plan_block(s, h, split, NSEG, true, lane);
        const uint32_t btype2 = B.use_dyn ? 2u : 1u;
        emit_block(s, B, btype2, last, mycodes + NSYM, myhdr + HDR_WORDS, lane);
        off = segment_offsets(s, h, btype2, split, NSEG, true, block2 + B.hdr_bits, d, lane);
        if (lane == 0) {
            const uint32_t eobB = btype2 == 2 ? s->lens[256] : 7u;
            d->btype = btype; d->hdr_bits = A.hdr_bits; d->total_bits = off + eobB;
            d->split_seg = split; d->btype2 = btype2; d->hdr_bits2 = B.hdr_bits; d->block2_bit = block2;
        }
    }
    if (lane == 0) {
        d->nbytes = huff_bytes;
        d->clen = clen;
        d->last = last;
        d->index_bytes = index_bytes;
        d->eob = 0;
        sizes[chunk] = huff_bytes;
    }
}

}  // namespace b200
