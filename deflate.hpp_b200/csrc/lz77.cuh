// lz77.cuh -- K1: per-chunk LZ77 tokenisers + histograms.
//
// Replaces the reference's LZ77::getMatches (include/deflate.hpp:310-383, "fast"), getMatchesSlow
// (:268-304, "better") and the histogram half of constructDynamicHuffmanTree (:402-418).  Not a port:
// the reference probes only 4-byte-aligned positions of a 32 KB chunk with one thread and a
// first-occurrence table (:373-376), or scans all earlier positions for every position (:280-296).
// Here one CTA owns a 64 KiB chunk staged in shared memory by a TMA bulk copy; candidates are found
// for ALL positions in parallel against a chunk-wide structure (hash table for "fast", hash chains
// for "better"), then each warp parses one 4 KiB segment greedily (or lazily), 32 positions per step
// with ballot-mask selection.  Tokens go to a global scratch buffer (one word per TOKEN, not per
// byte); literal/length + distance histograms are accumulated per segment with shared-memory atomics.
#pragma once
#include "common.cuh"

namespace b200 {

constexpr uint32_t LZ_DATA_PAD = 64;           // zeroed over-read slack after the chunk

__device__ __forceinline__ uint32_t mask_range(uint32_t a, uint32_t b) {   // bits [a, b), a <= b <= 32
    uint32_t hi = b >= 32 ? 0xFFFFFFFFu : ((1u << b) - 1u);
    uint32_t lo = a >= 32 ? 0xFFFFFFFFu : ((1u << a) - 1u);
    return hi & ~lo;
}

// --- TMA 1-D bulk copy global -> shared, completion on an mbarrier ----------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
    uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d),
        "l"(gsrc), "r"(bytes), "r"(b)
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(a), "r"(parity)
            : "memory");
    }
}

// Per-segment histogram in shared memory, two 16-bit counters per word (a segment has at most
// SEG = 4096 tokens, so a counter cannot carry into its neighbour).
constexpr uint32_t HIST_WORDS = NSYM / 2;
__device__ __forceinline__ void hist_add(uint32_t* h, uint32_t sym) { atomicAdd(&h[sym >> 1], 1u << ((sym & 1) * 16)); }
// global layout: [chunk][segment][NSYM] u16 -- exactly the shared-memory words (little endian), copied as u32
__device__ __forceinline__ void hist_store(const uint32_t* h, uint16_t* gh, uint32_t lane) {
    uint32_t* g32 = reinterpret_cast<uint32_t*>(gh);
    for (uint32_t i = lane; i < HIST_WORDS; i += 32) g32[i] = h[i];
}

// =====================================================================================================
// Level 1 (Huffman only, no matching -- reference deflate.hpp:706-708): every byte is a literal token.
// grid = chunks, 512 threads, warp s handles segment s.  No shared-memory staging needed.
// =====================================================================================================
constexpr uint32_t LZL_THREADS = NSEG * 32;
__global__ void __launch_bounds__(LZL_THREADS)
lz77_literal_kernel(const uint8_t* __restrict__ in, uint64_t n, uint32_t* __restrict__ tok,
                    uint32_t* __restrict__ ntok, uint16_t* __restrict__ hist, const ChunkSrc* __restrict__ srcs) {
    __shared__ uint32_t s_hist[NSEG * HIST_WORDS];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t chunk = blockIdx.x;
    const uint64_t base = srcs ? srcs[chunk].off : chunk * CHUNK;
    const uint32_t clen = srcs ? srcs[chunk].clen : (uint32_t)min((uint64_t)CHUNK, n - base);
    for (uint32_t i = tid; i < NSEG * HIST_WORDS; i += LZL_THREADS) s_hist[i] = 0;
    __syncthreads();
    const uint32_t seg_lo = warp * SEG, seg_hi = min(clen, seg_lo + SEG);
    uint32_t* hs = s_hist + warp * HIST_WORDS;
    // a segment of nothing but literals writes no tokens: the encoder reads the bytes from the input (NTOK_LITERALS)
    for (uint32_t p = seg_lo + lane; p < seg_hi; p += 32) hist_add(hs, in[base + p]);
    __syncwarp();
    if (lane == 0) ntok[chunk * NSEG + warp] = seg_lo < clen ? ((seg_hi - seg_lo) | NTOK_LITERALS) : 0;
    (void)tok;
    hist_store(hs, hist + (size_t)(chunk * NSEG + warp) * NSYM, lane);
}

// =====================================================================================================
// "fast" level, two-phase matcher (level 2 default).
//
// Persistent CTAs (grid = 2 x SMs) pull chunks from a global counter.  Per chunk:
//   phase B  all 8 warps sweep the chunk in tiles of 1024 positions against ONE chunk-wide hash table
//            (8 K x u32 in shared memory, 4-byte hash).  Per tile: every thread looks its buckets up,
//            __syncthreads, every thread inserts its positions with atomicMax -- lookups therefore see
//            exactly the positions before the tile and the table content is deterministic (highest
//            position wins).  Four candidates per position: the bucket and the distances 1, 2, 3 (runs,
//            short periods, RGB pixels -- what a tile-lagged table cannot see), the latter straight
//            from registers.  Each is verified on 4 bytes and scored on the next 4 (no loops);
//            the best score / nearest distance goes to a u16 candidate array in a per-CTA scratch
//            slot (128 KB per CTA, 38 MB in total: it lives in L2 and is never meant to reach HBM).
//   phase C  each warp parses one 4 KiB segment greedily: 32 positions per step, every lane extends
//            its own candidate to full length (4 bytes per iteration), ballot-mask selection, tokens
//            and shared-memory histograms exactly as in the single-phase kernel.
// Against lz77_kernel<0>: the whole 32 KiB window is reachable from every position instead of an
// 8 KiB segment-private table (large.bmp stand-in: 2x smaller output, corpus: 0.71 -> 0.62).
// =====================================================================================================
constexpr uint32_t LZF_THREADS = NSEG * 32;          // 512
constexpr uint32_t LZF_HASH_BITS = 13;
constexpr uint32_t LZF_PER_THREAD = 2;
constexpr uint32_t LZF_TILE = LZF_PER_THREAD * LZF_THREADS;     // 1024 positions between two table updates
constexpr uint32_t LZF_RING_BLOCKS = 4;                 // per-warp ring of candidate blocks (32 x u16 each) in shared memory
constexpr size_t LZF_SMEM_BYTES = CHUNK + LZ_DATA_PAD + (4u << LZF_HASH_BITS) + NSEG * HIST_WORDS * 4 + 32 +
                                  NSEG * LZF_RING_BLOCKS * 64 + (NSEG + 16) * 4;
// Adaptive matching (LZ4-style acceleration): where the data has no repeats (random / encrypted / already compressed)
// looking every position up is wasted work.  Tile 0 and every LZF_SAMPLE-th tile after tile 1 are always searched; if
// such a sample tile finds fewer than LZF_SAMPLE_MIN candidates among its 1024 positions, the tiles up to the next
// sample are not searched at all (no lookups, no inserts, candidate = none).  A segment without a single candidate is
// parsed by a literal-only loop.  The decision is per 4 KiB, so a chunk that turns compressible is picked up again
// after at most 3 tiles.  (B200_LZF_ADAPTIVE=0 switches it off; the corpus ratio is unchanged to 4 digits.)
constexpr uint32_t LZF_SAMPLE = 4;
constexpr uint32_t LZF_SAMPLE_MIN = 8;

__device__ __forceinline__ uint32_t lzf_hash(uint32_t w4) { return (w4 * 0x9E3779B1u) >> (32 - LZF_HASH_BITS); }

// candidate q for position p: 0 if the first 4 bytes differ, else 4 + (matching bytes among the next 4)
__device__ __forceinline__ uint32_t lzf_score(const uint8_t* d, uint32_t q, uint32_t w4, uint32_t w8) {
    if (ld4_unaligned(d, q) != w4) return 0;
    const uint32_t x = ld4_unaligned(d, q + 4) ^ w8;
    return x ? 4 + ((__ffs(x) - 1) >> 3) : 8;
}

__global__ void __launch_bounds__(LZF_THREADS, 2)
lz77_fast_kernel(const uint8_t* __restrict__ in, uint64_t n, uint32_t nchunks, uint32_t* __restrict__ tok,
                 uint32_t* __restrict__ ntok, uint16_t* __restrict__ hist, uint16_t* __restrict__ cand_scratch,
                 unsigned int* __restrict__ counter, const ChunkSrc* __restrict__ srcs, uint32_t adaptive) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* s_data = smem;
    uint32_t* s_tab = reinterpret_cast<uint32_t*>(smem + CHUNK + LZ_DATA_PAD);
    uint32_t* s_hist = s_tab + (1u << LZF_HASH_BITS);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_hist + NSEG * HIST_WORDS);
    uint32_t* s_next = reinterpret_cast<uint32_t*>(s_bar + 1);
    uint16_t* s_ring = reinterpret_cast<uint16_t*>(smem + CHUNK + LZ_DATA_PAD + (4u << LZF_HASH_BITS) + NSEG * HIST_WORDS * 4 + 32);
    uint32_t* s_segcnt = reinterpret_cast<uint32_t*>(smem + CHUNK + LZ_DATA_PAD + (4u << LZF_HASH_BITS) + NSEG * HIST_WORDS * 4 + 32 +
                                                     NSEG * LZF_RING_BLOCKS * 64);      // [NSEG] candidates per segment, [NSEG] = sample tile count

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t FULL = 0xFFFFFFFFu;
    uint16_t* cand = cand_scratch + (size_t)blockIdx.x * CHUNK;
    if (tid == 0) mbar_init(s_bar, 1);
    uint32_t parity = 0;

    for (;;) {
        __syncthreads();                                   // previous chunk fully done (s_next, smem reuse)
        if (tid == 0) *s_next = atomicAdd(counter, 1u);
        __syncthreads();
        const uint64_t chunk = *s_next;
        if (chunk >= nchunks) break;
        const uint64_t base = srcs ? srcs[chunk].off : chunk * CHUNK;
        const uint32_t clen = srcs ? srcs[chunk].clen : (uint32_t)min((uint64_t)CHUNK, n - base);
        const uint8_t* src = in + base;

        const bool aligned = ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
        const uint32_t bulk = aligned ? (clen & ~15u) : 0;
        if (tid == 0 && bulk) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic reads of s_data are done
            tma_load_1d(s_data, src, bulk, s_bar);
        }
        for (uint32_t i = bulk + tid; i < clen; i += LZF_THREADS) s_data[i] = src[i];
        for (uint32_t i = clen + tid; i < ((clen + 15u) & ~15u) + LZ_DATA_PAD && i < CHUNK + LZ_DATA_PAD; i += LZF_THREADS)
            s_data[i] = 0;
        for (uint32_t i = tid; i < (1u << LZF_HASH_BITS) + NSEG * HIST_WORDS; i += LZF_THREADS) s_tab[i] = 0;
        if (tid < NSEG + 1) s_segcnt[tid] = 0;
        if (bulk) { mbar_wait(s_bar, parity); parity ^= 1; }
        __syncthreads();

        // ---- phase B: one verified candidate distance per position ----------------------------------
        // Tile = 1024 positions, 2 per thread (tid, tid+512): each warp still holds 32 consecutive
        // positions per sub-tile, so neighbours' words are a shuffle away.
        bool skip = false;
        for (uint32_t t0 = 0; t0 < clen; t0 += LZF_TILE) {
            const uint32_t tile = t0 / LZF_TILE;
            const bool sample = tile == 0 || (tile % LZF_SAMPLE) == 1;
            if (skip && !sample) {
                // nothing to find here (the last sample tile said so): no candidates, no table update
                #pragma unroll
                for (uint32_t k = 0; k < LZF_PER_THREAD; k++) {
                    const uint32_t p = t0 + k * LZF_THREADS + tid;
                    if (p < clen) cand[p] = 0;
                }
                continue;
            }
            uint32_t found = 0;
            uint32_t w4[LZF_PER_THREAD], w8[LZF_PER_THREAD], hq[LZF_PER_THREAD];
            #pragma unroll
            for (uint32_t k = 0; k < LZF_PER_THREAD; k++) {
                const uint32_t p = t0 + k * LZF_THREADS + tid;
                const uint32_t* w = reinterpret_cast<const uint32_t*>(s_data) + (p >> 2);
                const uint32_t a = w[0], b = w[1], c = w[2], sh = (p & 3) * 8;
                w4[k] = __funnelshift_r(a, b, sh);
                w8[k] = __funnelshift_r(b, c, sh);
                hq[k] = s_tab[lzf_hash(w4[k])];
            }
            __syncthreads();                               // every lookup of this tile is done
            #pragma unroll
            for (uint32_t k = 0; k < LZF_PER_THREAD; k++) {
                const uint32_t p = t0 + k * LZF_THREADS + tid;
                if (p + 4 <= clen) atomicMax(&s_tab[lzf_hash(w4[k])], p);
            }
            #pragma unroll
            for (uint32_t k = 0; k < LZF_PER_THREAD; k++) {
                const uint32_t p = t0 + k * LZF_THREADS + tid;
                // the 4 bytes before p: lane-4's word, or a shared-memory read at the warp's left edge
                uint32_t prev4 = __shfl_up_sync(FULL, w4[k], 4);
                if (lane < 4) prev4 = p >= 4 ? ld4_unaligned(s_data, p - 4)
                                             : (p ? reinterpret_cast<const uint32_t*>(s_data)[0] << (8 * (4 - p)) : 0u);
                uint32_t best = 0, bdist = 0;
                if (p + 4 <= clen) {
                    const uint32_t q = hq[k];
                    if (q < p && p - q <= MAX_DIST) { best = lzf_score(s_data, q, w4[k], w8[k]); bdist = p - q; }
                    // distances 1..3 (runs, 2- and 3-byte periods, RGB pixels) straight from registers:
                    // bytes [p-d, p-d+4) = funnel(prev4, w4), bytes [p-d+4, p-d+8) = funnel(w4, w8)
                    #pragma unroll
                    for (uint32_t d = 3; d >= 1; d--) {
                        if (p >= d && __funnelshift_r(prev4, w4[k], 8 * (4 - d)) == w4[k]) {
                            const uint32_t x = __funnelshift_r(w4[k], w8[k], 8 * (4 - d)) ^ w8[k];
                            const uint32_t sc = x ? 4 + ((__ffs(x) - 1) >> 3) : 8;
                            if (sc >= best) { best = sc; bdist = d; }            // ties: the nearer one
                        }
                    }
                    if (!best) bdist = 0;
                }
                if (p < clen) cand[p] = (uint16_t)bdist;   // 32768 == 0x8000 still fits
                found += bdist != 0;
            }
            {
                // candidates of this tile: per segment (phase C takes a literal-only loop for segments without any) and,
                // for a sample tile, in total
                const uint32_t wsum = __reduce_add_sync(FULL, found);
                if (lane == 0 && wsum) {
                    atomicAdd(&s_segcnt[tile >> 2], wsum);
                    if (sample && tile) atomicAdd(&s_segcnt[NSEG], wsum);
                }
            }
            __syncthreads();                               // inserts land before the next tile looks up
            if (sample && tile && adaptive) {
                skip = s_segcnt[NSEG] < LZF_SAMPLE_MIN;
                __syncthreads();                           // everybody has read the count
                if (tid == 0) s_segcnt[NSEG] = 0;
            }
        }
        __syncthreads();
        // the block-wide barrier above also orders this CTA's cand[] stores before the loads below

        // ---- phase C: greedy parse, one warp per segment ----------------------------------------------
        const uint32_t seg_lo = warp * SEG;
        const uint32_t seg_hi = min(clen, seg_lo + SEG);
        uint32_t* hs = s_hist + warp * HIST_WORDS;
        uint32_t* mytok = tok + chunk * CHUNK + seg_lo;
        uint32_t nt = 0;
        if (seg_lo < clen && s_segcnt[warp] == 0) {
            // not one candidate in this segment: every byte is a literal, and no tokens are written -- the encoder reads
            // the bytes from the input (a quarter of a MB per random chunk that neither kernel has to move)
            for (uint32_t p = seg_lo + lane; p < seg_hi; p += 32) hist_add(hs, s_data[p]);
            nt = (seg_hi - seg_lo) | NTOK_LITERALS;
        } else if (seg_lo < clen) {
            uint32_t pos = seg_lo;
            // candidates travel through a per-warp ring of four 32-entry blocks in shared memory, filled by
            // cp.async (no destination register: a register queue has to be shifted one step after the load
            // was issued and stalls on it -- 10 % of this kernel's stall samples).  Blocks b0 .. b0+3 of the
            // current position are always requested; b0 and b0+1 are waited for.  (L1 is bypassed: the slot
            // is rewritten for every chunk this CTA processes.)
            uint16_t* ring = s_ring + warp * (LZF_RING_BLOCKS * 32);
            const uint32_t ring_sa = (uint32_t)__cvta_generic_to_shared(ring);
            uint32_t req_next = pos >> 5;                  // first block not requested yet
            while (pos < seg_hi) {
                const uint32_t b0 = pos >> 5;
                if (req_next < b0) req_next = b0;          // a long match jumped over blocks nobody needs
                #pragma unroll 1
                while (req_next < b0 + LZF_RING_BLOCKS) {
                    if (lane < 4 && req_next < CHUNK / 32)         // 64 bytes per block: four 16-byte copies, L2 only (.cg)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::
                                     "r"(ring_sa + (req_next % LZF_RING_BLOCKS) * 64 + lane * 16), "l"(cand + req_next * 32 + lane * 8) : "memory");
                    asm volatile("cp.async.commit_group;" ::: "memory");
                    req_next++;
                }
                asm volatile("cp.async.wait_group 2;" ::: "memory");        // all but the two newest blocks: b0, b0+1 are in
                __syncwarp();
                const uint32_t p = pos + lane;
                const uint32_t avail = p < seg_hi ? seg_hi - p : 0;
                uint32_t len = 0, dist = 0;
                if (avail >= 4) {
                    dist = ring[((p >> 5) % LZF_RING_BLOCKS) * 32 + (p & 31)];
                    if (dist) {
                        const uint32_t q = p - dist;
                        const uint32_t maxl = min(avail, MAX_MATCH);
                        uint32_t l = 4;
                        while (l < maxl) {
                            const uint32_t x = ld4_unaligned(s_data, p + l) ^ ld4_unaligned(s_data, q + l);
                            if (x) { l += (__ffs(x) - 1) >> 3; break; }
                            l += 4;
                        }
                        len = min(l, maxl);
                    }
                }
                const uint32_t byte = s_data[p];
                const uint32_t valid = min(32u, seg_hi - pos);
                // greedy in-order selection: the serial part only walks from match to match
                // (first match lane at or after the frontier, jump to its end) ...
                // (Measured and rejected: the walk replaced by three rounds of pointer doubling over per-lane successor
                // links -- the selected set is what is reachable from the window's first match, at most 8 matches -- gives the
                // same bytes in the same 9.5 ms: the walk's latency is hidden by the other 31 warps of the SM.)
                // Every lane works out beforehand which matches start at or behind the end of its own (`after`), so a step of the
                // walk is: lowest set bit, OR it in, fetch that lane's `after` -- 7 instructions instead of 12.
                const uint32_t mm = __ballot_sync(FULL, len >= 4);
                const uint32_t endl = lane + len;                        // where this lane's match would end
                const uint32_t after = endl < 32 ? mm & (FULL << endl) : 0u;
                uint32_t selmask = 0;
                for (uint32_t m = mm; m;) {
                    selmask |= m & (0u - m);
                    m = __shfl_sync(FULL, after, __clz(__brev(m)));
                }
                // end of the last selected match (0 if there is none)
                const uint32_t lastsel = __shfl_sync(FULL, endl, (31 - __clz(selmask)) & 31);
                const uint32_t cur = selmask ? lastsel : 0u;
                // ... and every lane then decides for itself whether a selected match covers it: the
                // nearest selected lane at or below it is the only one that can (matches do not overlap)
                const uint32_t below = selmask & (FULL >> (31 - lane));
                const uint32_t owner = below ? 31 - __clz(below) : lane;
                const uint32_t oend = __shfl_sync(FULL, endl, owner);
                const bool is_sel = (selmask >> lane) & 1;
                const bool is_lit = lane < valid && !(below && oend > lane);
                const uint32_t sel = __ballot_sync(FULL, is_sel || is_lit);
                if (is_sel || is_lit) {
                    const uint32_t rank = __popc(sel & ((1u << lane) - 1u));
                    if (is_sel) {
                        mytok[nt + rank] = tok_match(len, dist);
                        uint32_t idx, ne, ev, ds;
                        len_symbol(len, idx, ne, ev);
                        hist_add(hs, 257 + idx);
                        dist_symbol(dist, ds, ne, ev);
                        hist_add(hs, NLIT + ds);
                    } else {
                        mytok[nt + rank] = byte;
                        hist_add(hs, byte);
                    }
                }
                nt += __popc(sel);
                pos += max(valid, cur);
            }
        }
        __syncwarp();
        if (lane == 0) ntok[chunk * NSEG + warp] = nt;
        hist_store(hs, hist + (size_t)(chunk * NSEG + warp) * NSYM, lane);
    }
}

// =====================================================================================================
// "better" level (reference level 3, LZ77::getMatchesSlow, include/deflate.hpp:268-304).
//
// The reference scans ALL earlier positions of a 32 KB chunk for every position (O(n^2), ~1.1 s per
// chunk) and takes the longest match greedily.  Here one CTA owns a 64 KiB chunk for its whole
// lifetime in ~224 KB of shared memory:
//   phase A  warp 0 threads the chunk into hash chains (3-byte hash, 8 K u32 heads + a 64 K-entry u16
//            `prev` array in shared memory); 32 positions per step push themselves onto their bucket
//            with atomicExch (__match_any_sync, the first version, costs ~700 cycles per step when the
//            32 hashes are distinct -- which they normally are)
//   phase B  all 16 warps search: each lane walks the chain of its own position (nearest first, up to
//            `depth` candidates, 32 KiB distance limit, 4-byte-stride extension) and stores its best
//            (length, distance) in the chunk's token scratch
//   phase C  every warp parses one 4 KiB segment over those candidates with lazy evaluation
//            (a match is deferred when the next position has a longer one), emit tokens in place and
//            build the histograms, exactly as the fast kernel does.
// =====================================================================================================
constexpr uint32_t LZB_THREADS = 1024;          // 32 warps: the chain search is a chase of dependent shared-memory loads, one CTA per SM
                                                // (224 KB) -- twice the warps hide twice the latency; 16 of them parse afterwards
constexpr uint32_t LZB_HASH_BITS = 13;          // 8 K heads x u32 (atomicExch needs 32-bit words)
constexpr uint32_t LZB_NIL = 0xFFFFu;           // position 65535 can never be anybody's predecessor
constexpr uint32_t LZB_GOOD = 32;               // once a match this long is in hand, cut the remaining search to a quarter
constexpr uint32_t LZB_TOO_FAR = 4096;          // a 3-byte match farther than this costs more than literals
constexpr size_t LZB_SMEM_BYTES = CHUNK + LZ_DATA_PAD + CHUNK * 2 + (4u << LZB_HASH_BITS) + 16 + 80;      // + progress words

__device__ __forceinline__ uint32_t lzb_hash(uint32_t w4) { return ((w4 & 0xFFFFFFu) * 0x9E3779B1u) >> (32 - LZB_HASH_BITS); }

__global__ void __launch_bounds__(LZB_THREADS, 1)
lz77_better_kernel(const uint8_t* __restrict__ in, uint64_t n, uint32_t* __restrict__ tok,
                   uint32_t* __restrict__ ntok, uint16_t* __restrict__ hist, uint32_t depth_large, uint32_t nice_large,
                   uint32_t depth_small, uint32_t nice_small, uint32_t input_is_small, const ChunkSrc* __restrict__ srcs) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* s_data = smem;
    uint16_t* s_prev = reinterpret_cast<uint16_t*>(smem + CHUNK + LZ_DATA_PAD);
    uint32_t* s_head = reinterpret_cast<uint32_t*>(s_prev + CHUNK);
    uint32_t* s_hist = s_head;                                       // phase C reuses the head table
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + CHUNK + LZ_DATA_PAD + CHUNK * 2 + (4u << LZB_HASH_BITS));

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t chunk = blockIdx.x;
    const uint64_t base = srcs ? srcs[chunk].off : chunk * CHUNK;
    const uint32_t clen = srcs ? srcs[chunk].clen : (uint32_t)min((uint64_t)CHUNK, n - base);
    const uint8_t* src = in + base;
    const uint32_t FULL = 0xFFFFFFFFu;
    // search effort by the size of the INPUT this chunk belongs to (a batch mixes small and large files: ChunkSrc says which)
    const bool small_input = srcs ? (srcs[chunk].last & 2u) != 0 : input_is_small != 0;
    const uint32_t depth = small_input ? depth_small : depth_large;
    const uint32_t nice = small_input ? nice_small : nice_large;

    const bool aligned = ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    const uint32_t bulk = aligned ? (clen & ~15u) : 0;
    uint32_t* s_sync = reinterpret_cast<uint32_t*>(s_bar + 1);       // [0] positions linked so far, [1] next search tile
    uint32_t* s_segdone = s_sync + 4;                                // [NSEG] search tiles finished per segment; s_sync[2] = next segment to parse
    if (tid == 0) mbar_init(s_bar, 1);
    if (tid < 4 + NSEG) s_sync[tid] = 0;
    __syncthreads();
    if (tid == 0 && bulk) tma_load_1d(s_data, src, bulk, s_bar);
    for (uint32_t i = bulk + tid; i < clen; i += LZB_THREADS) s_data[i] = src[i];
    for (uint32_t i = clen + tid; i < ((clen + 15u) & ~15u) + LZ_DATA_PAD && i < CHUNK + LZ_DATA_PAD; i += LZB_THREADS)
        s_data[i] = 0;
    for (uint32_t i = tid; i < (1u << LZB_HASH_BITS); i += LZB_THREADS) s_head[i] = LZB_NIL;
    if (bulk) mbar_wait(s_bar, 0);
    __syncthreads();

    // ---- phase A: hash chains (one warp) -----------------------------------------------------------
    // 32 positions per step push themselves onto their bucket with atomicExch: every position gets the
    // previous head as its predecessor, so a bucket is one linked list through all its positions,
    // newest first.  Steps are ordered; inside a step the hardware serialises lanes that hit the same
    // bucket (the search below tolerates either order by skipping predecessors that are not earlier).
    // Four steps are in flight at a time: a warp's atomics on one address execute in program order, so the chains are
    // the same as with one step at a time, but the ~150-cycle round trip of the exchange is paid once per four steps.
    if (warp == 0) {
        constexpr uint32_t INFL = 8;          // steps in flight (the linker is the critical path once the search runs beside it)
        for (uint32_t t0 = 0; t0 < clen; t0 += 32 * INFL) {
            uint32_t old[INFL], hsh[INFL];
            #pragma unroll
            for (uint32_t k = 0; k < INFL; k++) hsh[k] = lzb_hash(ld4_unaligned(s_data, t0 + k * 32 + lane));   // (the pad behind the chunk is readable)
            #pragma unroll
            for (uint32_t k = 0; k < INFL; k++) {
                const uint32_t p = t0 + k * 32 + lane;
                old[k] = LZB_NIL;
                if (p + 3 <= clen) old[k] = atomicExch(&s_head[hsh[k]], p);
            }
            #pragma unroll
            for (uint32_t k = 0; k < INFL; k++) {
                const uint32_t p = t0 + k * 32 + lane;
                if (p < clen) s_prev[p] = (uint16_t)old[k];
            }
            // publish: the searchers of phase B run right behind (see below)
            __syncwarp();
            if (lane == 0) { __threadfence_block(); *reinterpret_cast<volatile uint32_t*>(&s_sync[0]) = t0 + 32 * INFL; }
        }
    }

    // ---- phase B: chain search, every position, all warps ----------------------------------------
    // The search at position p follows links of positions <= p only, so it does not have to wait for the whole chunk to be
    // linked: the other 31 warps pull 32-position tiles in order from a counter and wait (rarely: linking is twice as fast
    // as searching) until the linker has passed their tile; warp 0 joins when it is done.  Before, 31 warps sat at a barrier
    // for the 30 % of the kernel that phase A takes (ncu r02: 29.6 % of all stall samples on that barrier).  Chains and
    // search are what they were, so is every output byte.
    uint32_t* cand = tok + chunk * CHUNK;
    for (;;) {
        uint32_t t0 = 0;
        if (lane == 0) t0 = atomicAdd(&s_sync[1], 1u) * 32u;
        t0 = __shfl_sync(FULL, t0, 0);
        if (t0 >= clen) break;
        {
            const uint32_t need = min(t0 + 32u, clen);
            while (*reinterpret_cast<volatile uint32_t*>(&s_sync[0]) < need) __nanosleep(256);
            __threadfence_block();
        }
        const uint32_t p = t0 + lane;
        uint32_t best = 0, bdist = 0;
        if (p + 3 <= clen) {
            const uint32_t maxl = min(clen - p, MAX_MATCH);
            const uint32_t stop = min(nice, maxl);
            const uint32_t w4 = ld4_unaligned(s_data, p);
            uint32_t q = s_prev[p];
            uint32_t budget = depth;
            best = 2;
            while (q != LZB_NIL && budget--) {
                if (q >= p) { q = s_prev[q]; continue; }              // same-step neighbour linked the other way round
                if (p - q > MAX_DIST) break;
                // cheap reject: to beat `best` a candidate must agree on bytes [best-3, best] and at the head
                if ((best < 3 || ld4_unaligned(s_data, q + best - 3) == ld4_unaligned(s_data, p + best - 3)) &&
                    ((ld4_unaligned(s_data, q) ^ w4) & 0xFFFFFFu) == 0) {
                    uint32_t l = 3;
                    while (l < maxl) {
                        const uint32_t x = ld4_unaligned(s_data, p + l) ^ ld4_unaligned(s_data, q + l);
                        if (x) { l += (__ffs(x) - 1) >> 3; break; }
                        l += 4;
                    }
                    l = min(l, maxl);
                    if (l > best) {
                        best = l; bdist = p - q;
                        if (l >= stop) break;
                        if (l >= LZB_GOOD && budget > depth / 4) budget = depth / 4;   // good enough: search less (zlib's good_length)
                    }
                }
                q = s_prev[q];
            }
            if (best < 3 || (best == 3 && bdist > LZB_TOO_FAR)) { best = 0; bdist = 0; }
        }
        if (p < clen) cand[p] = best | (bdist << 16);
        // this tile's candidates are in place: count it for its segment (the parser of that segment waits for all of them)
        __syncwarp();
        if (lane == 0) { __threadfence_block(); atomicAdd(&s_segdone[t0 / SEG], 1u); }
    }

    // ---- phase C: lazy parse, one warp per segment -----------------------------------------------
    // No barrier between the phases either: a warp that finds no search tile left takes the next segment, waits until the
    // segment's tiles are all counted (and the linker is done with the head table, whose memory the histograms reuse) and
    // parses it while other warps are still searching later segments.
    for (;;) {
    uint32_t seg = 0;
    if (lane == 0) seg = atomicAdd(&s_sync[2], 1u);
    seg = __shfl_sync(FULL, seg, 0);
    if (seg >= NSEG) break;
    const uint32_t seg_lo = seg * SEG;
    const uint32_t seg_hi = min(clen, seg_lo + SEG);
    {
        const uint32_t tiles = seg_lo < clen ? (seg_hi - seg_lo + 31u) / 32u : 0u;
        while (*reinterpret_cast<volatile uint32_t*>(&s_segdone[seg]) < tiles ||
               *reinterpret_cast<volatile uint32_t*>(&s_sync[0]) < clen) __nanosleep(64);
        __threadfence_block();
    }
    uint32_t* h = s_hist + seg * HIST_WORDS;
    for (uint32_t i = lane; i < HIST_WORDS; i += 32) h[i] = 0;
    __syncwarp();
    uint32_t* mytok = tok + chunk * CHUNK + seg_lo;
    uint32_t nt = 0;
    if (seg_lo < clen) {
        uint32_t pos = seg_lo;
        while (pos < seg_hi) {
            const uint32_t p = pos + lane;
            const uint32_t valid = min(32u, seg_hi - pos);
            uint32_t len = 0, dist = 0;
            if (lane < valid) {
                const uint32_t c = cand[p];
                len = min(c & 0xFFFFu, seg_hi - p);          // tokens never cross a segment
                dist = c >> 16;
                if (len < 3 || (len == 3 && dist > LZB_TOO_FAR)) len = 0;
            }
            const uint8_t byte = s_data[p];
            __syncwarp();                                     // all candidate loads done before tokens overwrite them
            // lane 31 is only a lookahead unless the segment ends inside this window
            const uint32_t limit = valid < 32 ? valid : 31;
            const uint32_t mmask = __ballot_sync(FULL, len >= 3);
            uint32_t litmask = 0, selmask = 0, cur = 0;
            while (cur < limit) {
                const uint32_t m = mmask & ~((1u << cur) - 1u) & ((limit >= 32) ? FULL : ((1u << limit) - 1u));
                if (m == 0) { litmask |= mask_range(cur, limit); cur = limit; break; }
                const uint32_t j = __ffs(m) - 1;
                litmask |= mask_range(cur, j);
                const uint32_t lj = __shfl_sync(FULL, len, j);
                const uint32_t ln = __shfl_sync(FULL, len, (j + 1) & 31);
                if (j + 1 < valid && ln > lj) { litmask |= 1u << j; cur = j + 1; continue; }   // lazy: defer
                selmask |= 1u << j;
                cur = j + lj;
            }
            const uint32_t sel = litmask | selmask;
            if ((sel >> lane) & 1) {
                const uint32_t rank = __popc(sel & ((1u << lane) - 1u));
                if ((selmask >> lane) & 1) {
                    mytok[nt + rank] = tok_match(len, dist);
                    uint32_t idx, ne, ev, ds;
                    len_symbol(len, idx, ne, ev);
                    hist_add(h, 257 + idx);
                    dist_symbol(dist, ds, ne, ev);
                    hist_add(h, NLIT + ds);
                } else {
                    mytok[nt + rank] = byte;
                    hist_add(h, byte);
                }
            }
            nt += __popc(sel);
            pos += cur;
            __syncwarp();
        }
    }
    __syncwarp();
    if (lane == 0) ntok[chunk * NSEG + seg] = nt;
    hist_store(h, hist + (size_t)(chunk * NSEG + seg) * NSYM, lane);
    }
}

}  // namespace b200
