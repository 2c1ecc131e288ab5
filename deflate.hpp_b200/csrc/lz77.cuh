// lz77.cuh -- K1: per-chunk LZ77 tokeniser + histogram.
//
// Replaces the reference's LZ77::getMatches (include/deflate.hpp:310-383, "fast") and the histogram
// half of constructDynamicHuffmanTree (deflate.hpp:402-418).  Not a port: the reference probes only
// 4-byte-aligned positions of a 32 KB chunk with one thread and a first-occurrence table
// (deflate.hpp:373-376); here one CTA owns a 64 KiB chunk staged in shared memory by a TMA bulk copy,
// each of its 8 warps parses an 8 KiB segment with its own shared-memory hash table, the 32 lanes of
// a warp probe 32 consecutive positions at once, and the greedy selection is a warp-uniform loop
// over ballot masks.  Tokens go to a global scratch buffer (one word per TOKEN, not per byte) and the
// literal/length + distance histograms are accumulated with shared-memory atomics per warp.
#pragma once
#include "common.cuh"

namespace b200 {

constexpr uint32_t LZ_THREADS = NSEG * 32;     // 256
constexpr uint32_t LZ_HASH_BITS = 11;          // 2048 x u16 per warp
constexpr uint32_t LZ_WARM = 4096;             // bytes of the previous segment pre-inserted into the table
constexpr uint32_t LZ_DATA_PAD = 64;           // zeroed over-read slack after the chunk

constexpr size_t LZ_SMEM_BYTES = CHUNK + LZ_DATA_PAD + NSEG * (2u << LZ_HASH_BITS) + NSEG * NSYM * 4 + 16;

__device__ __forceinline__ uint32_t lz_hash(uint32_t w4) { return (w4 * 0x9E3779B1u) >> (32 - LZ_HASH_BITS); }

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

// MODE 0: greedy hash matcher (level 2 "fast").  MODE 1: literals only (level 1).
// grid = chunks, block = 256.  first_chunk lets a batch address its slice of the scratch buffers.
template <int MODE>
__global__ void __launch_bounds__(LZ_THREADS, 2)
lz77_kernel(const uint8_t* __restrict__ in, uint64_t n, uint32_t* __restrict__ tok,
            uint32_t* __restrict__ ntok, uint32_t* __restrict__ hist) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* s_data = smem;                                                  // CHUNK + pad
    uint16_t* s_tab = reinterpret_cast<uint16_t*>(smem + CHUNK + LZ_DATA_PAD);
    uint32_t* s_hist = reinterpret_cast<uint32_t*>(smem + CHUNK + LZ_DATA_PAD + NSEG * (2u << LZ_HASH_BITS));
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_hist + NSEG * NSYM);

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t chunk = blockIdx.x;
    const uint64_t base = chunk * CHUNK;
    const uint32_t clen = (uint32_t)min((uint64_t)CHUNK, n - base);
    const uint8_t* src = in + base;

    // ---- stage the chunk in shared memory (TMA bulk copy for the 16-byte-aligned body) -------
    const bool aligned = ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    const uint32_t bulk = aligned ? (clen & ~15u) : 0;
    if (tid == 0) mbar_init(s_bar, 1);
    __syncthreads();
    if (tid == 0 && bulk) tma_load_1d(s_data, src, bulk, s_bar);
    for (uint32_t i = bulk + tid; i < clen; i += LZ_THREADS) s_data[i] = src[i];
    for (uint32_t i = clen + tid; i < ((clen + 15u) & ~15u) + LZ_DATA_PAD && i < CHUNK + LZ_DATA_PAD; i += LZ_THREADS)
        s_data[i] = 0;
    {   // zero hash tables + histograms while the copy is in flight
        uint32_t* z = reinterpret_cast<uint32_t*>(s_tab);
        const uint32_t nz = (NSEG * (2u << LZ_HASH_BITS) + NSEG * NSYM * 4) / 4;
        for (uint32_t i = tid; i < nz; i += LZ_THREADS) z[i] = 0;
    }
    if (bulk) mbar_wait(s_bar, 0);
    __syncthreads();

    // ---- per-warp parse of one segment ----------------------------------------------------
    const uint32_t seg_lo = warp * SEG;
    const uint32_t seg_hi = min(clen, seg_lo + SEG);
    uint16_t* tab = s_tab + warp * (1u << LZ_HASH_BITS);
    uint32_t* h = s_hist + warp * NSYM;
    uint32_t* mytok = tok + chunk * CHUNK + seg_lo;
    uint32_t nt = 0;
    const uint32_t FULL = 0xFFFFFFFFu;

    if (seg_lo < clen) {
        if (MODE == 0) {
            // warm start: pre-insert the tail of the previous segment so matches can reach into it
            uint32_t ws = seg_lo > LZ_WARM ? seg_lo - LZ_WARM : 0;
            for (uint32_t p = ws + lane; p < seg_lo; p += 32) tab[lz_hash(ld4_unaligned(s_data, p))] = (uint16_t)p;
            __syncwarp();
        }
        uint32_t pos = seg_lo;
        while (pos < seg_hi) {
            const uint32_t p = pos + lane;
            const uint32_t avail = p < seg_hi ? seg_hi - p : 0;
            const uint32_t w4 = ld4_unaligned(s_data, p);
            uint32_t len = 0, dist = 0;
            if (MODE == 0) {
                const uint32_t hsh = lz_hash(w4);
                uint32_t q = tab[hsh];
                __syncwarp();
                if (avail >= 4) tab[hsh] = (uint16_t)p;
                __syncwarp();
                bool hit = avail >= 4 && q < p && (p - q) <= MAX_DIST && ld4_unaligned(s_data, q) == w4;
                if (!hit && avail >= 4 && p > 0 && ld4_unaligned(s_data, p - 1) == w4) { q = p - 1; hit = true; }  // run
                if (hit) {
                    const uint32_t maxl = min(avail, MAX_MATCH);
                    uint32_t l = 4;
                    while (l < maxl) {
                        uint32_t x = ld4_unaligned(s_data, p + l) ^ ld4_unaligned(s_data, q + l);
                        if (x) { l += (__ffs(x) - 1) >> 3; break; }
                        l += 4;
                    }
                    len = min(l, maxl);
                    dist = p - q;
                }
            }
            // greedy in-order selection over the 32 candidates: warp-uniform mask arithmetic
            const uint32_t valid = min(32u, seg_hi - pos);
            uint32_t litmask = 0, selmask = 0, advance = valid;
            if (MODE == 0) {
                const uint32_t mmask = __ballot_sync(FULL, len >= 4);
                uint32_t cur = 0;
                for (;;) {
                    uint32_t m = cur < 32 ? (mmask & ~((1u << cur) - 1u)) : 0;
                    if (m == 0) {
                        if (cur < valid) litmask |= mask_range(cur, valid);
                        advance = max(valid, cur);
                        break;
                    }
                    uint32_t j = __ffs(m) - 1;
                    litmask |= mask_range(cur, j);
                    selmask |= 1u << j;
                    cur = j + __shfl_sync(FULL, len, j);
                    if (cur >= 32) { advance = cur; break; }
                }
            } else {
                litmask = mask_range(0, valid);
            }
            const uint32_t sel = litmask | selmask;
            if ((sel >> lane) & 1) {
                const uint32_t rank = __popc(sel & ((1u << lane) - 1u));
                if ((selmask >> lane) & 1) {
                    mytok[nt + rank] = tok_match(len, dist);
                    uint32_t idx, ne, ev, ds;
                    len_symbol(len, idx, ne, ev);
                    atomicAdd(&h[257 + idx], 1u);
                    dist_symbol(dist, ds, ne, ev);
                    atomicAdd(&h[NLIT + ds], 1u);
                } else {
                    mytok[nt + rank] = w4 & 0xFFu;
                    atomicAdd(&h[w4 & 0xFFu], 1u);
                }
            }
            nt += __popc(sel);
            pos += advance;
        }
    }
    __syncwarp();
    if (lane == 0) ntok[chunk * NSEG + warp] = nt;
    uint32_t* gh = hist + (chunk * NSEG + warp) * NSYM;
    for (uint32_t i = lane; i < NSYM; i += 32) gh[i] = h[i];
}

// =====================================================================================================
// "fast" level, two-phase (replaces lz77_kernel<0> as the default level-2 matcher).
//
//   phase B  all 8 warps sweep the chunk in tiles of 256 positions against ONE chunk-wide hash table
//            (8 K x u32 in shared memory, 4-byte hash).  Per tile: every thread looks its bucket up,
//            __syncthreads, every thread inserts its position with atomicMax -- lookups therefore see
//            exactly the positions before the tile, and the table content is deterministic (highest
//            position wins).  Three candidates per position: the bucket, the nearest lower lane of the
//            same warp with the same hash (__match_any_sync: matches closer than a tile), and p-1 (runs).
//            The longest verified candidate, extended to at most LZF_CAP bytes, goes to the token
//            scratch as (len | dist << 16).
//   phase C  each warp parses one 8 KiB segment greedily over those candidates, 32 positions per step
//            with ballot-mask selection; a selected match that hit the cap is extended to its full
//            length (<= 258) by the whole warp, 128 bytes per step.
// Against the single-phase kernel: the whole 32 KiB window is reachable from every position instead of
// an 8 KiB segment-private table, ~40 % fewer tokens on the benchmark corpus (less work for K4 and for
// the inflater), and no position is probed twice.
// =====================================================================================================
constexpr uint32_t LZF_THREADS = 256;
constexpr uint32_t LZF_HASH_BITS = 13;
constexpr uint32_t LZF_CAP = 36;                 // phase-B extension cap (multiple of 4)
constexpr size_t LZF_SMEM_BYTES = CHUNK + LZ_DATA_PAD + (4u << LZF_HASH_BITS) + NSEG * NSYM * 4 + 16;

__device__ __forceinline__ uint32_t lzf_hash(uint32_t w4) { return (w4 * 0x9E3779B1u) >> (32 - LZF_HASH_BITS); }

// length of the common prefix of positions p and q (q < p), known to agree on 4 bytes, capped
__device__ __forceinline__ uint32_t lzf_extend(const uint8_t* d, uint32_t p, uint32_t q, uint32_t maxl) {
    uint32_t l = 4;
    while (l < maxl) {
        const uint32_t x = ld4_unaligned(d, p + l) ^ ld4_unaligned(d, q + l);
        if (x) { l += (__ffs(x) - 1) >> 3; break; }
        l += 4;
    }
    return min(l, maxl);
}

__global__ void __launch_bounds__(LZF_THREADS, 2)
lz77_fast_kernel(const uint8_t* __restrict__ in, uint64_t n, uint32_t* __restrict__ tok,
                 uint32_t* __restrict__ ntok, uint32_t* __restrict__ hist) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* s_data = smem;
    uint32_t* s_tab = reinterpret_cast<uint32_t*>(smem + CHUNK + LZ_DATA_PAD);
    uint32_t* s_hist = s_tab + (1u << LZF_HASH_BITS);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_hist + NSEG * NSYM);

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t chunk = blockIdx.x;
    const uint64_t base = chunk * CHUNK;
    const uint32_t clen = (uint32_t)min((uint64_t)CHUNK, n - base);
    const uint8_t* src = in + base;
    const uint32_t FULL = 0xFFFFFFFFu;

    const bool aligned = ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    const uint32_t bulk = aligned ? (clen & ~15u) : 0;
    if (tid == 0) mbar_init(s_bar, 1);
    __syncthreads();
    if (tid == 0 && bulk) tma_load_1d(s_data, src, bulk, s_bar);
    for (uint32_t i = bulk + tid; i < clen; i += LZF_THREADS) s_data[i] = src[i];
    for (uint32_t i = clen + tid; i < ((clen + 15u) & ~15u) + LZ_DATA_PAD && i < CHUNK + LZ_DATA_PAD; i += LZF_THREADS)
        s_data[i] = 0;
    for (uint32_t i = tid; i < (1u << LZF_HASH_BITS) + NSEG * NSYM; i += LZF_THREADS) s_tab[i] = 0;
    if (bulk) mbar_wait(s_bar, 0);
    __syncthreads();

    // ---- phase B: candidates for every position ---------------------------------------------------
    uint32_t* cand = tok + chunk * CHUNK;
    for (uint32_t t0 = 0; t0 < clen; t0 += LZF_THREADS) {
        const uint32_t p = t0 + tid;
        const bool valid = p + 4 <= clen;
        const uint32_t w4 = ld4_unaligned(s_data, p);
        const uint32_t h = lzf_hash(w4);
        const uint32_t q_tab = s_tab[h];
        const uint32_t peers = __match_any_sync(FULL, valid ? h : (0x80000000u | lane));
        const uint32_t lower = peers & ((1u << lane) - 1u);
        __syncthreads();                                  // every lookup of this tile is done
        if (valid) atomicMax(&s_tab[h], p);
        uint32_t best = 0, bdist = 0;
        if (valid) {
            const uint32_t maxl = min(min(clen - p, MAX_MATCH), LZF_CAP);
            if (q_tab < p && p - q_tab <= MAX_DIST && ld4_unaligned(s_data, q_tab) == w4) {
                best = lzf_extend(s_data, p, q_tab, maxl);
                bdist = p - q_tab;
            }
            if (lower && best < maxl) {
                const uint32_t q = p - lane + (31 - __clz(lower));
                if (ld4_unaligned(s_data, q) == w4) {
                    const uint32_t l = lzf_extend(s_data, p, q, maxl);
                    if (l >= best) { best = l; bdist = p - q; }           // ties: the nearer one
                }
            }
            if (p > 0 && best < maxl && bdist != 1 && ld4_unaligned(s_data, p - 1) == w4) {
                const uint32_t l = lzf_extend(s_data, p, p - 1, maxl);
                if (l >= best) { best = l; bdist = 1; }
            }
        }
        if (p < clen) cand[p] = best | (bdist << 16);
        __syncthreads();                                  // inserts land before the next tile looks up
    }
    // the block-wide barrier above also orders this CTA's cand[] stores before the loads below

    // ---- phase C: greedy parse, one warp per segment ----------------------------------------------
    const uint32_t seg_lo = warp * SEG;
    const uint32_t seg_hi = min(clen, seg_lo + SEG);
    uint32_t* hs = s_hist + warp * NSYM;
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
                len = min(c & 0xFFFFu, seg_hi - p);           // tokens never cross a segment
                dist = c >> 16;
                if (len < 4) len = 0;
            }
            const uint32_t byte = s_data[p];
            __syncwarp();                                     // candidate loads done before tokens overwrite them
            const uint32_t mmask = __ballot_sync(FULL, len >= 4);
            uint32_t litmask = 0, selmask = 0, cur = 0, advance = valid;
            for (;;) {
                const uint32_t m = cur < 32 ? (mmask & ~((1u << cur) - 1u)) : 0;
                if (m == 0) {
                    if (cur < valid) litmask |= mask_range(cur, valid);
                    advance = max(valid, cur);
                    break;
                }
                const uint32_t j = __ffs(m) - 1;
                litmask |= mask_range(cur, j);
                selmask |= 1u << j;
                uint32_t lj = __shfl_sync(FULL, len, j);
                if (lj == LZF_CAP) {
                    // the candidate hit the phase-B cap: extend it, 128 bytes per warp step
                    const uint32_t dj = __shfl_sync(FULL, dist, j);
                    const uint32_t pj = pos + j;
                    const uint32_t maxl = min(seg_hi - pj, MAX_MATCH);
                    while (lj < maxl) {
                        const uint32_t o = lj + 4 * lane;
                        const uint32_t x = ld4_unaligned(s_data, pj + o) ^ ld4_unaligned(s_data, pj - dj + o);
                        const uint32_t mm = __ballot_sync(FULL, x != 0);
                        if (mm) {
                            const uint32_t f = __ffs(mm) - 1;
                            const uint32_t xb = __shfl_sync(FULL, x, f);
                            lj += 4 * f + ((__ffs(xb) - 1) >> 3);
                            break;
                        }
                        lj += 128;
                    }
                    lj = min(lj, maxl);
                    if (lane == j) len = lj;
                }
                cur = j + lj;
                if (cur >= 32) { advance = cur; break; }
            }
            const uint32_t sel = litmask | selmask;
            if ((sel >> lane) & 1) {
                const uint32_t rank = __popc(sel & ((1u << lane) - 1u));
                if ((selmask >> lane) & 1) {
                    mytok[nt + rank] = tok_match(len, dist);
                    uint32_t idx, ne, ev, ds;
                    len_symbol(len, idx, ne, ev);
                    atomicAdd(&hs[257 + idx], 1u);
                    dist_symbol(dist, ds, ne, ev);
                    atomicAdd(&hs[NLIT + ds], 1u);
                } else {
                    mytok[nt + rank] = byte;
                    atomicAdd(&hs[byte], 1u);
                }
            }
            nt += __popc(sel);
            pos += advance;
            __syncwarp();
        }
    }
    __syncwarp();
    if (lane == 0) ntok[chunk * NSEG + warp] = nt;
    uint32_t* gh = hist + (chunk * NSEG + warp) * NSYM;
    for (uint32_t i = lane; i < NSYM; i += 32) gh[i] = hs[i];
}

// =====================================================================================================
// "better" level (reference level 3, LZ77::getMatchesSlow, include/deflate.hpp:268-304).
//
// The reference scans ALL earlier positions of a 32 KB chunk for every position (O(n^2), ~1.1 s per
// chunk) and takes the longest match greedily.  Here one CTA owns a 64 KiB chunk for its whole
// lifetime in ~224 KB of shared memory:
//   phase A  warp 0 threads the chunk into hash chains (3-byte hash, 16 K heads + a 64 K-entry `prev`
//            array, both u16 in shared memory); 32 positions per step, intra-step collisions resolved
//            with __match_any_sync so the chains are strictly ordered by position
//   phase B  all 16 warps search: each lane walks the chain of its own position (nearest first, up to
//            `depth` candidates, 32 KiB distance limit, 4-byte-stride extension) and stores its best
//            (length, distance) in the chunk's token scratch
//   phase C  warps 0..7 parse one 8 KiB segment each over those candidates with lazy evaluation
//            (a match is deferred when the next position has a longer one), emit tokens in place and
//            build the histograms, exactly as the fast kernel does.
// =====================================================================================================
constexpr uint32_t LZB_THREADS = 512;
constexpr uint32_t LZB_HASH_BITS = 14;
constexpr uint32_t LZB_NIL = 0xFFFFu;           // position 65535 can never be anybody's predecessor
constexpr uint32_t LZB_TOO_FAR = 4096;          // a 3-byte match farther than this costs more than literals
constexpr size_t LZB_SMEM_BYTES = CHUNK + LZ_DATA_PAD + CHUNK * 2 + (2u << LZB_HASH_BITS) + 16;

__device__ __forceinline__ uint32_t lzb_hash(uint32_t w4) { return ((w4 & 0xFFFFFFu) * 0x9E3779B1u) >> (32 - LZB_HASH_BITS); }

__global__ void __launch_bounds__(LZB_THREADS, 1)
lz77_better_kernel(const uint8_t* __restrict__ in, uint64_t n, uint32_t* __restrict__ tok,
                   uint32_t* __restrict__ ntok, uint32_t* __restrict__ hist, uint32_t depth, uint32_t nice) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* s_data = smem;
    uint16_t* s_prev = reinterpret_cast<uint16_t*>(smem + CHUNK + LZ_DATA_PAD);
    uint16_t* s_head = s_prev + CHUNK;
    uint32_t* s_hist = reinterpret_cast<uint32_t*>(s_head);          // phase C reuses the head table
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + CHUNK + LZ_DATA_PAD + CHUNK * 2 + (2u << LZB_HASH_BITS));

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t chunk = blockIdx.x;
    const uint64_t base = chunk * CHUNK;
    const uint32_t clen = (uint32_t)min((uint64_t)CHUNK, n - base);
    const uint8_t* src = in + base;
    const uint32_t FULL = 0xFFFFFFFFu;

    const bool aligned = ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    const uint32_t bulk = aligned ? (clen & ~15u) : 0;
    if (tid == 0) mbar_init(s_bar, 1);
    __syncthreads();
    if (tid == 0 && bulk) tma_load_1d(s_data, src, bulk, s_bar);
    for (uint32_t i = bulk + tid; i < clen; i += LZB_THREADS) s_data[i] = src[i];
    for (uint32_t i = clen + tid; i < ((clen + 15u) & ~15u) + LZ_DATA_PAD && i < CHUNK + LZ_DATA_PAD; i += LZB_THREADS)
        s_data[i] = 0;
    for (uint32_t i = tid; i < (1u << LZB_HASH_BITS) / 2; i += LZB_THREADS) reinterpret_cast<uint32_t*>(s_head)[i] = 0xFFFFFFFFu;
    if (bulk) mbar_wait(s_bar, 0);
    __syncthreads();

    // ---- phase A: ordered hash chains (one warp) -------------------------------------------------
    if (warp == 0) {
        for (uint32_t t0 = 0; t0 < clen; t0 += 32) {
            const uint32_t p = t0 + lane;
            const bool valid = p + 3 <= clen;
            const uint32_t h = lzb_hash(ld4_unaligned(s_data, p));
            const uint32_t key = valid ? h : (0x80000000u | lane);       // invalid lanes match nobody
            const uint32_t peers = __match_any_sync(FULL, key);
            const uint32_t lower = peers & ((1u << lane) - 1u);
            if (valid) {
                const uint32_t old = s_head[h];
                s_prev[p] = (uint16_t)(lower ? t0 + (31 - __clz(lower)) : old);
            } else if (p < clen) {
                s_prev[p] = (uint16_t)LZB_NIL;
            }
            __syncwarp();
            if (valid && (peers >> lane) == 1u) s_head[h] = (uint16_t)p;  // highest lane of its group
            __syncwarp();
        }
    }
    __syncthreads();

    // ---- phase B: chain search, every position, all warps ----------------------------------------
    uint32_t* cand = tok + chunk * CHUNK;
    for (uint32_t t0 = warp * 32; t0 < clen; t0 += LZB_THREADS) {
        const uint32_t p = t0 + lane;
        uint32_t best = 0, bdist = 0;
        if (p + 3 <= clen) {
            const uint32_t maxl = min(clen - p, MAX_MATCH);
            const uint32_t stop = min(nice, maxl);
            const uint32_t w4 = ld4_unaligned(s_data, p);
            uint32_t q = s_prev[p];
            uint32_t budget = depth;
            best = 2;
            while (q != LZB_NIL && budget-- && p - q <= MAX_DIST) {
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
                    if (l > best) { best = l; bdist = p - q; if (l >= stop) break; }
                }
                q = s_prev[q];
            }
            if (best < 3 || (best == 3 && bdist > LZB_TOO_FAR)) { best = 0; bdist = 0; }
        }
        if (p < clen) cand[p] = best | (bdist << 16);
    }
    __syncthreads();   // also orders the global cand[] stores before phase C's loads (same CTA)

    // ---- phase C: lazy parse per segment (warps 0..7) --------------------------------------------
    for (uint32_t i = tid; i < NSEG * NSYM; i += LZB_THREADS) s_hist[i] = 0;
    __syncthreads();
    if (warp >= NSEG) return;
    const uint32_t seg_lo = warp * SEG;
    const uint32_t seg_hi = min(clen, seg_lo + SEG);
    uint32_t* h = s_hist + warp * NSYM;
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
                    atomicAdd(&h[257 + idx], 1u);
                    dist_symbol(dist, ds, ne, ev);
                    atomicAdd(&h[NLIT + ds], 1u);
                } else {
                    mytok[nt + rank] = byte;
                    atomicAdd(&h[byte], 1u);
                }
            }
            nt += __popc(sel);
            pos += cur;
            __syncwarp();
        }
    }
    __syncwarp();
    if (lane == 0) ntok[chunk * NSEG + warp] = nt;
    uint32_t* gh = hist + (chunk * NSEG + warp) * NSYM;
    for (uint32_t i = lane; i < NSYM; i += 32) gh[i] = h[i];
}

}  // namespace b200
