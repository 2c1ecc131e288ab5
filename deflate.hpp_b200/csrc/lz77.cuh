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

}  // namespace b200
