// encode.cuh -- K3 (chunk offset scan) and K4 (token -> bit packing, written straight to its final
// place in the output stream).
//
// Replaces the reference's compressBuffer (include/deflate.hpp:630-674), makeUncompressedBlock
// (:387-399), Bitstream::addBits (:97-116, one byte-loop per code) and Bitstream::copyBitstream
// (:143-150, a second byte-by-byte re-pack of every block to concatenate at bit granularity).
// Here chunk sizes are exact before encoding (K2), an exclusive scan gives every chunk its final
// byte offset, and one CTA per chunk packs the bits in shared memory -- each warp owns one segment
// whose starting bit offset K2 computed -- then copies the bytes out with 16-byte stores.
#pragma once
#include "common.cuh"

namespace b200 {

// ---- K3: exclusive scan of per-chunk byte sizes (single CTA) --------------------------------------
// offsets[i] = base + sum(sizes[0..i)), offsets[n] = base + total.  *total_out = offsets[n].
constexpr uint32_t SCAN_THREADS = 1024;
__global__ void __launch_bounds__(SCAN_THREADS)
scan_sizes_kernel(const uint32_t* __restrict__ sizes, uint32_t n, const uint64_t* __restrict__ base_ptr,
                  uint64_t* __restrict__ offsets, uint64_t* __restrict__ total_out) {
    __shared__ uint64_t s_warp[32];
    __shared__ uint64_t s_carry;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = base_ptr ? *base_ptr : 0;
    __syncthreads();
    const uint32_t per = (n + SCAN_THREADS - 1) / SCAN_THREADS;   // contiguous items per thread
    const uint32_t lo = min(n, tid * per), hi = min(n, lo + per);
    uint64_t sum = 0;
    for (uint32_t i = lo; i < hi; i++) sum += sizes[i];
    uint64_t incl = sum;
    for (int o = 1; o < 32; o <<= 1) {
        uint64_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if ((int)lane >= o) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint64_t w = s_warp[lane];
        uint64_t wi = w;
        for (int o = 1; o < 32; o <<= 1) {
            uint64_t v = __shfl_up_sync(0xFFFFFFFFu, wi, o);
            if ((int)lane >= o) wi += v;
        }
        s_warp[lane] = wi - w;   // exclusive
    }
    __syncthreads();
    uint64_t run = s_carry + s_warp[warp] + (incl - sum);
    for (uint32_t i = lo; i < hi; i++) { offsets[i] = run; run += sizes[i]; }
    if (tid == SCAN_THREADS - 1) { offsets[n] = run; if (total_out) *total_out = run; }
}

// ---- K4: encode ------------------------------------------------------------------------------------
constexpr uint32_t ENC_THREADS = NSEG * 32;
constexpr uint32_t ENC_STAGE_BYTES = CHUNK + 80;       // block bytes (<= stored size) + phase + marker
constexpr size_t ENC_SMEM_BYTES = ENC_STAGE_BYTES + 2 * NSYM * 4;     // staging image + the code tables of up to two blocks

// OR `nbits` (<= 48) of v into the staging bit array at absolute bit position `bit`.
__device__ __forceinline__ void stage_bits(uint32_t* stage, uint32_t bit, uint64_t v, uint32_t nbits) {
    (void)nbits;
    const uint32_t w = bit >> 5, ph = bit & 31;
    const uint64_t a = v << ph;
    const uint32_t lo = (uint32_t)a, mid = (uint32_t)(a >> 32);
    const uint32_t hi = ph ? (uint32_t)(v >> (64 - ph)) : 0;
    if (lo) atomicOr(&stage[w], lo);
    if (mid) atomicOr(&stage[w + 1], mid);
    if (hi) atomicOr(&stage[w + 2], hi);
}

// Copy n bytes from global memory to global memory (any alignment on either side), all threads of the CTA: destination-aligned
// 16-byte stores, the source read as aligned 32-bit words and realigned with funnel shifts, four vectors (20 loads) in flight
// per thread; the ragged ends go bytewise, so nothing outside [dst, dst + n) is written.  (The stored third of the corpus first
// went through a byte loop into the shared-memory staging image -- one 1-byte global load per thread and trip, 128 dependent
// trips per chunk: 35 % of this kernel's stall samples -- then through the image with vectors: 2.12 -> 1.58 ms per GiB.
// Straight to its place it needs no image, no barrier, and the CTA is gone as soon as its stores are issued.)
__device__ __forceinline__ void vec_copy(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, uint32_t n, uint32_t tid) {
    const uint32_t head = min(n, (16u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u);
    if (tid < head) dst[tid] = src[tid];
    const uint8_t* s = src + head;
    uint8_t* d = dst + head;
    const uint32_t m = n - head;
    const uint32_t k = (uint32_t)(reinterpret_cast<uintptr_t>(s) & 3u), sh = k * 8;
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(s - k);
    // a vector reads the words [4 j, 4 j + 4] when the source is not word aligned: all of them must lie inside the source
    const uint32_t nvec = k ? (m >= 20 ? (m - 4) / 16 : 0) : m / 16;
    for (uint32_t j0 = 0; j0 < nvec; j0 += 4 * ENC_THREADS) {
        uint32_t w[4][5];
        #pragma unroll
        for (uint32_t u = 0; u < 4; u++) {
            const uint32_t j = j0 + u * ENC_THREADS + tid;
            if (j < nvec) {
                #pragma unroll
                for (uint32_t q = 0; q < 4; q++) w[u][q] = __ldg(sw + 4 * j + q);
                w[u][4] = k ? __ldg(sw + 4 * j + 4) : 0u;
            }
        }
        #pragma unroll
        for (uint32_t u = 0; u < 4; u++) {
            const uint32_t j = j0 + u * ENC_THREADS + tid;
            if (j < nvec) {
                uint4 v;
                v.x = __funnelshift_r(w[u][0], w[u][1], sh);
                v.y = __funnelshift_r(w[u][1], w[u][2], sh);
                v.z = __funnelshift_r(w[u][2], w[u][3], sh);
                v.w = __funnelshift_r(w[u][3], w[u][4], sh);
                *reinterpret_cast<uint4*>(d + 16 * j) = v;
            }
        }
    }
    for (uint32_t i = 16 * nvec + tid; i < m; i += ENC_THREADS) d[i] = s[i];
}

// grid = chunks, block = 256 (warp s encodes segment s).
__global__ void __launch_bounds__(ENC_THREADS, 3)
encode_kernel(const uint8_t* __restrict__ in, const uint32_t* __restrict__ tok, const uint32_t* __restrict__ ntok,
              const uint32_t* __restrict__ codes, const uint32_t* __restrict__ hdr,
              const BlockDesc* __restrict__ desc, const uint64_t* __restrict__ offsets,
              const uint64_t* __restrict__ extra_base, uint8_t* __restrict__ out, const ChunkSrc* __restrict__ srcs) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t* stage = reinterpret_cast<uint32_t*>(smem);
    uint32_t* s_codes = reinterpret_cast<uint32_t*>(smem + ENC_STAGE_BYTES);

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t chunk = blockIdx.x;
    const BlockDesc& d = desc[chunk];
    // extra_base: where this shard starts inside `out` when `out` is the stream being assembled from several
    // GPUs (out may then be another GPU's memory, mapped over NVLink: the copy-out below IS the gather)
    const uint64_t dst = offsets[chunk] + (extra_base ? *extra_base : 0);
    const uint32_t phase = (uint32_t)(dst & 15);          // staging byte k <-> out[(dst & ~15) + k]
    const uint32_t nbytes = d.nbytes;
    const uint32_t stage_bytes = phase + nbytes;
    const uint32_t stage_words = (stage_bytes + 3) / 4;

    if (d.btype == 0) {
        // stored blocks: [hdr byte][LEN][NLEN][raw bytes], <= 65535 bytes each (deflate.hpp:387-399), written straight to
        // their place in the stream (no staging image, no barrier)
        uint8_t* ob = out + dst;
        const uint8_t* src = in + (srcs ? srcs[chunk].off : chunk * CHUNK);
        const uint32_t nblk = (d.clen + 65534u) / 65535u;
        uint32_t done = 0, o = 0;
        for (uint32_t b = 0; b < nblk; b++) {
            const uint32_t bl = min(65535u, d.clen - done);
            if (tid == 0) {
                ob[o] = (d.last && b == nblk - 1) ? 1 : 0;
                ob[o + 1] = bl & 0xFF; ob[o + 2] = bl >> 8;
                ob[o + 3] = (~bl) & 0xFF; ob[o + 4] = ((~bl) >> 8) & 0xFF;
            }
            vec_copy(ob + o + 5, src + done, bl, tid);
            o += 5 + bl; done += bl;
        }
        if (!d.last && tid == 0) {   // separator: two empty stored blocks
            ob[o] = 0; ob[o + 1] = 0; ob[o + 2] = 0; ob[o + 3] = 0xFF; ob[o + 4] = 0xFF;
            ob[o + 5] = 0; ob[o + 6] = 0; ob[o + 7] = 0; ob[o + 8] = 0xFF; ob[o + 9] = 0xFF;
        }
        return;
    }
    for (uint32_t i = tid; i < ((stage_words + 3) & ~3u) + 4; i += ENC_THREADS) stage[i] = 0;
    const uint32_t split = d.split_seg;                    // segments [split, 16) belong to the chunk's second block
    for (uint32_t i = tid; i < (split ? 2 * NSYM : NSYM); i += ENC_THREADS) s_codes[i] = codes[chunk * 2 * NSYM + i];
    __syncthreads();

    {
        // segment index: four empty stored blocks per segment whose padding bits carry the segments' bit lengths
        if (tid < d.index_bytes / 5) {
            const uint32_t w = tid >> 2;
            const uint32_t word = w == 0 ? (INDEX_MAGIC | ((d.index_bytes / INDEX_BYTES_PER_SEG - 1) << 10)) : d.seg_bitoff[w] - d.seg_bitoff[w - 1];
            const uint32_t nib = (word >> (4 * (tid & 3))) & 15u;
            stage_bits(stage, (phase + 5 * tid) * 8, (uint64_t)(0x80u | (nib << 3)) | (0xFFFFull << 24), 40);
        }
        const uint32_t bit0 = (phase + d.index_bytes) * 8;
        // header bits
        const uint32_t hw = (d.hdr_bits + 31) / 32;
        for (uint32_t i = tid; i < hw; i += ENC_THREADS) {
            uint32_t v = hdr[chunk * 2 * HDR_WORDS + i];
            const uint32_t rem = d.hdr_bits - i * 32;
            if (rem < 32) v &= (1u << rem) - 1u;
            stage_bits(stage, bit0 + i * 32, v, 32);
        }
        if (split) {
            const uint32_t hw2 = (d.hdr_bits2 + 31) / 32;
            for (uint32_t i = tid; i < hw2; i += ENC_THREADS) {
                uint32_t v = hdr[(chunk * 2 + 1) * HDR_WORDS + i];
                const uint32_t rem = d.hdr_bits2 - i * 32;
                if (rem < 32) v &= (1u << rem) - 1u;
                stage_bits(stage, bit0 + d.block2_bit + i * 32, v, 32);
            }
        }
        const uint32_t* cd = s_codes + ((split && warp >= split) ? NSYM : 0u);
        // payload: warp = segment; 32 tokens per step, warp scan of bit lengths
        const uint32_t nt_raw = d.clen ? ntok[chunk * NSEG + warp] : 0u;  // an empty input: header + end of block only
        const uint32_t nt = nt_raw & ~NTOK_LITERALS;
        const bool lit_only = (nt_raw & NTOK_LITERALS) != 0;               // no tokens were written: token i = input byte i
        const uint32_t* mytok = tok + chunk * CHUNK + warp * SEG;
        const uint8_t* mysrc = in + (srcs ? srcs[chunk].off : chunk * CHUNK) + warp * SEG;
        auto ld_tok = [&](uint32_t k) -> uint32_t { return lit_only ? (uint32_t)mysrc[k] : mytok[k]; };
        uint32_t bit = bit0 + d.seg_bitoff[warp];
        // the tokens of the step after the current one are always in flight (a step is ~60 instructions: shorter than
        // an L2 / DRAM round trip; ncu r01f: 13 % of this kernel's stall samples were long-scoreboard)
        uint32_t t_next = lane < nt ? ld_tok(lane) : 0u;
        uint32_t t_next2 = 32 + lane < nt ? ld_tok(32 + lane) : 0u;
        for (uint32_t b = 0; b < nt; b += 32) {
            const uint32_t i = b + lane;
            uint64_t v = 0;
            uint32_t nb = 0;
            const uint32_t t = t_next;
            t_next = t_next2;
            if (b + 64 + lane < nt) t_next2 = ld_tok(b + 64 + lane);
            if (i < nt) {
                const uint32_t dist = tok_dist(t);
                if (dist == 0) {
                    const uint32_t c = cd[t & 0xFFu];
                    v = c >> 8; nb = c & 0xFFu;
                } else {
                    uint32_t idx, ne, ev;
                    len_symbol(tok_len(t), idx, ne, ev);
                    uint32_t c = cd[257 + idx];
                    v = c >> 8; nb = c & 0xFFu;
                    v |= (uint64_t)ev << nb; nb += ne;
                    uint32_t ds;
                    dist_symbol(dist, ds, ne, ev);
                    c = cd[NLIT + ds];
                    v |= (uint64_t)(c >> 8) << nb; nb += c & 0xFFu;
                    v |= (uint64_t)ev << nb; nb += ne;
                }
            }
            uint32_t incl = nb;
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t u = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if ((int)lane >= o) incl += u;
            }
            if (nb) stage_bits(stage, bit + incl - nb, v, nb);
            bit += __shfl_sync(0xFFFFFFFFu, incl, 31);
        }
        if (tid == 0) {
            if (split) {                                   // end of the first block, right in front of the second header
                const uint32_t c1 = s_codes[256];
                stage_bits(stage, bit0 + d.block2_bit - (c1 & 0xFFu), c1 >> 8, c1 & 0xFFu);
            }
            const uint32_t c = s_codes[(split ? NSYM : 0u) + 256];
            const uint32_t eob_len = c & 0xFFu;
            stage_bits(stage, bit0 + d.total_bits - eob_len, c >> 8, eob_len);
            if (!d.last) {   // separator: 3 zero bits, pad, 00 00 FF FF, then 00 | 00 00 FF FF
                const uint32_t mb = phase + d.index_bytes + (d.total_bits + 3 + 7) / 8;
                stage_bits(stage, (mb + 2) * 8, 0xFFFFu, 16);   // atomic: may share a word with payload bits
                stage_bits(stage, (mb + 7) * 8, 0xFFFFu, 16);
            }
        }
    }
    __syncthreads();

    // ---- copy out: aligned 16-byte vectors in the middle, bytes at the ragged ends ------------
    uint8_t* gbase = out + (dst - phase);                 // 16-byte aligned iff out is
    const bool valigned = ((reinterpret_cast<uintptr_t>(gbase) & 15) == 0);
    const uint32_t vlo = valigned ? ((phase + 15) & ~15u) : stage_bytes;   // first fully-owned vector
    const uint32_t vhi = valigned ? max(vlo, stage_bytes & ~15u) : stage_bytes;
    for (uint32_t i = phase + tid; i < min(vlo, stage_bytes); i += ENC_THREADS) gbase[i] = smem[i];
    for (uint32_t i = vlo + tid * 16; i < vhi; i += ENC_THREADS * 16)
        *reinterpret_cast<uint4*>(gbase + i) = *reinterpret_cast<const uint4*>(smem + i);
    for (uint32_t i = max(vhi, min(vlo, stage_bytes)) + tid; i < stage_bytes; i += ENC_THREADS) gbase[i] = smem[i];
}

}  // namespace b200
