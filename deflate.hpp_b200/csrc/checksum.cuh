// checksum.cuh -- Adler-32 (RFC 1950) on the device.
//
// The reference's inflate::decompressZlib (include/inflate.hpp:326-361) skips the two header bytes and never
// looks at the Adler-32 trailer; with B200_F_STRICT this library verifies it (and the header check bits) on the
// decoded bytes where they already are -- in HBM.  SURVEY.md 8(f) rank 2.
//
// adler(d[0..n)) = (B << 16) | A with A = 1 + sum d_i, B = sum of the running A, both mod 65521.  For a block of
// L bytes with s1 = sum d_i and s2 = sum i * d_i (i = offset inside the block):  A' = A + s1,
// B' = B + L * A + (L * s1 - s2).  One CTA per 64 KiB block computes (s1, s2) with coalesced 16-byte loads; the
// blocks are then folded in order by one warp (lane-strided partial folds, combined the same way).
#pragma once
#include "common.cuh"

namespace b200 {

constexpr uint32_t ADLER_MOD = 65521;
constexpr uint32_t ADLER_THREADS = 256;
struct AdlerPart { uint32_t a, b; };      // block contribution: a = s1 mod M, b = (L * s1 - s2) mod M

__global__ void __launch_bounds__(ADLER_THREADS)
adler_partial_kernel(const uint8_t* __restrict__ data, uint64_t n, AdlerPart* __restrict__ parts) {
    __shared__ unsigned long long s_s1[ADLER_THREADS / 32], s_s2[ADLER_THREADS / 32];
    const uint64_t base = (uint64_t)blockIdx.x * CHUNK;
    const uint32_t L = (uint32_t)min((uint64_t)CHUNK, n - base);
    const uint8_t* p = data + base;
    unsigned long long s1 = 0, s2 = 0;
    if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
        const uint32_t vecs = L >> 4;
        for (uint32_t v = threadIdx.x; v < vecs; v += ADLER_THREADS) {
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(p) + v);
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
            uint32_t t1 = 0, t2 = 0;                      // sum of the 16 bytes, sum of j * byte_j (j = 0..15)
            #pragma unroll
            for (uint32_t k = 0; k < 4; k++) {
                #pragma unroll
                for (uint32_t j = 0; j < 4; j++) {
                    const uint32_t d = (w[k] >> (8 * j)) & 0xFFu;
                    t1 += d; t2 += (4 * k + j) * d;
                }
            }
            s1 += t1;
            s2 += (unsigned long long)(v * 16) * t1 + t2;
        }
        for (uint32_t i = (vecs << 4) + threadIdx.x; i < L; i += ADLER_THREADS) { s1 += p[i]; s2 += (unsigned long long)i * p[i]; }
    } else {
        for (uint32_t i = threadIdx.x; i < L; i += ADLER_THREADS) { s1 += p[i]; s2 += (unsigned long long)i * p[i]; }
    }
    for (int o = 16; o; o >>= 1) {
        s1 += __shfl_xor_sync(0xFFFFFFFFu, s1, o);
        s2 += __shfl_xor_sync(0xFFFFFFFFu, s2, o);
    }
    if ((threadIdx.x & 31) == 0) { s_s1[threadIdx.x >> 5] = s1; s_s2[threadIdx.x >> 5] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t1 = 0, t2 = 0;
        for (uint32_t k = 0; k < ADLER_THREADS / 32; k++) { t1 += s_s1[k]; t2 += s_s2[k]; }
        AdlerPart r;
        r.a = (uint32_t)(t1 % ADLER_MOD);
        r.b = (uint32_t)(((unsigned long long)L * t1 - t2) % ADLER_MOD);     // L * s1 >= s2 because i < L
        parts[blockIdx.x] = r;
    }
}

// Folds the block contributions in order.  Every full block has L = CHUNK; only the last may be shorter.
// state (A, B) -> after a block: B += L * A + b, A += a.  Lanes fold contiguous runs of blocks starting from
// (A, B) = (0, 0); a run of blocks is itself "a block" of the summed length, so lane results combine the same way.
__global__ void adler_fold_kernel(const AdlerPart* __restrict__ parts, uint64_t nblocks, uint64_t n, uint32_t* __restrict__ out) {
    const uint32_t lane = threadIdx.x;
    const uint64_t per = (nblocks + 31) / 32;
    const uint64_t lo = min(nblocks, lane * per), hi = min(nblocks, lo + per);
    unsigned long long A = 0, B = 0, len = 0;              // contribution of blocks [lo, hi) to a zero start state
    for (uint64_t k = lo; k < hi; k++) {
        const unsigned long long L = min((uint64_t)CHUNK, n - k * CHUNK);
        B = (B + (L % ADLER_MOD) * A + parts[k].b) % ADLER_MOD;
        A = (A + parts[k].a) % ADLER_MOD;
        len += L;
    }
    // in-order combine across lanes by lane 0
    __shared__ unsigned long long sA[32], sB[32], sL[32];
    sA[lane] = A; sB[lane] = B; sL[lane] = len;
    __syncwarp();
    if (lane == 0) {
        unsigned long long a = 1, b = 0;                    // Adler-32 start state
        for (uint32_t k = 0; k < 32; k++) {
            b = (b + (sL[k] % ADLER_MOD) * a + sB[k]) % ADLER_MOD;
            a = (a + sA[k]) % ADLER_MOD;
        }
        *out = (uint32_t)((b << 16) | a);
    }
}

}  // namespace b200
