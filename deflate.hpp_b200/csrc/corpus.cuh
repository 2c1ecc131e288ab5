// corpus.cuh -- device generator for the synthetic mixed-entropy corpus of BASELINE configs 3 and 5.
// Bit-identical to the host restatement in oracle/corpus_oracle.c (tests/test_corpus.py checks it).
// Chunk c (64 KiB) has kind c % 3: 0 text-like, 1 image-like, 2 random.  See DESIGN.md "Corpus".
#pragma once
#include "common.cuh"

namespace b200 {

constexpr uint64_t GOLDEN = 0x9E3779B97F4A7C15ull;

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__device__ __forceinline__ uint32_t vocab_word(uint64_t seed, uint32_t k, uint8_t* w) {
    const uint64_t base = (seed ^ 0x5DEECE66Dull) + (uint64_t)k * 16u * GOLDEN;
    const uint32_t len = 2u + (uint32_t)(mix64(base + GOLDEN) % 9u);
    for (uint32_t i = 0; i < len; i++) w[i] = (uint8_t)('a' + mix64(base + (uint64_t)(i + 2) * GOLDEN) % 26u);
    return len;
}

// Text chunks are inherently sequential (variable-length words): one thread per text chunk.
__global__ void corpus_text_kernel(uint8_t* __restrict__ out, uint64_t seed, uint64_t first_chunk, uint64_t n_chunks) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_chunks) return;
    const uint64_t c = first_chunk + i;
    if (c % 3u != 0) return;
    uint8_t* dst = out + i * CHUNK;
    const uint64_t s = seed ^ (c * GOLDEN);
    uint64_t k = 0;
    uint32_t pos = 0, col = 0;
    uint8_t w[10];
    while (pos < CHUNK) {
        const uint64_t r = mix64(s + (++k) * GOLDEN);
        const uint32_t e = (uint32_t)(r % 12u);
        const uint32_t idx = ((1u << e) | ((uint32_t)(r >> 8) & ((1u << e) - 1u))) - 1u;
        const uint32_t len = vocab_word(seed, idx, w);
        for (uint32_t j = 0; j < len && pos < CHUNK; j++) dst[pos++] = w[j];
        col += len;
        if (pos < CHUNK) {
            if (col >= 72) { dst[pos++] = '\n'; col = 0; }
            else { dst[pos++] = ' '; col++; }
        }
    }
}

// Image-like and random chunks are stateless per byte / per 8-byte word: one thread per 8 bytes.
__global__ void corpus_flat_kernel(uint8_t* __restrict__ out, uint64_t seed, uint64_t first_chunk, uint64_t n_chunks) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;   // 8-byte word index
    const uint64_t i = t / (CHUNK / 8);
    if (i >= n_chunks) return;
    const uint64_t c = first_chunk + i;
    const uint32_t kind = (uint32_t)(c % 3u);
    if (kind == 0) return;
    const uint32_t wd = (uint32_t)(t % (CHUNK / 8));
    uint64_t v = 0;
    if (kind == 2) {
        v = mix64((seed ^ (c * GOLDEN)) + (uint64_t)(wd + 1) * GOLDEN);
    } else {
        const uint64_t g0 = c * (uint64_t)CHUNK + (uint64_t)wd * 8;
        for (uint32_t b = 0; b < 8; b++) {
            const uint64_t g = g0 + b;
            const uint64_t pix = g / 3u;
            const uint32_t comp = (uint32_t)(g % 3u);
            const uint32_t x = (uint32_t)(pix & 1023u);
            const uint64_t y = pix >> 10;
            const uint32_t val = (uint32_t)(((x >> 2) + (y >> 2)) & 255u);
            uint32_t byte;
            if (comp == 0) byte = val + (uint32_t)(mix64((seed ^ 0x1234567ull) + (pix + 1) * GOLDEN) % 3u);
            else if (comp == 1) byte = 2u * val;
            else byte = 255u - val;
            v |= (uint64_t)(byte & 0xFFu) << (8 * b);
        }
    }
    *reinterpret_cast<uint64_t*>(out + i * CHUNK + (uint64_t)wd * 8) = v;
}

}  // namespace b200
