/* TEST INFRASTRUCTURE ONLY -- host-side restatement of the synthetic corpus generator.
 *
 * The reference ships no large input (large.bmp is missing from its checkout), so
 * BASELINE.json's "1 GB synthetic mixed-entropy corpus (text-like + image-like + random) in 64 KB
 * chunks" is defined here, integer-only so that the device generator
 * (deflate.hpp_b200/csrc/corpus.cuh) produces the same bytes bit for bit.  tests/test_corpus.py
 * freezes SHA-256 digests of the first chunk of each kind and compares host vs device output.
 *
 * Chunk c (global index, 65 536 bytes) has kind c % 3:
 *   0  T  text-like : words from a 4096-word vocabulary, octave-uniform ("Zipf-like") word index,
 *                     single spaces, '\n' once a line reaches 72 columns
 *   1  I  image-like: 24-bpp scanlines 1024 px wide addressed by the global byte offset;
 *                     pixel (x,y): v = ((x>>2)+(y>>2)) & 255 -> bytes (v + noise%3, 2v & 255, 255-v)
 *   2  R  random    : raw splitmix64 output
 * RNG: splitmix64; per-chunk stream k-th value = mix(seed ^ (c * GOLDEN) + (k+1) * GOLDEN).
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#define CORPUS_CHUNK 65536u
#define GOLDEN 0x9E3779B97F4A7C15ull
#define VOCAB_WORDS 4096
#define VOCAB_MAXLEN 10

static uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* Vocabulary: word k has length 2 + r % 9 (2..10), letters 'a' + r % 26.  Stateless in k. */
static unsigned vocab_word(uint64_t seed, unsigned k, uint8_t* w) {
    uint64_t base = (seed ^ 0x5DEECE66Dull) + (uint64_t)k * 16u * GOLDEN;
    unsigned len = 2u + (unsigned)(mix64(base + GOLDEN) % 9u);
    for (unsigned i = 0; i < len; i++) w[i] = (uint8_t)('a' + mix64(base + (uint64_t)(i + 2) * GOLDEN) % 26u);
    return len;
}

static void gen_text(uint64_t seed, uint64_t c, uint8_t* out) {
    uint64_t s = seed ^ (c * GOLDEN);
    uint64_t k = 0;
    unsigned pos = 0, col = 0;
    uint8_t w[VOCAB_MAXLEN];
    while (pos < CORPUS_CHUNK) {
        uint64_t r = mix64(s + (++k) * GOLDEN);
        unsigned e = (unsigned)(r % 12u);
        unsigned idx = ((1u << e) | ((unsigned)(r >> 8) & ((1u << e) - 1u))) - 1u; /* 0..4094 */
        unsigned len = vocab_word(seed, idx, w);
        for (unsigned i = 0; i < len && pos < CORPUS_CHUNK; i++) out[pos++] = w[i];
        col += len;
        if (pos < CORPUS_CHUNK) {
            if (col >= 72) { out[pos++] = '\n'; col = 0; }
            else { out[pos++] = ' '; col++; }
        }
    }
}

static void gen_image(uint64_t seed, uint64_t c, uint8_t* out) {
    uint64_t g0 = c * (uint64_t)CORPUS_CHUNK;
    for (unsigned j = 0; j < CORPUS_CHUNK; j++) {
        uint64_t g = g0 + j;
        uint64_t pix = g / 3u;
        unsigned comp = (unsigned)(g % 3u);
        unsigned x = (unsigned)(pix & 1023u);
        uint64_t y = pix >> 10;
        unsigned v = (unsigned)(((x >> 2) + (y >> 2)) & 255u);
        unsigned b;
        if (comp == 0) b = v + (unsigned)(mix64((seed ^ 0x1234567ull) + (pix + 1) * GOLDEN) % 3u);
        else if (comp == 1) b = 2u * v;
        else b = 255u - v;
        out[j] = (uint8_t)b;
    }
}

static void gen_random(uint64_t seed, uint64_t c, uint8_t* out) {
    uint64_t s = seed ^ (c * GOLDEN);
    for (unsigned wd = 0; wd < CORPUS_CHUNK / 8u; wd++) {
        uint64_t r = mix64(s + (uint64_t)(wd + 1) * GOLDEN);
        for (unsigned b = 0; b < 8; b++) out[wd * 8u + b] = (uint8_t)(r >> (8u * b));
    }
}

/* Fill out[0 .. nchunks*65536) with chunks first_chunk .. first_chunk+nchunks-1. */
void oracle_corpus_generate(uint64_t seed, uint64_t first_chunk, uint64_t nchunks, uint8_t* out) {
    for (uint64_t i = 0; i < nchunks; i++) {
        uint64_t c = first_chunk + i;
        uint8_t* dst = out + i * (size_t)CORPUS_CHUNK;
        switch (c % 3u) {
            case 0: gen_text(seed, c, dst); break;
            case 1: gen_image(seed, c, dst); break;
            default: gen_random(seed, c, dst); break;
        }
    }
}
