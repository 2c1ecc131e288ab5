/* TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or called from the product path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use anything in oracle/.
 *
 * CPU restatement (plain C) of the reference INFLATE path, HyperBitGore/deflate.hpp:
 *   inflate::realDecompress            include/inflate.hpp:277-322
 *   inflate::decompressHuffmanBlock    include/inflate.hpp:226-275
 *   inflate::decodeTree & friends      include/inflate.hpp:136-224
 *   Bitwrapper::readBits / readByte    include/inflate.hpp:78-122
 *   FlatHuffmanTree::construct         include/common.hpp:104-145  (canonical code assignment)
 *   generateFixedCodes / ...Distance   include/common.hpp:442-495
 *   generateLengthLookup / Distance    include/common.hpp:508-575
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file against (a) the reference's own
 * fixtures zlib.dat / weird.dat (tests/golden/, SHA-1 of the decoded bytes), (b) the unmodified
 * reference compiled into oracle/_ref/libref_deflate.so when present, (c) zlib on generated streams.
 *
 * The reference decodes one bit at a time, extends `code` MSB-first and accepts on an exact
 * (code, length) hit in a tree keyed by the canonical code (inflate.hpp:232-235, common.hpp:201-220).
 * For a prefix-free canonical code that is the textbook count/first-code decode, which is what
 * decode_sym() below does -- same accepted symbols, same bit consumption.
 *
 * Reference behaviours kept on purpose (each one is covered by a test):
 *   - BTYPE 3 has no case label: the header is consumed and the loop goes on   (inflate.hpp:292-313)
 *   - NLEN of a stored block is read and never verified                        (inflate.hpp:296-297)
 *   - a back-reference whose distance exceeds the bytes produced so far copies nothing, silently
 *     (size_t underflow makes `i < buffer.size()` false)                       (inflate.hpp:268-270)
 *   - literal/length and distance code-length lists are parsed by two independent calls, `last_code`
 *     reset in between, so a repeat code cannot cross the boundary            (inflate.hpp:216-220)
 *   - running past the input ends the call with the "Reading bits beyond the alloted buffer size!"
 *     error (ORACLE_E_OVERRUN here).  The reference's bound test is `offset > size`, i.e. it may read
 *     the byte AT index `size` (one past the end, inflate.hpp:81,97,106); this restatement reads that
 *     byte as 0 (what a NUL-terminated Python bytes object gives the compiled reference).
 * Undefined behaviour in the reference (uninitialised `dss`, value_lookup_table overrun on repeat
 * runs past the list end, oversubscribed trees) is reported as ORACLE_E_DATA instead.
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_OK 0
#define ORACLE_E_OVERRUN (-1) /* reference throws std::runtime_error */
#define ORACLE_E_DATA (-2)    /* reference behaviour undefined / never terminates normally */
#define ORACLE_E_NOMEM (-3)

typedef struct {
    const uint8_t* data;
    size_t size;
    size_t offset;      /* byte index   (Bitwrapper::offset)     */
    unsigned bit_offset;/* bit in byte  (Bitwrapper::bit_offset) */
    int overrun;
} bitreader;

/* Bitwrapper::readBits, inflate.hpp:78-102 (LSB-first, value assembled low bits first). */
static uint32_t read_bits(bitreader* br, unsigned bits) {
    uint32_t val = 0;
    unsigned total = 0;
    if (br->offset > br->size) { br->overrun = 1; return 0; }
    while (bits > 0) {
        unsigned remaining = 8 - br->bit_offset;
        unsigned take = bits < remaining ? bits : remaining;
        uint32_t byte = br->offset < br->size ? br->data[br->offset] : 0;
        uint32_t chunk = (byte >> br->bit_offset) & ((1u << take) - 1u);
        val |= chunk << total;
        br->bit_offset += take;
        total += take;
        bits -= take;
        if (br->bit_offset > 7) { br->offset++; br->bit_offset = 0; }
        if (br->offset > br->size) { br->overrun = 1; return val; }
    }
    return val;
}

/* Bitwrapper::readByte, inflate.hpp:104-110 */
static uint8_t read_byte(bitreader* br) {
    br->bit_offset = 0;
    if (br->offset > br->size) { br->overrun = 1; return 0; }
    uint8_t b = br->offset < br->size ? br->data[br->offset] : 0;
    br->offset++;
    return b;
}

/* Bitwrapper::moveByte(true), inflate.hpp:112-119 */
static void align_byte(bitreader* br) {
    if (br->bit_offset != 0) { br->offset++; br->bit_offset = 0; }
}

/* Canonical code table: FlatHuffmanTree::construct, common.hpp:104-145.
 * count[l] symbols of length l; symbols sorted by (len, value); first code per length from the
 * bl_count / next_code recurrence (common.hpp:117-138). */
typedef struct {
    uint16_t count[17];
    uint16_t first[17];   /* first canonical code of each length */
    uint16_t index[17];   /* offset of that length's first symbol in sorted[] */
    uint16_t sorted[320];
    int nsyms;            /* symbols with len > 0 */
} hufftab;

static int build_table(hufftab* t, const uint8_t* lens, int n) {
    memset(t, 0, sizeof(*t));
    for (int i = 0; i < n; i++) {
        if (lens[i] > 15) return ORACLE_E_DATA;
        if (lens[i]) { t->count[lens[i]]++; t->nsyms++; }
    }
    uint32_t code = 0, idx = 0;
    long left = 1;
    for (int l = 1; l <= 15; l++) {
        code = (code + t->count[l - 1]) << 1;   /* common.hpp:127-130 (count[0] stays 0) */
        t->first[l] = (uint16_t)code;
        t->index[l] = (uint16_t)idx;
        idx += t->count[l];
        left = (left << 1) - t->count[l];
        if (left < 0) return ORACLE_E_DATA;     /* oversubscribed: reference tree is corrupt (UB) */
    }
    uint16_t next[17];
    memcpy(next, t->index, sizeof(next));
    for (int i = 0; i < n; i++)
        if (lens[i]) t->sorted[next[lens[i]]++] = (uint16_t)i;
    return ORACLE_OK;
}

/* One symbol: the bit-at-a-time loop of inflate.hpp:231-235 / 252-259.  `maxbits` is 16 for the
 * distance tree (inflate.hpp:252) and unbounded for literal/length (runs until the reader overruns). */
static int decode_sym(bitreader* br, const hufftab* t, int maxbits) {
    uint32_t code = 0;
    for (int len = 1; len <= maxbits; len++) {
        code = (code << 1) | read_bits(br, 1);
        if (br->overrun) return ORACLE_E_OVERRUN;
        if (len <= 15) {
            uint32_t rel = code - t->first[len];
            if (code >= t->first[len] && rel < t->count[len]) return t->sorted[t->index[len] + rel];
        }
    }
    return ORACLE_E_DATA; /* distance: `dss` would be used uninitialised (inflate.hpp:251-261) */
}

/* generateLengthLookup / generateDistanceLookup, common.hpp:508-575 (RFC 1951 3.2.5) */
static const uint16_t LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31,
                                      35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2,
                                      3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193,
                                       257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145,
                                       8193, 12289, 16385, 24577};
static const uint8_t DIST_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6,
                                       7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

typedef struct {
    uint8_t* buf;
    size_t len, cap;
} outbuf;

static int out_push(outbuf* o, uint8_t b) {
    if (o->len == o->cap) {
        size_t ncap = o->cap ? o->cap * 2 : 1 << 16;
        uint8_t* nb = (uint8_t*)realloc(o->buf, ncap);
        if (!nb) return ORACLE_E_NOMEM;
        o->buf = nb;
        o->cap = ncap;
    }
    o->buf[o->len++] = b;
    return ORACLE_OK;
}

/* readDynamicTreeCodes, inflate.hpp:166-206: `iterations` code lengths, repeat codes 16/17/18,
 * `last_code` starts at 0 for every list. */
static int read_code_lengths(bitreader* br, const hufftab* pre, int iterations, uint8_t* lens, int cap) {
    int i = 0;
    uint8_t last = 0;
    while (i < iterations) {
        int sym = decode_sym(br, pre, 1 << 30);
        if (sym < 0) return sym;
        int rep;
        switch (sym) {
            case 16:
                rep = 3 + (int)read_bits(br, 2);
                if (br->overrun) return ORACLE_E_OVERRUN;
                if (i + rep > cap) return ORACLE_E_DATA;
                for (int j = 0; j < rep; j++) lens[i++] = last;
                break;
            case 17:
                rep = 3 + (int)read_bits(br, 3);
                if (br->overrun) return ORACLE_E_OVERRUN;
                if (i + rep > cap) return ORACLE_E_DATA;
                for (int j = 0; j < rep; j++) lens[i++] = 0;
                break;
            case 18:
                rep = 11 + (int)read_bits(br, 7);
                if (br->overrun) return ORACLE_E_OVERRUN;
                if (i + rep > cap) return ORACLE_E_DATA;
                for (int j = 0; j < rep; j++) lens[i++] = 0;
                break;
            default:
                lens[i++] = (uint8_t)sym;
                last = (uint8_t)sym;
        }
    }
    /* the reference keeps entries past `iterations` (values >= the list size) which would index
     * value_lookup_table[300] out of range when large; treat any overshoot as invalid */
    if (i != iterations) return ORACLE_E_DATA;
    return ORACLE_OK;
}

/* decodeTree, inflate.hpp:208-224 + readCodeLengthTree :136-164 */
static int read_dynamic_tables(bitreader* br, hufftab* lit, hufftab* dist) {
    static const uint8_t ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    unsigned hlit = read_bits(br, 5), hdist = read_bits(br, 5), hclen = read_bits(br, 4);
    if (br->overrun) return ORACLE_E_OVERRUN;
    uint8_t prelens[19] = {0};
    for (unsigned i = 0; i < hclen + 4; i++) {
        prelens[ORDER[i]] = (uint8_t)read_bits(br, 3);
        if (br->overrun) return ORACLE_E_OVERRUN;
    }
    hufftab pre;
    int rc = build_table(&pre, prelens, 19);
    if (rc) return rc;
    uint8_t lens[320];
    memset(lens, 0, sizeof(lens));
    rc = read_code_lengths(br, &pre, 257 + (int)hlit, lens, 288);
    if (rc) return rc;
    rc = build_table(lit, lens, 257 + (int)hlit);
    if (rc) return rc;
    memset(lens, 0, sizeof(lens));
    rc = read_code_lengths(br, &pre, 1 + (int)hdist, lens, 32);
    if (rc) return rc;
    return build_table(dist, lens, 1 + (int)hdist);
}

/* decompressHuffmanBlock, inflate.hpp:226-275 */
static int inflate_huffman_block(bitreader* br, outbuf* o, const hufftab* lit, const hufftab* dist) {
    for (;;) {
        int sym = decode_sym(br, lit, 1 << 30);
        if (sym < 0) return sym;
        if (sym < 256) {
            int rc = out_push(o, (uint8_t)sym);
            if (rc) return rc;
        } else if (sym == 256) {
            return ORACLE_OK;
        } else {
            if (sym > 285) return ORACLE_E_DATA; /* rl.findCode miss: length 0, extra_bits -1 */
            uint32_t length = LEN_BASE[sym - 257];
            if (LEN_EXTRA[sym - 257]) {
                length += read_bits(br, LEN_EXTRA[sym - 257]);
                if (br->overrun) return ORACLE_E_OVERRUN;
            }
            int ds = decode_sym(br, dist, 16);
            if (ds < 0) return ds;
            if (ds > 29) return ORACLE_E_DATA;
            uint32_t distance = DIST_BASE[ds];
            if (DIST_EXTRA[ds]) {
                distance += read_bits(br, DIST_EXTRA[ds]);
                if (br->overrun) return ORACLE_E_OVERRUN;
            }
            /* inflate.hpp:268-270: i = size - distance underflows when distance > size, and the
             * loop condition `i < buffer.size()` is then false from the start: nothing is copied */
            if (distance <= o->len) {
                size_t from = o->len - distance;
                for (uint32_t j = 0; j < length; j++) {
                    int rc = out_push(o, o->buf[from + j]);
                    if (rc) return rc;
                }
            }
        }
    }
}

static void fixed_tables(hufftab* lit, hufftab* dist) {
    uint8_t lens[288];
    int i = 0;
    for (; i < 144; i++) lens[i] = 8;   /* common.hpp:446-448 */
    for (; i < 256; i++) lens[i] = 9;   /* common.hpp:449-451 */
    for (; i < 280; i++) lens[i] = 7;   /* common.hpp:453-469 */
    for (; i < 288; i++) lens[i] = 8;   /* common.hpp:470-480 */
    build_table(lit, lens, 288);
    uint8_t dl[32];
    for (i = 0; i < 32; i++) dl[i] = 5; /* common.hpp:484-495 */
    build_table(dist, dl, 32);
}

/* realDecompress, inflate.hpp:277-322.  On success *out is a malloc'ed buffer the caller frees with
 * oracle_free(); on error *out holds whatever was decoded before the error (may be NULL). */
int oracle_inflate(const void* in, size_t n, uint8_t** out, size_t* out_n) {
    bitreader br = {(const uint8_t*)in, n, 0, 0, 0};
    outbuf o = {NULL, 0, 0};
    hufftab fixed_lit, fixed_dist, lit, dist;
    fixed_tables(&fixed_lit, &fixed_dist);
    int rc = ORACLE_OK;
    for (;;) {
        unsigned final = read_bits(&br, 1);
        unsigned type = read_bits(&br, 2);
        if (br.overrun) { rc = ORACLE_E_OVERRUN; break; }
        if (type == 0) {
            align_byte(&br);
            unsigned len = read_bits(&br, 16);
            (void)read_bits(&br, 16); /* NLEN: read, never checked (inflate.hpp:297) */
            if (br.overrun) { rc = ORACLE_E_OVERRUN; break; }
            for (unsigned i = 0; i < len && !rc; i++) {
                uint8_t b = read_byte(&br);
                if (br.overrun) { rc = ORACLE_E_OVERRUN; break; }
                rc = out_push(&o, b);
            }
            if (rc) break;
        } else if (type == 1) {
            rc = inflate_huffman_block(&br, &o, &fixed_lit, &fixed_dist);
            if (rc) break;
        } else if (type == 2) {
            rc = read_dynamic_tables(&br, &lit, &dist);
            if (rc) break;
            rc = inflate_huffman_block(&br, &o, &lit, &dist);
            if (rc) break;
        } /* type 3: no case in the reference's switch -- ignored */
        if (final) break;
    }
    *out = o.buf;
    *out_n = o.len;
    return rc;
}

/* inflate::decompressZlib, inflate.hpp:326-335 / 352-361: skip the 2-byte zlib header.  The FDICT
 * test there is `extract1BitLeft(*(uint8_t*)in + 1, 5)`, i.e. bit 26 of (first byte + 1), which is
 * always 0 for a byte value -- so the skip is always 2 and the Adler-32 trailer is never looked at. */
int oracle_inflate_zlib(const void* in, size_t n, uint8_t** out, size_t* out_n) {
    if (n < 2) { *out = NULL; *out_n = 0; return ORACLE_E_OVERRUN; }
    return oracle_inflate((const uint8_t*)in + 2, n - 2, out, out_n);
}

void oracle_free(void* p) { free(p); }
