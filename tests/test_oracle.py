"""CPU tests (-m "not gpu"): pin the oracle (oracle/inflate_oracle.c, oracle/corpus_oracle.c) against
the reference's own fixtures, the unmodified reference (oracle/_ref) and zlib.  SURVEY.md 8(c)."""
import hashlib
import zlib

import pytest

from conftest import gold, zlib_raw_inflate
import datagen

# SHA-1 of the decoded fixtures; equal to zlib's and the reference's output
GOLDEN_SHA1 = {"zlib.dat": ("bcc0f00aa007", 72541), "weird.dat": ("ff59bf286819", 6050)}
# compressed sizes the unmodified reference produces: level -> size
REF_SIZES = {"test.bmp": {0: 21904, 1: 21904, 2: 5346, 3: 3124}, "tiny.bmp": {0: 264, 1: 264, 2: 99, 3: 67}}


@pytest.mark.parametrize("name", ["zlib.dat", "weird.dat"])
def test_oracle_fixtures(oracle, name):
    raw = gold(name)
    rc, out = oracle.inflate_zlib(raw)
    assert rc == 0
    sha, size = GOLDEN_SHA1[name]
    assert len(out) == size and hashlib.sha1(out).hexdigest().startswith(sha)
    assert out == zlib.decompress(raw)


@pytest.mark.parametrize("name", ["zlib.dat", "weird.dat"])
def test_reference_fixtures(ref, name):
    raw = gold(name)
    n, out = ref.inflate_zlib(raw)
    sha, size = GOLDEN_SHA1[name]
    assert n == size and hashlib.sha1(out).hexdigest().startswith(sha)


@pytest.mark.parametrize("name", ["test.bmp", "tiny.bmp"])
@pytest.mark.parametrize("level", [0, 1, 2, 3])
def test_oracle_equals_reference_on_reference_streams(oracle, ref, name, level):
    """The oracle inflater must reproduce the reference inflater on reference-compressed streams --
    including level 2, whose streams decode to WRONG bytes: same wrong bytes."""
    data = gold(name)
    c = ref.compress(data, level)
    assert len(c) == REF_SIZES[name][level]
    n, r_out = ref.inflate(c)
    rc, o_out = oracle.inflate(c)
    assert rc == 0 and n == len(o_out) and r_out == o_out
    z_out, _ = zlib_raw_inflate(c)
    assert z_out == o_out
    if level != 2:
        assert o_out == data
    else:
        assert o_out != data and len(o_out) == len(data)   # the reference's fast level corrupts


@pytest.mark.parametrize("kind", sorted(datagen.KINDS))
def test_oracle_vs_zlib_streams(oracle, kind):
    data = datagen.KINDS[kind](70000)
    for name, stream in datagen.foreign_streams(data).items():
        rc, out = oracle.inflate(stream)
        assert rc == 0 and out == data, name


def test_oracle_vs_reference_on_zlib_streams(oracle, ref):
    data = datagen.text_like(30000) + datagen.image_like(30000)
    for name, stream in datagen.foreign_streams(data).items():
        n, r_out = ref.inflate(stream)
        rc, o_out = oracle.inflate(stream)
        assert rc == 0 and r_out == o_out == data, name


def test_oracle_errors_match_reference(oracle, ref):
    data = datagen.text_like(5000)
    good = datagen.foreign_streams(data)["zlib6"]
    for cut in (len(good) // 2, len(good) - 3, 5, 1, 0):
        rc, _ = oracle.inflate(good[:cut])
        n, _ = ref.inflate(good[:cut])
        assert (rc == -1) == (n == -1), cut
        assert rc != 0


def test_oracle_reference_quirks(oracle, ref):
    # NLEN is never verified (inflate.hpp:296-297)
    bad_nlen = bytes([0x01, 0x03, 0x00, 0x00, 0x00]) + b"abc"
    rc, out = oracle.inflate(bad_nlen)
    assert rc == 0 and out == b"abc"
    assert ref.inflate(bad_nlen)[1] == b"abc"
    # distance beyond the produced output copies nothing (inflate.hpp:268-270):
    # fixed block: literal 'a' (0x61 -> code 0x91, 8 bits), match len 3 (sym 257) dist 4 (code 3), EOB
    stream = datagen.too_far_stream()
    rc, out = oracle.inflate(stream)
    assert rc == 0 and out == b"a"
    assert ref.inflate(stream)[1] == b"a"
    with pytest.raises(zlib.error):
        zlib.decompressobj(-15).decompress(stream)


def test_corpus_digests(oracle):
    """Freeze the synthetic corpus (BASELINE config 3): SHA-256 of the first chunk of each kind."""
    c = oracle.corpus(20261018, 0, 3)
    digests = [hashlib.sha256(c[i * 65536:(i + 1) * 65536]).hexdigest()[:16] for i in range(3)]
    assert digests == CORPUS_DIGESTS, digests
    # chunks are addressable independently
    assert oracle.corpus(20261018, 2, 1) == c[2 * 65536:]


CORPUS_DIGESTS = ['06ff5f27ec279c51', '72bb2045a8a9de18', 'b6398b943ef16d63']


# ---- segment index: the format, pinned on the CPU against a committed GPU-compressed stream ----------------
class _Bits:
    def __init__(self, data, pos=0):
        self.d, self.pos = data, pos          # pos in bits

    def get(self, n):
        v = 0
        for i in range(n):
            v |= ((self.d[self.pos >> 3] >> (self.pos & 7)) & 1) << i
            self.pos += 1
        return v


def _canon(lens):
    """code lengths -> {(length, code): symbol} (RFC 1951 3.2.2)."""
    bl = [0] * 16
    for l in lens:
        bl[l] += 1
    bl[0] = 0
    nxt, code = [0] * 16, 0
    for b in range(1, 16):
        code = (code + bl[b - 1]) << 1
        nxt[b] = code
    out = {}
    for s, l in enumerate(lens):
        if l:
            out[(l, nxt[l])] = s
            nxt[l] += 1
    return out


def _sym(br, table):
    code = 0
    for l in range(1, 16):
        code = (code << 1) | br.get(1)
        if (l, code) in table:
            return table[(l, code)]
    raise AssertionError("bad code")


_LEN_BASE = [3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258]
_LEN_EXTRA = [0] * 8 + [1] * 4 + [2] * 4 + [3] * 4 + [4] * 4 + [5] * 4 + [0]
_DIST_BASE = [1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145,
              8193, 12289, 16385, 24577]
_DIST_EXTRA = [0, 0, 0, 0] + [i // 2 for i in range(2, 28)]


def _dynamic_block_boundaries(data, bitpos, seg=4096):
    """Decodes ONE dynamic Huffman block starting at `bitpos` (its 3 header bits included); returns the bit position
    of the first symbol, of every symbol that starts at a multiple of `seg` output bytes, and the output length."""
    br = _Bits(data, bitpos)
    hdr = br.get(3)
    assert hdr >> 1 == 2, "fixture chunks are dynamic blocks"
    hlit, hdist, hclen = br.get(5) + 257, br.get(5) + 1, br.get(4) + 4
    order = [16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15]
    pl = [0] * 19
    for i in range(hclen):
        pl[order[i]] = br.get(3)
    pre = _canon(pl)
    lens = []
    while len(lens) < hlit + hdist:
        s = _sym(br, pre)
        if s < 16:
            lens.append(s)
        elif s == 16:
            lens += [lens[-1]] * (3 + br.get(2))
        elif s == 17:
            lens += [0] * (3 + br.get(3))
        else:
            lens += [0] * (11 + br.get(7))
    lit, dist = _canon(lens[:hlit]), _canon(lens[hlit:hlit + hdist])
    marks, out = {0: br.pos}, 0
    while True:
        if out % seg == 0:
            marks[out] = br.pos
        s = _sym(br, lit)
        if s < 256:
            out += 1
        elif s == 256:
            break
        else:
            out += _LEN_BASE[s - 257] + br.get(_LEN_EXTRA[s - 257])
            d = _sym(br, dist)
            br.get(_DIST_EXTRA[d])
    return marks, out, br.pos


def test_segment_index_fixture(oracle, ref):
    """tests/golden/b200_indexed.deflate was compressed on a B200 (tools/make_index_fixture.py).  The index in
    front of each full chunk -- 64 empty stored blocks whose padding bits spell 16 words (csrc/common.cuh) -- must
    state exactly the bit lengths an independent bit-by-bit decoder sees between 4 KiB output boundaries, every
    token must end on those boundaries, and the oracle (the reference's inflater restated) must skip the index."""
    import datagen
    c = gold("b200_indexed.deflate")
    data = datagen.text_like(65536, seed=51) + datagen.image_like(65536, seed=52) + datagen.text_like(65536 + 3000, seed=53)
    rc, out = oracle.inflate(c)
    assert rc == 0 and out == data
    n, r_out = ref.inflate(c)                    # the unmodified reference inflater skips the index as well
    assert n == len(data) and r_out == data
    assert zlib.decompressobj(-15).decompress(c) == data
    pos = 0                                     # byte offset of the current chunk
    for chunk in range(3):
        words = [0] * 16
        for g in range(64):
            b = c[pos + 5 * g: pos + 5 * g + 5]
            assert (b[0] & 0x87) == 0x80 and b[1:] == b"\x00\x00\xff\xff", (chunk, g)
            words[g >> 2] |= ((b[0] >> 3) & 15) << (4 * (g & 3))
        assert words[0] == 0x2B5 | (15 << 10)
        marks, produced, endbit = _dynamic_block_boundaries(c, (pos + 320) * 8)
        # no token crosses a segment boundary (the mark at 65536 is the end-of-block symbol)
        assert produced == 65536 and sorted(marks) == [4096 * s for s in range(17)]
        for s in range(1, 16):
            assert words[s] == marks[4096 * s] - marks[4096 * (s - 1)], (chunk, s)
        # the separator: two empty stored blocks with zero padding; the next chunk starts right after it
        sep = (endbit + 3 + 7) // 8            # 3 header bits of the first one, padded to a byte
        assert c[sep:sep + 4] == b"\x00\x00\xff\xff" and c[sep + 4:sep + 9] == b"\x00\x00\x00\xff\xff"
        pos = sep + 9
    assert (c[pos] & 0x87) != 0x80 or c[pos + 1:pos + 5] != b"\x00\x00\xff\xff"          # the partial chunk has no index


def test_segment_index_short_fixture(oracle, ref):
    """tests/golden/b200_indexed_short.deflate: ONE short chunk (40 000 bytes of text, ten segments), compressed on a B200
    (tools/make_index_fixture.py).  Its index has its own size -- four groups per segment, word 0 = magic | (10 - 1) << 10
    (csrc/common.cuh) -- and states the bit lengths an independent decoder sees between 4 KiB output boundaries; the block
    is final, nothing follows it; zlib, the oracle and the unmodified reference inflater skip the index."""
    import datagen
    c = gold("b200_indexed_short.deflate")
    data = datagen.text_like(40000, seed=61)
    nseg = (len(data) + 4095) // 4096
    assert nseg == 10
    rc, out = oracle.inflate(c)
    assert rc == 0 and out == data
    n, r_out = ref.inflate(c)
    assert n == len(data) and r_out == data
    o = zlib.decompressobj(-15)
    assert o.decompress(c) == data and o.eof and o.unused_data == b""
    words = [0] * nseg
    for g in range(4 * nseg):
        b = c[5 * g: 5 * g + 5]
        assert (b[0] & 0x87) == 0x80 and b[1:] == b"\x00\x00\xff\xff", g
        words[g >> 2] |= ((b[0] >> 3) & 15) << (4 * (g & 3))
    assert words[0] == 0x2B5 | ((nseg - 1) << 10)
    # the 41st group does not exist: the Huffman block starts right behind the index, and it is the final one
    b = c[20 * nseg: 20 * nseg + 5]
    assert not ((b[0] & 0x87) == 0x80 and b[1:] == b"\x00\x00\xff\xff")
    assert c[20 * nseg] & 1 == 1
    marks, produced, endbit = _dynamic_block_boundaries(c, 20 * nseg * 8)
    assert produced == len(data) and sorted(marks) == [4096 * s for s in range(nseg)]
    for s in range(1, nseg):
        assert words[s] == marks[4096 * s] - marks[4096 * (s - 1)], s
    assert (endbit + 7) // 8 == len(c)
