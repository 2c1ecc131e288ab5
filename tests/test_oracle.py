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
