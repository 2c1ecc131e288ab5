"""GPU parity tests for the inflater (-m gpu): bit-exact against the oracle (C restatement of the
reference inflater), the unmodified reference when its library travelled, and zlib."""
import hashlib
import zlib

import numpy as np
import pytest

from conftest import gold, zlib_raw_inflate
import datagen

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,sha,size", [("zlib.dat", "bcc0f00aa007", 72541), ("weird.dat", "ff59bf286819", 6050)])
def test_reference_fixtures(b200, oracle, name, sha, size):
    """zlib.dat: 3 dynamic blocks with cross-block back-references; weird.dat: HLIT 286, length 258."""
    raw = gold(name)
    out = b200.decompress_zlib(raw)
    assert len(out) == size and hashlib.sha1(out).hexdigest().startswith(sha)
    rc, o = oracle.inflate_zlib(raw)
    assert rc == 0 and out == o
    assert b200.decompress(raw[2:]) == out
    # caller-buffer overload truncates silently (inflate.hpp:345)
    assert b200.decompress_zlib(raw, out_size=size // 3) == out[:size // 3]
    assert b200.decompress_zlib(raw, out_size=size + 100) == out


@pytest.mark.parametrize("level", [0, 1, 2, 3])
def test_reference_compressed_streams(b200, oracle, ref, level):
    """GPU inflater == reference inflater on reference-compressed streams, including the level-2
    streams that decode to wrong bytes (same wrong bytes expected)."""
    for name in ("test.bmp", "tiny.bmp"):
        c = ref.compress(gold(name), level)
        n, r_out = ref.inflate(c)
        assert n >= 0
        assert b200.decompress(c) == r_out
    data = datagen.text_like(40000)
    c = ref.compress(data, level if level != 3 else 1)
    n, r_out = ref.inflate(c)
    if n >= 0:
        assert b200.decompress(c) == r_out
    else:
        with pytest.raises(b200.B200Error):
            b200.decompress(c)


@pytest.mark.parametrize("kind", sorted(datagen.KINDS))
def test_foreign_streams(b200, oracle, kind):
    data = datagen.KINDS[kind](150000)
    for name, stream in datagen.foreign_streams(data).items():
        out = b200.decompress(stream)
        assert out == data, name
        rc, o = oracle.inflate(stream)
        assert rc == 0 and o == out, name


def test_edge_sizes(b200):
    src = datagen.text_like(210000, seed=9)
    for n in datagen.EDGE_SIZES:
        for lvl in (0, 6):
            c = datagen.foreign_streams(src[:n])["zlib6" if lvl else "stored"]
            assert b200.decompress(c) == src[:n], (n, lvl)


def test_errors(b200, oracle):
    data = datagen.text_like(60000)
    good = datagen.foreign_streams(data)["zlib6"]
    for cut in (len(good) // 2, len(good) - 4, 7, 1):
        rc, _ = oracle.inflate(good[:cut])
        assert rc != 0
        with pytest.raises(b200.B200Error) as e:
            b200.decompress(good[:cut])
        assert e.value.code in (1, 2)
        if rc == -1:
            assert e.value.code == 1 and "beyond the alloted buffer size" in str(e.value)
    # garbage
    junk = datagen.random_bytes(5000, seed=11)
    rc, _ = oracle.inflate(junk)
    if rc != 0:
        with pytest.raises(b200.B200Error):
            b200.decompress(junk)


def test_reference_quirks_and_strict(b200, oracle):
    bad_nlen = bytes([0x01, 0x03, 0x00, 0x00, 0x00]) + b"abc"
    assert b200.decompress(bad_nlen) == oracle.inflate(bad_nlen)[1] == b"abc"
    with pytest.raises(b200.B200Error):
        b200.decompress(bad_nlen, flags=b200.F_STRICT)
    # distance beyond the produced output: the reference copies nothing; strict mode rejects
    stream = datagen.too_far_stream()
    rc, o = oracle.inflate(stream)
    assert rc == 0 and o == b"a"
    assert b200.decompress(stream) == b"a"
    with pytest.raises(b200.B200Error):
        b200.decompress(stream, flags=b200.F_STRICT)
    # BTYPE 3 is skipped by the reference (no case label): header 0b110 (non-final, type 3) then a
    # final stored block.  After the 3 header bits the next block header follows immediately.
    bt3 = bytes([0b00001110, 0x02, 0x00, 0xFD, 0xFF]) + b"hi"
    rc, o = oracle.inflate(bt3)
    assert rc == 0 and o == b"hi"
    assert b200.decompress(bt3) == b"hi"
    with pytest.raises(b200.B200Error):
        b200.decompress(bt3, flags=b200.F_STRICT)


@pytest.mark.parametrize("two_pass", [False, True])
def test_batch_mixed(b200, oracle, monkeypatch, two_pass):
    """BASELINE config 4 at reduced count: many independent small streams, one warp each (default) or one
    thread each through the two-pass path (B200_BATCH_TP=1)."""
    import torch
    if two_pass:
        monkeypatch.setenv("B200_BATCH_TP", "1")
    rng = np.random.default_rng(5)
    kinds = sorted(datagen.KINDS)
    streams, expect = [], []
    fixtures = [gold("zlib.dat")[2:], gold("weird.dat")[2:]]
    for i in range(600):
        n = int(np.exp(rng.uniform(np.log(1024), np.log(65536))))
        data = datagen.KINDS[kinds[i % len(kinds)]](n, seed=100 + i)
        prods = datagen.foreign_streams(data)
        name = sorted(prods)[i % len(prods)]
        streams.append(prods[name])
        expect.append(data)
    for f in fixtures:
        streams.append(f)
        expect.append(oracle.inflate(f)[1])
    streams.append(streams[0][: len(streams[0]) // 2])   # one broken stream must not disturb the others
    expect.append(None)
    in_off = np.cumsum([0] + [len(s) + 3 for s in streams[:-1]]).astype(np.uint64)   # odd alignment on purpose
    in_len = np.array([len(s) for s in streams], dtype=np.uint64)
    blob = bytearray(int(in_off[-1] + in_len[-1]) + 64)
    for o, s in zip(in_off, streams):
        blob[int(o):int(o) + len(s)] = s
    caps = np.array([len(e) if e is not None else 70000 for e in expect], dtype=np.uint64)
    out_off = np.cumsum([0] + list(caps[:-1])).astype(np.uint64)
    d_in = torch.frombuffer(blob, dtype=torch.uint8).cuda()
    d_out = torch.zeros(int(out_off[-1] + caps[-1]), dtype=torch.uint8, device="cuda")
    t = lambda a: torch.from_numpy(a.view(np.int64)).cuda()
    d_in_off, d_in_len, d_out_off, d_caps = t(in_off), t(in_len), t(out_off), t(caps)
    d_out_len = torch.zeros(len(streams), dtype=torch.int64, device="cuda")
    d_status = torch.full((len(streams),), -1, dtype=torch.int32, device="cuda")
    ctx = b200.Context(0)
    ctx.inflate_batch_dev(d_in.data_ptr(), d_in_off.data_ptr(), d_in_len.data_ptr(), d_out.data_ptr(),
                          d_out_off.data_ptr(), d_caps.data_ptr(), d_out_len.data_ptr(), d_status.data_ptr(),
                          len(streams))
    torch.cuda.synchronize()
    status = d_status.cpu().numpy()
    lens = d_out_len.cpu().numpy()
    host = d_out.cpu().numpy().tobytes()
    for i, e in enumerate(expect):
        if e is None:
            assert status[i] != 0
            continue
        assert status[i] == 0 and lens[i] == len(e), i
        assert host[int(out_off[i]):int(out_off[i]) + len(e)] == e, i


def test_own_streams_chunk_parallel_vs_sequential(b200, monkeypatch):
    data = datagen.text_like(400000) + datagen.random_bytes(100000) + datagen.image_like(300000)
    c = b200.compress(data, 2)
    assert b200.decompress(c) == data
    monkeypatch.setenv("B200_INFLATE_SEQUENTIAL", "1")
    assert b200.decompress(c) == data


def test_small_streams_shared_memory_window(b200, oracle, monkeypatch):
    """Short streams into small buffers are decoded by one warp with the output window in shared memory
    (inflate_small_kernel); B200_NO_SMALL_INFLATE=1 sends them down the regular paths.  Same bytes, same sizes, same
    errors -- own streams of every level, ten foreign producers, the reference's fixtures, truncating buffers, damaged
    input -- and the oracle agrees."""
    from conftest import gold
    bmp = gold("test.bmp")
    cases = []
    for data in (bmp, gold("tiny.bmp"), datagen.text_like(150000), datagen.image_like(70000), datagen.random_bytes(5000),
                 b"", b"a", b"\x00" * 160000, datagen.text_like(163840)):
        for level in (0, 1, 2, 3):
            c = b200.compress(data, level)
            if len(c) <= 65536:
                cases.append((c, data))
    for name, z in datagen.foreign_streams(bmp + datagen.text_like(60000)).items():
        if len(z) <= 65536:
            cases.append((z, bmp + datagen.text_like(60000)))
    assert len(cases) > 30
    raw = [gold("zlib.dat")[2:], gold("weird.dat")[2:], datagen.too_far_stream()]

    def run():
        out = []
        for c, data in cases:
            out.append(b200.decompress(c))                                    # size unknown: probe, then exact
            out.append(b200.decompress(c, out_size=len(data)))                # exact buffer
            out.append(b200.decompress(c, out_size=len(data) // 3))           # truncating buffer (inflate.hpp:345)
            out.append(b200.decompress(c, out_size=len(data) + 1000))
        for z in raw:
            out.append(b200.decompress(z))
            out.append(b200.decompress(z, out_size=100))
        for c, data in cases[::5]:                                             # damaged input: same outcome either way
            for cut in (len(c) // 2, max(len(c) - 1, 0)):
                try:
                    out.append(b200.decompress(c[:cut], out_size=len(data)))
                except b200.B200Error as e:
                    out.append(("error", e.code))
            bad = bytearray(c)
            if bad:
                bad[len(bad) // 2] ^= 0x5A
            try:
                out.append(b200.decompress(bytes(bad), out_size=len(data)))
            except b200.B200Error as e:
                out.append(("error", e.code))
        return out

    small = run()
    monkeypatch.setenv("B200_NO_SMALL_INFLATE", "1")
    regular = run()
    monkeypatch.delenv("B200_NO_SMALL_INFLATE")
    assert len(small) == len(regular)
    for k, (a, b) in enumerate(zip(small, regular)):
        assert a == b, k
    k = 0
    for c, data in cases:
        assert small[k] == data and small[k + 1] == data and small[k + 2] == data[:len(data) // 3] and small[k + 3] == data
        k += 4
    for z in raw:
        rc, o = oracle.inflate(z)
        assert rc == 0 and small[k] == o and small[k + 1] == o[:100]
        k += 2


def test_false_sync_markers(b200):
    """Stored data full of 00 00 FF FF patterns must not confuse the chunk-parallel path."""
    data = (b"\x00\x00\xff\xff" * 40000) + datagen.random_bytes(70000) + b"\x00\x00\xff\xff" * 10
    for level in (0, 2):
        c = b200.compress(data, level)
        assert b200.decompress(c) == data
    # the full 9-byte chunk separator inside stored (incompressible) user data: candidates are found that
    # are not chunk starts; validation must reject the optimistic layout and still decode correctly
    sep = b"\x00\x00\xff\xff\x00\x00\x00\xff\xff"
    noisy = b"".join(datagen.random_bytes(3000, seed=50 + i) + sep for i in range(60)) + datagen.text_like(100000)
    for level in (0, 2, 3):
        c = b200.compress(noisy, level)
        assert b200.decompress(c) == noisy
    z = datagen.foreign_streams(datagen.random_bytes(200000) + b"\x00\x00\xff\xff" * 1000)["stored"]
    assert b200.decompress(z) == datagen.random_bytes(200000) + b"\x00\x00\xff\xff" * 1000


# ---- segment index (include/b200_deflate.h, csrc/common.cuh) -------------------------------------------
INDEX_GROUPS, INDEX_BYTES = 64, 320


def _index_at(c, off):
    """Parses the 64 index groups at c[off:]; returns the 16 words or None."""
    words = [0] * 16
    for g in range(INDEX_GROUPS):
        b = c[off + 5 * g: off + 5 * g + 5]
        if len(b) < 5 or (b[0] & 0x87) != 0x80 or b[1:] != b"\x00\x00\xff\xff":
            return None
        words[g >> 2] |= ((b[0] >> 3) & 15) << (4 * (g & 3))
    return words


def _compress_dev(b200, data, level, flags=0):
    import torch
    ctx = b200.Context(0)
    src = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
    cap = b200.deflate_bound(len(data))
    dst = torch.zeros(cap, dtype=torch.uint8, device="cuda")
    n = ctx.compress_dev(src.data_ptr(), len(data), level, dst.data_ptr(), cap, flags=flags)
    torch.cuda.synchronize()
    return bytes(dst[:n].cpu().numpy())


@pytest.mark.parametrize("level", [1, 2, 3])
def test_segment_index_streams(b200, oracle, level):
    """Full Huffman-coded chunks start with the segment index: 64 empty stored blocks that zlib, the
    reference inflater and its restatement skip, and that this inflater uses to decode 16 segments in
    parallel.  Without it (B200_F_NO_INDEX) the stream is exactly 320 bytes per indexed chunk smaller and
    decodes to the same bytes."""
    data = datagen.text_like(5 * 65536 + 1234, seed=3)
    c = _compress_dev(b200, data, level)
    plain = _compress_dev(b200, data, level, flags=b200.F_NO_INDEX)
    words = _index_at(c, 0)
    assert words is not None and (words[0] & 0x3FF) == 0x2B5 and (words[0] >> 10) == 15
    assert all(0 < w < 65536 for w in words[1:])
    assert _index_at(plain, 0) is None
    assert len(c) - len(plain) == 5 * INDEX_BYTES          # five full chunks; the partial one is a single segment: no index
    for s in (c, plain):
        out, unused = zlib_raw_inflate(s)
        assert out == data and unused == b""
        rc, o = oracle.inflate(s)
        assert rc == 0 and o == data
        assert b200.decompress(s) == data
        assert b200.decompress(s, out_size=100000) == data[:100000]     # truncating overload


@pytest.mark.parametrize("level", [2, 3])
def test_segment_index_short_chunk(b200, oracle, ref, level):
    """The short last chunk of a stream carries an index of its own size -- four groups per segment, word 0 says how
    many -- if it has 8 segments or more, so that it is decoded by one thread per segment too; shorter ones (test.bmp: 6
    segments) have none and go to the one-warp decoder with the shared-memory window, which is faster there."""
    from conftest import gold
    for data in (gold("test.bmp"), datagen.text_like(65536 + 9 * 4096 + 17, seed=9), datagen.text_like(40000, seed=10),
                 datagen.text_like(65536 + 3 * 4096 + 17, seed=11), datagen.text_like(7 * 4096 + 1, seed=12)):
        nseg = (len(data) % 65536 + 4095) // 4096
        c = _compress_dev(b200, data, level)
        plain = _compress_dev(b200, data, level, flags=b200.F_NO_INDEX)
        full = len(data) // 65536
        assert len(c) - len(plain) == full * INDEX_BYTES + (20 * nseg if nseg >= 8 else 0)
        if full == 0 and nseg >= 8:
            words = [0] * nseg
            for g in range(4 * nseg):
                b = c[5 * g: 5 * g + 5]
                assert (b[0] & 0x87) == 0x80 and b[1:] == b"\x00\x00\xff\xff"
                words[g >> 2] |= ((b[0] >> 3) & 15) << (4 * (g & 3))
            assert (words[0] & 0x3FF) == 0x2B5 and (words[0] >> 10) == nseg - 1
            assert all(0 < w < 65536 for w in words[1:])
        for s in (c, plain):
            out, unused = zlib_raw_inflate(s)
            assert out == data and unused == b""
            rc, o = oracle.inflate(s)
            assert rc == 0 and o == data
            n, r_out = ref.inflate(s)
            assert n == len(data) and r_out == data
            assert b200.decompress(s) == data
            assert b200.decompress(s, out_size=len(data)) == data
            assert b200.decompress(s, out_size=5000) == data[:5000]


def test_segment_index_not_trusted(b200):
    """The index lives in bits every DEFLATE decoder must ignore, so a stream with a wrong index is
    still a valid stream with the same content: the decoder has to notice and fall back."""
    data = datagen.text_like(3 * 65536, seed=4) + datagen.image_like(2 * 65536)
    c = bytearray(_compress_dev(b200, data, 2))
    assert _index_at(c, 0) is not None
    for g, flip in ((7, 0x08), (20, 0x40), (63, 0x10)):    # segment lengths of chunk 0
        bad = bytearray(c)
        bad[5 * g] ^= flip
        assert zlib_raw_inflate(bytes(bad))[0] == data
        assert b200.decompress(bytes(bad)) == data
    bad = bytearray(c)
    bad[0] ^= 0x08                                          # magic gone: chunk 0 is simply not indexed
    assert b200.decompress(bytes(bad)) == data
    # small inputs and chunks that barely compress carry no index
    assert _index_at(_compress_dev(b200, gold("test.bmp"), 2), 0) is None
    noisy = datagen.random_bytes(65536 * 2, seed=8)
    assert _index_at(_compress_dev(b200, noisy, 1), 0) is None


@pytest.mark.parametrize("overlap", ["0", "1"])
def test_chunk_groups(b200, monkeypatch, overlap):
    """Long streams are inflated in groups of chunks (bounded scratch, optionally the copy pass of one group
    overlapping pass A of the next): force tiny groups so that an 11-chunk stream takes six of them."""
    import torch
    monkeypatch.setenv("B200_INFLATE_GROUP", "2")
    monkeypatch.setenv("B200_INFLATE_OVERLAP", overlap)
    data = datagen.text_like(4 * 65536, seed=21) + datagen.random_bytes(2 * 65536, seed=22) + datagen.image_like(4 * 65536 + 777)
    c = _compress_dev(b200, data, 2)
    ctx = b200.Context(0)
    src = torch.frombuffer(bytearray(c), dtype=torch.uint8).cuda()
    for cap in (len(data), len(data) + 5000, 3 * 65536 + 11):
        dst = torch.zeros(cap + 16, dtype=torch.uint8, device="cuda")
        w, full = ctx.inflate_dev(src.data_ptr(), len(c), dst.data_ptr(), cap)
        torch.cuda.synchronize()
        assert full == len(data) and w == min(cap, len(data))
        assert bytes(dst[:w].cpu().numpy()) == data[:w]
        assert int(dst[cap:].sum()) == 0                     # nothing written past the caller's capacity


# ---- zlib framing: Adler-32 on the device, strict header / trailer checks ------------
def test_adler32_device(b200):
    import torch
    ctx = b200.Context(0)
    src = datagen.text_like(3 * 65536 + 999, seed=31) + datagen.random_bytes(2 * 65536 + 1, seed=32) + b"\xff" * 70000
    buf = torch.frombuffer(bytearray(b"\x00" + src), dtype=torch.uint8).cuda()      # +1: an unaligned view below
    for n in (0, 1, 15, 16, 17, 255, 65535, 65536, 65537, 200000, len(src)):
        assert ctx.adler32_dev(buf.data_ptr() + 1, n) == zlib.adler32(src[:n]), n
    aligned = torch.frombuffer(bytearray(src), dtype=torch.uint8).cuda()
    for n in (16, 4096, 65536, 65536 * 2 + 48, len(src)):
        assert ctx.adler32_dev(aligned.data_ptr(), n) == zlib.adler32(src[:n]), n


def test_zlib_strict_checks_header_and_trailer(b200):
    """The reference skips two bytes and ignores the Adler-32 trailer (inflate.hpp:326-361); so does this library
    unless B200_F_STRICT is set, in which case header check bits and the trailer (computed on the GPU over the
    decoded bytes) must hold."""
    data = datagen.text_like(300000, seed=33)
    z = zlib.compress(data, 6)
    assert b200.decompress_zlib(z) == data
    assert b200.decompress_zlib(z, flags=b200.F_STRICT) == data
    assert b200.decompress_zlib(z, out_size=len(data), flags=b200.F_STRICT) == data
    for name in ("zlib.dat", "weird.dat"):                     # the reference's own fixtures carry valid trailers
        assert b200.decompress_zlib(gold(name), flags=b200.F_STRICT) == zlib.decompress(gold(name))
    bad_trailer = z[:-1] + bytes([z[-1] ^ 1])
    assert b200.decompress_zlib(bad_trailer) == data             # like the reference: not looked at
    with pytest.raises(b200.B200Error) as e:
        b200.decompress_zlib(bad_trailer, flags=b200.F_STRICT)
    assert e.value.code == 2
    bad_header = bytes([z[0], z[1] ^ 1]) + z[2:]
    assert b200.decompress_zlib(bad_header) == data
    with pytest.raises(b200.B200Error):
        b200.decompress_zlib(bad_header, flags=b200.F_STRICT)
    # a GPU-compressed raw stream wrapped as zlib by hand
    raw = b200.compress(data, 2)
    wrapped = b"\x78\x9c" + raw + zlib.adler32(data).to_bytes(4, "big")
    assert b200.decompress_zlib(wrapped, flags=b200.F_STRICT) == data
    assert zlib.decompress(wrapped) == data


def test_host_inflate_pipelined():
    """The host-buffer inflate call streams long inputs in slices (H2D of slice k + 1, kernels of slice k, D2H of
    slice k - 1 overlap).  The default slice is 256 MiB; run a child process with 1 MiB slices so that a 20 MB
    input takes the pipelined path, including truncation and streams it has to hand back to the plain path."""
    import os
    import subprocess
    import sys
    from conftest import ROOT
    script = r"""
import sys, zlib
sys.path.insert(0, %r)
sys.path.insert(0, %r)
import deflate_hpp_b200 as b200
import datagen
data = datagen.text_like(7 * 1024 * 1024, seed=41) + datagen.random_bytes(3 * 1024 * 1024 + 17, seed=42) + datagen.image_like(9 * 1024 * 1024 + 5)
for level in (2, 0):
    c = b200.compress(data, level)
    assert len(c) > 4 * 1024 * 1024 or level == 2
    assert b200.decompress(c, out_size=len(data)) == data
    assert b200.decompress(c, out_size=len(data) + 12345) == data
    assert b200.decompress(c, out_size=5 * 1024 * 1024 + 3) == data[:5 * 1024 * 1024 + 3]
    assert b200.decompress(c) == data
z = zlib.compressobj(1, zlib.DEFLATED, -15)
foreign = z.compress(data) + z.flush()
assert b200.decompress(foreign, out_size=len(data)) == data          # no chunk structure: plain path
cut = b200.compress(data, 2)[:6 * 1024 * 1024]
try:
    b200.decompress(cut, out_size=len(data))
    raise SystemExit("truncated stream did not raise")
except b200.B200Error as e:
    assert e.code in (1, 2)
print("ok", b200.launch_count())
""" % (ROOT, os.path.join(ROOT, "tests"))
    env = dict(os.environ, B200_HOST_INFLATE_SLICE=str(1 << 20))
    r = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout[-2000:] + r.stderr[-4000:]


def test_zlib_strict_with_truncating_buffer(b200):
    """with a caller buffer smaller than the decoded size, strict mode must still verify the trailer
    (the whole stream is decoded on the device; only the copy back is truncated)."""
    data = datagen.text_like(300000, seed=35)
    z = zlib.compress(data, 6)
    assert b200.decompress_zlib(z, out_size=1000, flags=b200.F_STRICT) == data[:1000]
    bad = z[:-1] + bytes([z[-1] ^ 1])
    assert b200.decompress_zlib(bad, out_size=1000) == data[:1000]            # reference behaviour: trailer ignored
    with pytest.raises(b200.B200Error) as e:
        b200.decompress_zlib(bad, out_size=1000, flags=b200.F_STRICT)
    assert e.value.code == 2


def test_one_warp_decoder_beyond_4gib(b200, monkeypatch):
    """the sequential one-warp decoder keeps 64-bit output positions (rebased every 2 GiB).  The stream is
    built on the device: 4.25 GiB in stored blocks of 65535 bytes, then a zlib-6 coded tail (literals and matches
    decoded with positions beyond 4 GiB).  Decoded by ONE warp (B200_INFLATE_SEQUENTIAL=1): exact size, right bytes
    on both sides of the 2 GiB and 4 GiB marks and in the tail; a smaller capacity truncates and still reports the
    full size."""
    import torch
    nblocks = (17 << 28) // 65535 + 1
    nstored = nblocks * 65535
    tail = datagen.text_like(300000, seed=77)
    ztail = zlib.compressobj(6, zlib.DEFLATED, -15)
    ztail = ztail.compress(tail) + ztail.flush()
    dev = "cuda"
    payload = (torch.arange(nstored, device=dev, dtype=torch.int64) * 2654435761 >> 7).to(torch.uint8)
    stream = torch.empty(nblocks * 65540 + len(ztail), dtype=torch.uint8, device=dev)
    blocks = stream[:nblocks * 65540].view(nblocks, 65540)
    blocks[:, 0] = 0
    blocks[:, 1:3] = 0xFF
    blocks[:, 3:5] = 0
    blocks[:, 5:] = payload.view(nblocks, 65535)
    stream[nblocks * 65540:] = torch.frombuffer(bytearray(ztail), dtype=torch.uint8).to(dev)
    n = nstored + len(tail)
    assert n > (1 << 32) + (1 << 28)
    monkeypatch.setenv("B200_INFLATE_SEQUENTIAL", "1")
    ctx = b200.Context(0)
    out = torch.zeros(n + 64, dtype=torch.uint8, device=dev)
    w, full = ctx.inflate_dev(stream.data_ptr(), stream.numel(), out.data_ptr(), n + 64)
    assert w == full == n
    for lo in range(0, nstored, 1 << 30):
        assert torch.equal(out[lo:min(nstored, lo + (1 << 30))], payload[lo:min(nstored, lo + (1 << 30))]), lo
    assert bytes(out[nstored:n].cpu().numpy()) == tail
    assert int(out[n:].sum().item()) == 0
    out.zero_()
    cap = (1 << 32) + 5
    w, full = ctx.inflate_dev(stream.data_ptr(), stream.numel(), out.data_ptr(), cap)
    assert w == cap and full == n
    assert torch.equal(out[cap - 4096:cap], payload[cap - 4096:cap]) and int(out[cap:cap + 4096].sum().item()) == 0


def test_inflate_shard_windows(b200):
    """b200_inflate_shard_dev: a stream of this library cut into byte ranges; every range (plus slack) decodes exactly the
    chunks that START inside it, the pieces tile the output (what ShardedDeflate.inflate does on N GPUs)."""
    import torch
    nchunks = 40
    n = nchunks * b200.CHUNK - 777                            # the last chunk is partial
    ctx = b200.Context(0)
    src = torch.empty(nchunks * b200.CHUNK, dtype=torch.uint8, device="cuda")
    ctx.corpus_generate_dev(src.data_ptr(), 20261018, 5, nchunks)
    cap = b200.deflate_bound(n)
    dst = torch.empty(cap, dtype=torch.uint8, device="cuda")
    cn = ctx.compress_dev(src.data_ptr(), n, 2, dst.data_ptr(), cap)
    slack = b200.CHUNK + 4096
    for world in (1, 2, 3, 7, 64):
        pos = 0
        for r in range(world):
            lo, hi = cn * r // world, cn * (r + 1) // world
            ws = 0 if r == 0 else max(0, (lo - 16) & ~15)
            we = min(cn, hi + slack)
            win = dst[ws:we].clone()
            out = torch.zeros(n + 64, dtype=torch.uint8, device="cuda")
            out_n, nch, nxt = ctx.inflate_shard_dev(win.data_ptr(), we - ws, lo - ws, hi - ws, r == 0, we == cn, out.data_ptr(), n + 64)
            assert torch.equal(out[:out_n], src[pos:pos + out_n]), (world, r)
            assert out_n == nch * b200.CHUNK or pos + out_n == n
            pos += out_n
        assert pos == n, world
