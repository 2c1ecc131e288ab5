"""GPU parity tests for the block-parallel inflater of FOREIGN single streams (csrc/inflate_foreign.cuh, SURVEY.md 8(f)
rank 1): streams this library did not write -- zlib at several levels and strategies, sync- / full-flushed, pigz-style
with preset dictionaries, and streams of the reference's own compressor -- must decode bit-exactly like zlib, the oracle
and the reference inflater do, through the same public calls."""
import zlib

import numpy as np
import pytest

from conftest import gold, ROOT
import datagen

pytestmark = pytest.mark.gpu


def mixed(n, seed=0):
    """text / image / random / runs interleaved in pieces of 100-300 KB: many blocks, cross-block references, stored blocks"""
    rng = np.random.default_rng(seed)
    kinds = [datagen.text_like, datagen.image_like, datagen.random_bytes, datagen.runs, datagen.low_entropy]
    out = bytearray()
    k = 0
    while len(out) < n:
        out += kinds[k % len(kinds)](int(rng.integers(100_000, 300_000)), seed=seed * 100 + k)
        k += 1
    return bytes(out[:n])


def raw(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, memlevel=8, flush_every=None, mode=zlib.Z_SYNC_FLUSH):
    co = zlib.compressobj(level, zlib.DEFLATED, -15, memlevel, strategy)
    if flush_every is None:
        return co.compress(data) + co.flush()
    out = b""
    for i in range(0, len(data), flush_every):
        out += co.compress(data[i:i + flush_every]) + co.flush(mode)
    return out + co.flush()


def launches_of(b200, kernel_id_name, fn):
    """run fn() with profiling on a fresh context-free call path: count launches through the global counter instead"""
    before = b200.launch_count()
    r = fn()
    return r, b200.launch_count() - before


@pytest.mark.parametrize("producer", ["zlib1", "zlib6", "zlib9", "mem9", "filtered", "rle", "huffman_only", "fixed", "sync_64k",
                                      "full_flush_100k", "stored"])
def test_foreign_parallel_vs_zlib(b200, oracle, producer, monkeypatch):
    data = mixed(6_000_000, seed=3)
    stream = {
        "zlib1": lambda: raw(data, 1), "zlib6": lambda: raw(data, 6), "zlib9": lambda: raw(data, 9),
        "mem9": lambda: raw(data, 6, memlevel=9), "filtered": lambda: raw(data, 6, zlib.Z_FILTERED),
        "rle": lambda: raw(data, 6, zlib.Z_RLE), "huffman_only": lambda: raw(data, 6, zlib.Z_HUFFMAN_ONLY),
        "fixed": lambda: raw(data, 6, zlib.Z_FIXED), "sync_64k": lambda: raw(data, 6, flush_every=65536),
        "full_flush_100k": lambda: raw(data, 6, flush_every=100_000, mode=zlib.Z_FULL_FLUSH), "stored": lambda: raw(data, 0),
    }[producer]()
    out = b200.decompress(stream)
    assert out == data
    # the same stream through the sequential one-warp decoder: identical
    monkeypatch.setenv("B200_NO_FOREIGN_PARALLEL", "1")
    assert b200.decompress(stream) == data
    monkeypatch.delenv("B200_NO_FOREIGN_PARALLEL")
    # caller-buffer overloads: exact size, larger, truncating (the truncating one falls back to the sequential decoder)
    assert b200.decompress(stream, out_size=len(data)) == data
    assert b200.decompress(stream, out_size=len(data) + 999) == data
    assert b200.decompress(stream, out_size=1_234_567) == data[:1_234_567]
    if producer in ("zlib6", "stored"):
        rc, o = oracle.inflate(stream)
        assert rc == 0 and o == data


@pytest.mark.parametrize("tab", [0, 1, 2, 3])
def test_foreign_table_variants(b200, tab, monkeypatch):
    """every table-width / CTA-size variant of the decode kernels (B200_FOREIGN_TAB) gives the same bytes: text (long
    literal codes -> the canonical search behind the direct table), image-like data and a Huffman-only stream"""
    import torch
    monkeypatch.setenv("B200_FOREIGN_TAB", str(tab))
    ctx = b200.Context(0)
    data = mixed(5_000_000, seed=21)
    for stream in (raw(data, 6), raw(data, 9, zlib.Z_FILTERED), raw(data[:2_000_000], 6, zlib.Z_HUFFMAN_ONLY)):
        want = zlib.decompressobj(-15).decompress(stream)
        comp = torch.frombuffer(bytearray(stream), dtype=torch.uint8).cuda()
        out = torch.zeros(len(want) + 64, dtype=torch.uint8, device="cuda")
        ctx.profile(True)
        w, full = ctx.inflate_dev(comp.data_ptr(), comp.numel(), out.data_ptr(), len(want) + 64)
        ctx.profile(False)
        assert "foreign_decode_kernel<emit>" in ctx.profile_read()
        assert w == full == len(want) and bytes(out[:w].cpu().numpy()) == want
    ctx.close()


def test_foreign_parallel_is_taken(b200):
    """the block-parallel kernels really run for a long zlib stream (and not for a short one)"""
    import torch
    data = mixed(4_000_000, seed=5)
    stream = raw(data, 6)
    ctx = b200.Context(0)
    comp = torch.frombuffer(bytearray(stream), dtype=torch.uint8).cuda()
    out = torch.zeros(len(data) + 64, dtype=torch.uint8, device="cuda")
    ctx.profile(True)
    w, full = ctx.inflate_dev(comp.data_ptr(), comp.numel(), out.data_ptr(), len(data) + 64)
    ctx.profile(False)
    k = ctx.profile_read()
    assert w == full == len(data) and bytes(out[:w].cpu().numpy()) == data
    for name in ("foreign_find_blocks_kernel", "foreign_decode_kernel<count>", "foreign_decode_kernel<emit>", "foreign_copy_kernel",
                 "foreign_window_kernels", "foreign_resolve_kernel"):
        assert name in k, (name, sorted(k))
    assert "inflate_batch_kernel" not in k                     # the sequential decoder was not needed
    small = raw(data[:100_000], 6)
    comp = torch.frombuffer(bytearray(small), dtype=torch.uint8).cuda()
    ctx.profile(True)
    w, full = ctx.inflate_dev(comp.data_ptr(), comp.numel(), out.data_ptr(), len(data))
    ctx.profile(False)
    assert w == 100_000 and "foreign_find_blocks_kernel" not in ctx.profile_read()


def test_foreign_pigz_style_with_dictionary(b200):
    """pieces compressed in parallel, each primed with the previous 32 KiB (matches cross piece borders), sync-flushed"""
    data = mixed(9_000_000, seed=7)
    piece = 1 << 20
    parts = []
    npieces = (len(data) + piece - 1) // piece
    for i in range(npieces):
        lo = i * piece
        zd = data[max(0, lo - 32768):lo]
        co = zlib.compressobj(6, zlib.DEFLATED, -15, 8, zlib.Z_DEFAULT_STRATEGY, zd) if zd else zlib.compressobj(6, zlib.DEFLATED, -15)
        parts.append(co.compress(data[lo:lo + piece]) + (co.flush() if i == npieces - 1 else co.flush(zlib.Z_SYNC_FLUSH)))
    stream = b"".join(parts)
    assert zlib.decompressobj(-15).decompress(stream) == data
    assert b200.decompress(stream) == data


def test_foreign_reference_compressed(b200, ref):
    """a long stream written by the REFERENCE compressor (level 0 stored and level 1; its blocks are 32 KB and joined at
    bit granularity) -- GPU inflater == reference inflater, parallel path and sequential path alike"""
    data = mixed(1_500_000, seed=9)
    for level in (0, 1):
        c = ref.compress(data, level)
        n, r_out = ref.inflate(c)
        assert n >= 0
        assert b200.decompress(c) == r_out


def test_foreign_zlib_framing_and_fixtures(b200):
    """decompressZlib over a long zlib-framed stream (strict mode checks the Adler-32 on the device), and the reference's
    fixtures still decode (they are short: sequential path)"""
    data = mixed(5_000_000, seed=11)
    z = zlib.compress(data, 6)
    assert b200.decompress_zlib(z) == data
    assert b200.decompress_zlib(z, flags=b200.F_STRICT) == data
    for name in ("zlib.dat", "weird.dat"):
        assert b200.decompress_zlib(gold(name)) == zlib.decompress(gold(name))


def test_foreign_errors_match_sequential(b200, monkeypatch):
    """damaged long streams: same error class and, where decoding goes on, the same bytes as the sequential decoder"""
    data = mixed(3_000_000, seed=13)
    stream = bytearray(raw(data, 6))
    rng = np.random.default_rng(5)
    cases = [bytes(stream[:len(stream) // 2]), bytes(stream[:len(stream) - 3])]
    for _ in range(6):
        s = bytearray(stream)
        p = int(rng.integers(1000, len(s) - 1000))
        s[p] ^= 1 << int(rng.integers(0, 8))
        cases.append(bytes(s))

    def run(s):
        try:
            return ("ok", b200.decompress(s, out_size=len(data) + 4096))
        except b200.B200Error as e:
            return ("err", e.code)
    for s in cases:
        got = run(s)
        monkeypatch.setenv("B200_NO_FOREIGN_PARALLEL", "1")
        want = run(s)
        monkeypatch.delenv("B200_NO_FOREIGN_PARALLEL")
        assert got == want


def test_foreign_too_far_quirk_in_long_stream(b200, oracle):
    """a distance that reaches before the start of the output (first unit): the reference copies nothing, strict rejects"""
    body = raw(mixed(2_000_000, seed=17), 6)
    # prepend a non-final fixed block: literal 'a', match(len 3, dist 4) [too far], then the real stream
    bits = []

    def put(v, n, rev=False):
        for i in (range(n - 1, -1, -1) if rev else range(n)):
            bits.append((v >> i) & 1)
    put(0, 1); put(1, 2)
    put(0x30 + 0x61, 8, rev=True)
    put(0b0000001, 7, rev=True)
    put(3, 5, rev=True)
    put(0, 7, rev=True)
    # stored empty block to realign to a byte boundary: BFINAL 0, BTYPE 00, pad, LEN 0, NLEN FFFF
    put(0, 3)
    while len(bits) % 8:
        bits.append(0)
    head = bytes(sum(b << i for i, b in enumerate(bits[k:k + 8])) for k in range(0, len(bits), 8)) + b"\x00\x00\xff\xff"
    stream = head + body
    rc, o = oracle.inflate(stream)
    assert rc == 0
    assert b200.decompress(stream) == o
    with pytest.raises(b200.B200Error):
        b200.decompress(stream, flags=b200.F_STRICT)
