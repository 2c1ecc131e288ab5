"""GPU tests for the wire formats next to the path: zlib (RFC 1950) and gzip (RFC 1952) framing
emitted by the compressor with the checksum of the INPUT computed on the device, CRC-32 / Adler-32 device kernels against
zlib's, and the gzip inflate entry point."""
import gzip
import zlib

import numpy as np
import pytest

from conftest import gold
import datagen

pytestmark = pytest.mark.gpu

SIZES = [0, 1, 5, 255, 256, 257, 4095, 65535, 65536, 65537, 200001, 1 << 20, 3 * (1 << 20) + 77]


def test_crc32_and_adler32_device(b200):
    import torch
    ctx = b200.Context(0)
    src = datagen.text_like(700000, seed=3) + datagen.random_bytes(900001, seed=4)
    buf = torch.frombuffer(bytearray(src) + bytearray(64), dtype=torch.uint8).cuda()
    for n in [0, 1, 2, 15, 16, 17, 255, 256, 257, 65535, 65536, 65537, 65536 * 3 + 5, len(src)]:
        assert ctx.crc32_dev(buf.data_ptr(), n) == zlib.crc32(src[:n]), n
        assert ctx.adler32_dev(buf.data_ptr(), n) == zlib.adler32(src[:n]), n
    for off in (1, 3, 8, 13):                                   # unaligned starts
        n = 300000
        assert ctx.crc32_dev(buf.data_ptr() + off, n) == zlib.crc32(src[off:off + n]), off


@pytest.mark.parametrize("level", [0, 1, 2, 3])
def test_zlib_and_gzip_emit(b200, level):
    """zlib.decompress / gzip.decompress (independent implementations) accept the framed streams and verify the
    checksums the GPU computed; the library's own strict inflaters do too."""
    src = datagen.text_like(2_000_000, seed=5) + datagen.image_like(1_500_000) + datagen.random_bytes(300_000)
    for n in SIZES:
        data = src[:n]
        z = b200.compress(data, level, flags=b200.F_ZLIB)
        assert z[:2] == b"\x78\x9c"
        assert zlib.decompress(z) == data, n
        assert z[-4:] == zlib.adler32(data).to_bytes(4, "big")
        assert b200.decompress_zlib(z, flags=b200.F_STRICT) == data
        g = b200.compress(data, level, flags=b200.F_GZIP)
        assert gzip.decompress(g) == data, n
        assert g[-8:-4] == zlib.crc32(data).to_bytes(4, "little") and g[-4:] == (n & 0xFFFFFFFF).to_bytes(4, "little")
        assert b200.decompress_gzip(g, flags=b200.F_STRICT) == data
        raw = b200.compress(data, level)
        assert z[2:-4] == raw and g[10:-8] == raw                  # the frame goes AROUND the very same raw stream


def test_gzip_inflate_foreign_members(b200):
    """.gz made by Python's gzip module (FNAME, mtime), with and without optional header fields"""
    data = datagen.text_like(900_000, seed=8)
    import io
    bio = io.BytesIO()
    with gzip.GzipFile(filename="some_name.txt", mode="wb", fileobj=bio, mtime=12345) as f:
        f.write(data)
    g = bio.getvalue()
    assert g[3] & 8                                             # FNAME present
    assert b200.decompress_gzip(g) == data
    assert b200.decompress_gzip(g, flags=b200.F_STRICT) == data
    assert b200.decompress_gzip(g, out_size=1000) == data[:1000]
    plain = gzip.compress(data, 6, mtime=0)
    assert b200.decompress_gzip(plain, flags=b200.F_STRICT) == data
    bad = plain[:-6] + bytes([plain[-6] ^ 1]) + plain[-5:]
    assert b200.decompress_gzip(bad) == data                    # trailer ignored unless strict (like decompressZlib)
    with pytest.raises(b200.B200Error) as e:
        b200.decompress_gzip(bad, flags=b200.F_STRICT)
    assert e.value.code == 2
    with pytest.raises(b200.B200Error):
        b200.decompress_gzip(b"\x1f\x8b\x07" + plain[3:])      # CM != 8
    with pytest.raises(b200.B200Error):
        b200.decompress_gzip(plain[:12])


def test_framed_device_api(b200):
    """b200_deflate_compress_dev with B200_F_ZLIB / B200_F_GZIP on device buffers; framing a shard is refused"""
    import torch
    ctx = b200.Context(0)
    nchunks = 300
    n = nchunks * b200.CHUNK - 12345
    src = torch.empty(nchunks * b200.CHUNK, dtype=torch.uint8, device="cuda")
    ctx.corpus_generate_dev(src.data_ptr(), 20261018, 0, nchunks)
    host = bytes(src[:n].cpu().numpy())
    cap = b200.deflate_bound(n)
    dst = torch.zeros(cap, dtype=torch.uint8, device="cuda")
    cn = ctx.compress_dev(src.data_ptr(), n, 2, dst.data_ptr(), cap, flags=b200.F_ZLIB)
    assert zlib.decompress(bytes(dst[:cn].cpu().numpy())) == host
    cn = ctx.compress_dev(src.data_ptr(), n, 2, dst.data_ptr(), cap, flags=b200.F_GZIP)
    assert gzip.decompress(bytes(dst[:cn].cpu().numpy())) == host
    with pytest.raises(b200.B200Error):
        ctx.compress_dev(src.data_ptr(), n, 2, dst.data_ptr(), cap, flags=b200.F_GZIP | b200.F_NOT_LAST)


def test_file_api_streaming(tmp_path):
    """deflate::compress(file, file, level) / inflate::decompress(file, file) as bounded-memory streaming pipelines
    (SURVEY.md 8(f) rank 3).  Run in a child process with 1 MiB slices so that a 7 MB file takes 8 slices / windows:
    the streamed output must be byte-identical to the in-memory call, framed variants must satisfy zlib / gzip, the
    windowed inflate must equal the input, and files it cannot window (foreign streams, separator bytes inside stored
    data) must still decode."""
    import os
    import subprocess
    import sys
    from conftest import ROOT
    script = r"""
import sys, zlib, gzip, os
sys.path.insert(0, %r)
sys.path.insert(0, %r)
import deflate_hpp_b200 as b200
import datagen
tmp = %r
data = datagen.text_like(3 * 1024 * 1024 + 11, seed=51) + datagen.random_bytes(2 * 1024 * 1024 + 7, seed=52) + datagen.image_like(2 * 1024 * 1024 + 333)
src = os.path.join(tmp, "in.bin")
open(src, "wb").write(data)
for level in (0, 2, 3):
    dst = os.path.join(tmp, "out.%%d" %% level)
    a, b = b200.compress_file(src, dst, level)
    c = open(dst, "rb").read()
    assert a == len(data) and b == len(c)
    assert c == b200.compress(data, level), level          # streamed in slices == in memory
    back = os.path.join(tmp, "back.%%d" %% level)
    a2, b2 = b200.decompress_file(dst, back)
    assert a2 == len(c) and b2 == len(data) and open(back, "rb").read() == data
dst = os.path.join(tmp, "out.z")
b200.compress_file(src, dst, 2, flags=b200.F_ZLIB)
assert zlib.decompress(open(dst, "rb").read()) == data
dst = os.path.join(tmp, "out.gz")
b200.compress_file(src, dst, 2, flags=b200.F_GZIP)
assert gzip.decompress(open(dst, "rb").read()) == data
# edge sizes: empty, tiny, exactly one slice, one slice + 1
for n in (0, 1, 70000, 1 << 20, (1 << 20) + 1, 2 << 20):
    open(src, "wb").write(data[:n])
    dst = os.path.join(tmp, "edge")
    b200.compress_file(src, dst, 2)
    c = open(dst, "rb").read()
    assert c == b200.compress(data[:n], 2), n
    b200.decompress_file(dst, dst + ".back")
    assert open(dst + ".back", "rb").read() == data[:n], n
# a foreign stream in a file (zlib level 6, 3+ MB compressed): not windowable, decoded as a whole
z = zlib.compressobj(6, zlib.DEFLATED, -15)
foreign = z.compress(data) + z.flush()
open(src, "wb").write(foreign)
b200.decompress_file(src, src + ".out")
assert open(src + ".out", "rb").read() == data
# stored user data that contains the chunk separator pattern: windows do not validate, the whole-file path takes over
sep = b"\x00\x00\xff\xff\x00\x00\x00\xff\xff"
evil = (datagen.random_bytes(100000, seed=5) + sep) * 40
c = b200.compress(evil, 2)
open(src, "wb").write(c)
b200.decompress_file(src, src + ".out")
assert open(src + ".out", "rb").read() == evil
try:
    b200.compress_file(os.path.join(tmp, "does_not_exist"), dst, 2)
    raise SystemExit("missing file did not raise")
except b200.B200Error as e:
    assert e.code == b200.api.E_IO if hasattr(b200, "api") else e.code == 7
print("ok", b200.launch_count())
""" % (ROOT, os.path.join(ROOT, "tests"), str(tmp_path))
    env = dict(os.environ, B200_FILE_SLICE=str(1 << 20))
    r = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout[-2000:] + r.stderr[-4000:]
