"""Shared fixtures.  `-m "not gpu"` runs here on CPU; `-m gpu` runs on a B200 and calls the CUDA path
through the C ABI.  The oracle (oracle/) is the checker only -- it is never the thing under test in
the gpu tests and never on the product path."""
import ctypes
import os
import subprocess
import sys
import zlib

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _make_oracle():
    so = os.path.join(ROOT, "oracle", "liboracle.so")
    srcs = [os.path.join(ROOT, "oracle", f) for f in os.listdir(os.path.join(ROOT, "oracle")) if f.endswith(".c")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    return so


class Oracle:
    """ctypes view of oracle/liboracle.so (plain-C restatement of the reference algorithms)."""

    def __init__(self):
        self.lib = ctypes.CDLL(_make_oracle())
        L = self.lib
        L.oracle_inflate.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p),
                                     ctypes.POINTER(ctypes.c_size_t)]
        L.oracle_inflate_zlib.argtypes = L.oracle_inflate.argtypes
        L.oracle_free.argtypes = [ctypes.c_void_p]
        L.oracle_corpus_generate.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_char_p]

    def _run(self, fn, data):
        p, n = ctypes.c_void_p(), ctypes.c_size_t()
        rc = fn(data, len(data), ctypes.byref(p), ctypes.byref(n))
        out = ctypes.string_at(p.value, n.value) if p.value else b""
        self.lib.oracle_free(p)
        return rc, out

    def inflate(self, data):
        """-> (rc, bytes).  rc 0 ok, -1 overrun (reference throws), -2 invalid data."""
        return self._run(self.lib.oracle_inflate, data)

    def inflate_zlib(self, data):
        return self._run(self.lib.oracle_inflate_zlib, data)

    def corpus(self, seed, first_chunk, nchunks):
        buf = ctypes.create_string_buffer(nchunks * 65536)
        self.lib.oracle_corpus_generate(seed, first_chunk, nchunks, buf)
        return buf.raw


class Reference:
    """ctypes view of oracle/_ref/libref_deflate.so: the UNMODIFIED reference compiled from
    /root/reference (see oracle/Makefile).  Absent -> tests that need it are skipped."""

    def __init__(self, path):
        self.lib = ctypes.CDLL(path)
        L = self.lib
        L.ref_quiet(1)
        for f in ("ref_compress",):
            getattr(L, f).restype = ctypes.c_longlong
            getattr(L, f).argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t]
        for f in ("ref_inflate", "ref_inflate_into", "ref_inflate_zlib"):
            getattr(L, f).restype = ctypes.c_longlong
            getattr(L, f).argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t]

    def compress(self, data, level):
        cap = len(data) * 2 + 1024
        buf = ctypes.create_string_buffer(cap)
        n = self.lib.ref_compress(data, len(data), level, buf, cap)
        assert 0 <= n <= cap
        return buf.raw[:n]

    def _inf(self, fn, data, cap):
        buf = ctypes.create_string_buffer(cap + 1)
        n = fn(data, len(data), buf, cap)
        return (n, None) if n < 0 else (n, buf.raw[:min(n, cap)])

    def inflate(self, data, cap=1 << 24):
        """-> (size or -1 if the reference threw, bytes or None)."""
        return self._inf(self.lib.ref_inflate, data, cap)

    def inflate_zlib(self, data, cap=1 << 24):
        return self._inf(self.lib.ref_inflate_zlib, data, cap)


@pytest.fixture(scope="session")
def oracle():
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    if os.path.isdir("/root/reference/include"):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True, capture_output=True)
    p = os.path.join(ROOT, "oracle", "_ref", "libref_deflate.so")
    if not os.path.exists(p):
        pytest.skip("oracle/_ref/libref_deflate.so not built (no /root/reference here)")
    return Reference(p)


def gold(name):
    with open(os.path.join(GOLD, name), "rb") as f:
        return f.read()


def zlib_raw_inflate(data):
    """Independent third-party decoder (zlib 1.3): raw RFC 1951 stream, must end exactly at BFINAL."""
    o = zlib.decompressobj(-15)
    out = o.decompress(data)
    assert o.eof, "zlib: stream not terminated"
    return out, o.unused_data


def zlib_raw_deflate(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, wbits=-15, flush_every=None, flush_mode=None):
    co = zlib.compressobj(level, zlib.DEFLATED, wbits, 8, strategy)
    if flush_every is None:
        return co.compress(data) + co.flush()
    out = b""
    for i in range(0, len(data), flush_every):
        out += co.compress(data[i:i + flush_every]) + co.flush(flush_mode)
    return out + co.flush()


@pytest.fixture(scope="session")
def b200():
    """The product, through its C ABI (ctypes mirror).  Import fails loudly if the .so is missing."""
    import deflate_hpp_b200
    return deflate_hpp_b200
