"""Deterministic test inputs and foreign-stream producers shared by the CPU and GPU tests."""
import zlib

import numpy as np

EDGE_SIZES = [0, 1, 2, 3, 4, 5, 31, 32, 33, 257, 258, 259, 260, 8191, 8192, 8193, 32767, 32768, 32769,
              65535, 65536, 65537, 131072, 200001]


def text_like(n, seed=1):
    rng = np.random.default_rng(seed)
    words = [bytes(rng.integers(97, 123, rng.integers(2, 10), dtype=np.uint8)) for _ in range(512)]
    out = bytearray()
    while len(out) < n:
        out += words[int(rng.integers(0, 512) ** 2 // 512)] + b" "
    return bytes(out[:n])


def image_like(n, seed=2):
    rng = np.random.default_rng(seed)
    i = np.arange(n, dtype=np.int64)
    pix, comp = i // 3, i % 3
    x, y = pix % 1024, pix // 1024
    v = ((x >> 2) + (y >> 2)) & 255
    noise = rng.integers(0, 3, n)
    b = np.where(comp == 0, v + noise, np.where(comp == 1, 2 * v, 255 - v)) & 255
    return b.astype(np.uint8).tobytes()


def random_bytes(n, seed=3):
    return np.random.default_rng(seed).integers(0, 256, n, dtype=np.uint8).tobytes()


def runs(n, seed=4):
    rng = np.random.default_rng(seed)
    out = bytearray()
    while len(out) < n:
        out += bytes([int(rng.integers(0, 256))]) * int(rng.integers(1, 700))
    return bytes(out[:n])


def low_entropy(n, seed=5, k=4):
    return np.random.default_rng(seed).integers(0, k, n, dtype=np.uint8).tobytes()


KINDS = {"text": text_like, "image": image_like, "random": random_bytes, "runs": runs, "lowent": low_entropy}


def foreign_streams(data):
    """name -> raw RFC 1951 stream of `data` from an independent producer (zlib 1.3)."""
    def z(level, strategy=zlib.Z_DEFAULT_STRATEGY, flush_every=None, mode=None):
        co = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
        if flush_every is None:
            return co.compress(data) + co.flush()
        out = b""
        for i in range(0, len(data), flush_every):
            out += co.compress(data[i:i + flush_every]) + co.flush(mode)
        return out + co.flush()

    return {
        "zlib1": z(1), "zlib6": z(6), "zlib9": z(9), "fixed": z(6, zlib.Z_FIXED), "stored": z(0),
        "huffman_only": z(6, zlib.Z_HUFFMAN_ONLY), "rle": z(6, zlib.Z_RLE),
        "sync_flush_8k": z(6, flush_every=8192, mode=zlib.Z_SYNC_FLUSH),
        "full_flush_8k": z(6, flush_every=8192, mode=zlib.Z_FULL_FLUSH),
        "full_flush_64k": z(6, flush_every=65536, mode=zlib.Z_FULL_FLUSH),
    }


def too_far_stream():
    """Fixed-Huffman block: literal 'a', then a match (len 3, dist 4) that reaches before the start of
    the output, then EOB.  zlib rejects it; the reference silently copies nothing (inflate.hpp:268-270)."""
    bits = []

    def put(v, n, rev=False):
        for i in (range(n - 1, -1, -1) if rev else range(n)):
            bits.append((v >> i) & 1)
    put(1, 1); put(1, 2)                      # BFINAL, BTYPE=01
    put(0x30 + 0x61, 8, rev=True)             # literal 'a'
    put(0b0000001, 7, rev=True)               # length symbol 257 (len 3)
    put(3, 5, rev=True)                       # distance symbol 3 (dist 4 > 1 byte produced)
    put(0, 7, rev=True)                       # EOB
    while len(bits) % 8:
        bits.append(0)
    return bytes(sum(b << i for i, b in enumerate(bits[k:k + 8])) for k in range(0, len(bits), 8))
