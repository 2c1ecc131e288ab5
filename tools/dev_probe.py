"""Development probe (not a pytest): quick end-to-end check + rough timings on a GPU box.

    gpurun -- 'python tools/dev_probe.py [MiB]'
"""
import ctypes
import hashlib
import os
import sys
import time
import zlib

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import deflate_hpp_b200 as d  # noqa: E402

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
GOLD = os.path.join(ROOT, "tests", "golden")


def load_ref():
    p = os.path.join(ROOT, "oracle", "_ref", "libref_deflate.so")
    if not os.path.exists(p):
        return None
    r = ctypes.CDLL(p)
    r.ref_quiet(1)
    r.ref_inflate.restype = ctypes.c_longlong
    r.ref_inflate.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t]
    r.ref_compress.restype = ctypes.c_longlong
    r.ref_compress.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t]
    return r


REF = load_ref()


def ref_inflate(c, cap):
    buf = ctypes.create_string_buffer(cap + 16)
    n = REF.ref_inflate(c, len(c), buf, cap + 16)
    return None if n < 0 else buf.raw[:n]


def zraw(c):
    o = zlib.decompressobj(-15)
    out = o.decompress(c)
    assert o.eof, "zlib: stream not terminated"
    return out


def check_roundtrip(name, data, level):
    c = d.compress(data, level)
    z = zraw(c)
    ok_z = z == data
    ok_r = None
    if REF is not None and len(data) <= 1 << 20:
        ok_r = ref_inflate(c, len(data)) == data
    g = d.decompress(c)
    ok_g = g == data
    print(f"  {name:28s} L{level} n={len(data):9d} -> {len(c):9d} ({len(c) / max(1, len(data)):.4f}) "
          f"zlib={ok_z} ref={ok_r} gpu={ok_g}", flush=True)
    return ok_z and ok_g and (ok_r is not False)


def main():
    mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    print(torch.cuda.get_device_name(0), flush=True)
    ok = True
    import numpy as np
    rng = np.random.default_rng(1)
    cases = {
        "empty": b"",
        "one": b"A",
        "abc": b"abc",
        "run1000": b"A" * 1000,
        "zeros100k": bytes(100000),
        "tiny.bmp": open(os.path.join(GOLD, "tiny.bmp"), "rb").read(),
        "test.bmp": open(os.path.join(GOLD, "test.bmp"), "rb").read(),
        "rand70000": rng.integers(0, 256, 70000, dtype=np.uint8).tobytes(),
        "text200k": (b"the quick brown fox jumps over the lazy dog. " * 5000)[:200000],
        "exact64k": rng.integers(0, 4, 65536, dtype=np.uint8).tobytes(),
        "exact128k": rng.integers(0, 16, 131072, dtype=np.uint8).tobytes(),
    }
    for name, data in cases.items():
        for level in (0, 1, 2, 3):
            try:
                ok &= check_roundtrip(name, data, level)
            except Exception as e:  # noqa: BLE001
                ok = False
                print(f"  {name} L{level}: EXC {e!r}", flush=True)
    # inflate of foreign streams
    for f in ("zlib.dat", "weird.dat"):
        raw = open(os.path.join(GOLD, f), "rb").read()
        try:
            out = d.decompress_zlib(raw)
            good = out == zlib.decompress(raw)
            print(f"  decompressZlib({f}) -> {len(out)} sha1={hashlib.sha1(out).hexdigest()[:12]} ok={good}", flush=True)
            ok &= good
        except Exception as e:  # noqa: BLE001
            ok = False
            print(f"  {f}: EXC {e!r}", flush=True)
    text = cases["text200k"] + cases["test.bmp"] * 3
    for lvl, strat in ((1, 0), (6, 0), (9, 0), (6, zlib.Z_FIXED), (0, 0), (6, zlib.Z_HUFFMAN_ONLY), (6, zlib.Z_RLE)):
        co = zlib.compressobj(lvl, zlib.DEFLATED, -15, 8, strat)
        c = co.compress(text) + co.flush()
        try:
            out = d.decompress(c)
            print(f"  zlib L{lvl} strat{strat}: {len(c)} -> {len(out)} ok={out == text}", flush=True)
            ok &= out == text
        except Exception as e:  # noqa: BLE001
            ok = False
            print(f"  zlib L{lvl} strat{strat}: EXC {e!r}", flush=True)
    print("CORRECTNESS", "PASS" if ok else "FAIL", flush=True)

    # ---- timings on the synthetic corpus, device resident ----
    nchunks = mib * 16
    n = nchunks * d.CHUNK
    ctx = d.Context(0)
    src = torch.empty(n, dtype=torch.uint8, device="cuda")
    ctx.corpus_generate_dev(src.data_ptr(), 20261018, 0, nchunks)
    torch.cuda.synchronize()
    cap = d.deflate_bound(n)
    dst = torch.empty(cap, dtype=torch.uint8, device="cuda")
    back = torch.empty(n, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for level in (2, 3, 1, 0):
        cn = ctx.compress_dev(src.data_ptr(), n, level, dst.data_ptr(), cap, stream=st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        e0.record()
        for _ in range(reps):
            cn = ctx.compress_dev(src.data_ptr(), n, level, dst.data_ptr(), cap, stream=st)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"compress L{level}: {n / 2**20:.0f} MiB -> {cn} ({cn / n:.4f})  {ms:.2f} ms  {n / ms / 1e6:.2f} GB/s", flush=True)
        w, full = ctx.inflate_dev(dst.data_ptr(), cn, back.data_ptr(), n, stream=st)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            w, full = ctx.inflate_dev(dst.data_ptr(), cn, back.data_ptr(), n, stream=st)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        same = bool(torch.equal(back, src))
        print(f"inflate  L{level}: {full} bytes  {ms:.2f} ms  {n / ms / 1e6:.2f} GB/s  equal={same}", flush=True)
        if level == 2 and mib <= 64:
            host = dst[:cn].cpu().numpy().tobytes()
            t = time.time()
            z = zraw(host)
            print(f"  zlib check of GPU stream: {z == src.cpu().numpy().tobytes()} ({time.time() - t:.1f}s)", flush=True)


if __name__ == "__main__":
    main()
