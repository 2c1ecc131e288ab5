"""Round-2 tuning probe (not a test, not the bench): one GPU call, many settings.

    gpurun -- 'python tools/probe_r02.py [sections]'     sections: better,fast,foreign,batch (default all)
Context knobs are read from the environment when a context is created, so every setting gets a fresh Context.
"""
import json
import os
import sys
import time
import zlib

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import deflate_hpp_b200 as d  # noqa: E402
import bench  # noqa: E402

SEED, CHUNK = 20261018, 65536
want = set(sys.argv[1].split(",")) if len(sys.argv) > 1 else {"better", "fast", "foreign"}
dev = torch.device("cuda", 0)
nchunks = 16384
n = nchunks * CHUNK
src = torch.empty(n, dtype=torch.uint8, device=dev)
d.Context.corpus_generate_dev(src.data_ptr(), SEED, 0, nchunks)
cap = d.deflate_bound(n)
dst = torch.empty(cap + 4096, dtype=torch.uint8, device=dev)
back = torch.empty(n, dtype=torch.uint8, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timed(fn, steps, warm=1):
    for _ in range(warm):
        r = fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        r = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, r


def with_env(env):
    for k, v in env.items():
        os.environ[k] = str(v)
    c = d.Context(0)
    for k in env:
        del os.environ[k]
    return c


def compress_point(env, level, steps=3):
    ctx = with_env(env)
    ms, cn = timed(lambda: ctx.compress_dev(src.data_ptr(), n, level, dst.data_ptr(), cap), steps)
    ctx.profile(True)
    ctx.compress_dev(src.data_ptr(), n, level, dst.data_ptr(), cap)
    ctx.profile(False)
    k = {a: round(b[0], 3) for a, b in ctx.profile_read().items()}
    w, full = ctx.inflate_dev(dst.data_ptr(), cn, back.data_ptr(), n)
    ok = bool(full == n and torch.equal(back, src))
    print(json.dumps({"env": env, "level": level, "GBps": round(n / ms / 1e6, 2), "ms": round(ms, 3), "ratio": round(cn / n, 5),
                      "round_trip": ok, "kernels_ms": k}), flush=True)
    ctx.close()


if "fast" in want:
    for env in ({}, {"B200_NO_SPLIT": 1}, {"B200_BATCH_CHUNKS": 8192}, {"B200_BATCH_CHUNKS": 16384}):
        compress_point(env, 2, steps=5)

if "better" in want:
    for env in ({}, {"B200_NO_SPLIT": 1}):
        compress_point(env, 3, steps=2)

if "foreign" in want:
    host = src.cpu().numpy()
    t0 = time.time()
    stream = bench.pigz_style_stream(host, 6)
    print("zlib-6 stream", len(stream), "bytes in", round(time.time() - t0, 1), "s", flush=True)
    import numpy as np
    comp = torch.from_numpy(np.frombuffer(stream, dtype=np.uint8).copy()).to(dev)
    for env in ({"B200_FOREIGN_TAB": 0}, {"B200_FOREIGN_TAB": 1}, {"B200_FOREIGN_TAB": 2}, {"B200_FOREIGN_TAB": 3}):
        ctx = with_env(env)
        ms, (w, full) = timed(lambda: ctx.inflate_dev(comp.data_ptr(), comp.numel(), back.data_ptr(), n), 3)
        ctx.profile(True)
        ctx.inflate_dev(comp.data_ptr(), comp.numel(), back.data_ptr(), n)
        ctx.profile(False)
        k = {a: (round(b[0], 3), b[1]) for a, b in ctx.profile_read().items()}
        print(json.dumps({"foreign": env, "GBps": round(n / ms / 1e6, 2), "ms": round(ms, 2), "ok": bool(full == n and torch.equal(back, src)),
                          "kernels": k}), flush=True)
        ctx.close()
    # a zlib level-1 and a level-9 stream of the first 256 MiB (different block shapes)
    for lv in (1, 9):
        st = bench.pigz_style_stream(host[:256 << 20], lv)
        cc = torch.from_numpy(np.frombuffer(st, dtype=np.uint8).copy()).to(dev)
        ctx = d.Context(0)
        ms, (w, full) = timed(lambda: ctx.inflate_dev(cc.data_ptr(), cc.numel(), back.data_ptr(), n), 3)
        print(json.dumps({"foreign_level": lv, "GBps": round((256 << 20) / ms / 1e6, 2), "ms": round(ms, 2),
                          "ok": bool(full == (256 << 20) and torch.equal(back[:256 << 20], src[:256 << 20]))}), flush=True)
        ctx.close()
print("probe done")
