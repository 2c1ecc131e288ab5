"""Small-input latency probe: which kernels a short stream goes through and what one call costs.
    python tools/probe_small.py"""
import os
import statistics
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import deflate_hpp_b200 as d  # noqa: E402

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import datagen  # noqa: E402

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
bmp = open(os.path.join(ROOT, "tests", "golden", "test.bmp"), "rb").read()
cases = {"test.bmp": bmp, "text 100 KiB": datagen.text_like(100 << 10), "text 9 KiB": datagen.text_like(9 << 10)}
ctx = d.Context(0)
for name, data in cases.items():
    n = len(data)
    src = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
    cap = d.deflate_bound(n)
    dst = torch.empty(cap + 64, dtype=torch.uint8, device="cuda")
    back = torch.zeros(n + 64, dtype=torch.uint8, device="cuda")
    for level in (2, 3):
        for flags in (0, d.F_NO_INDEX):
            cn = ctx.compress_dev(src.data_ptr(), n, level, dst.data_ptr(), cap, flags=flags)
            for env in ({}, {"B200_NO_SMALL_INFLATE": "1"}):
                os.environ.update(env)
                ts = []
                for i in range(30):
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    w, full = ctx.inflate_dev(dst.data_ptr(), cn, back.data_ptr(), n)
                    ts.append(time.perf_counter() - t0)
                ok = bool(full == n and torch.equal(back[:n], src))
                ctx.profile(True)
                ctx.inflate_dev(dst.data_ptr(), cn, back.data_ptr(), n)
                ctx.profile(False)
                k = {a: round(b[0] * 1000, 1) for a, b in ctx.profile_read().items()}
                for e in env:
                    del os.environ[e]
                print(f"{name:14s} level {level} {'no index' if flags else 'index   '} {'regular' if env else 'default'}: {cn:6d} B, "
                      f"inflate_dev median {statistics.median(ts) * 1e6:7.1f} us, ok {ok}, kernels us {k}", flush=True)
