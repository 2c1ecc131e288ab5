"""e2e probe of the host-buffer C ABI (pinned and pageable buffers); knobs come from the environment of THIS process
(the per-process default context reads them once).  usage: python tools/probe_e2e.py [MiB] [pinned|pageable|both] [raw]
`raw` adds what the PCIe link of this box moves for the same byte counts with nothing else going on (the ceiling of any e2e number)."""
import ctypes
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import deflate_hpp_b200 as d  # noqa: E402

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n = mib << 20
L = d.lib()
src = torch.empty(n, dtype=torch.uint8, device="cuda")
d.Context.corpus_generate_dev(src.data_ptr(), 20261018, 0, n // 65536)
cap = d.deflate_bound(n)
res = {"env": {k: v for k, v in os.environ.items() if k.startswith("B200_")}, "mib": mib}
kinds = ("pinned", "pageable")
if len(sys.argv) > 2 and sys.argv[2] in ("pinned", "pageable"):
    kinds = (sys.argv[2],)
if "raw" in sys.argv[2:]:
    # pinned 1 GiB one way, then n bytes in + 0.62 n bytes out at the same time (compress), and the reverse (inflate)
    a = torch.empty(n, dtype=torch.uint8).pin_memory()
    b = torch.empty(n, dtype=torch.uint8).pin_memory()
    da = torch.empty(n, dtype=torch.uint8, device="cuda")
    db = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    m = int(n * 0.6207)

    def timed(fn, reps=4):
        ts = []
        for i in range(reps + 1):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            if i:
                ts.append(time.perf_counter() - t0)
        return sum(ts) / len(ts)

    def h2d():
        with torch.cuda.stream(s1):
            da.copy_(a, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            b.copy_(db, non_blocking=True)

    def both_c():
        with torch.cuda.stream(s1):
            da.copy_(a, non_blocking=True)
        with torch.cuda.stream(s2):
            b[:m].copy_(db[:m], non_blocking=True)

    def both_i():
        with torch.cuda.stream(s1):
            da[:m].copy_(a[:m], non_blocking=True)
        with torch.cuda.stream(s2):
            b.copy_(db, non_blocking=True)

    res["raw_pcie"] = {"h2d_GBps": round(n / timed(h2d) / 1e9, 2), "d2h_GBps": round(n / timed(d2h) / 1e9, 2),
                       "compress_shape_ceiling_GBps": round(n / timed(both_c) / 1e9, 2),
                       "inflate_shape_ceiling_GBps": round(n / timed(both_i) / 1e9, 2),
                       "note": "ceilings: n bytes one way + 0.62 n the other way at the same time, per n"}
    del a, b, da, db
for kind in kinds:
    h_in = torch.empty(n, dtype=torch.uint8)
    h_out = torch.empty(cap, dtype=torch.uint8)
    h_back = torch.empty(n, dtype=torch.uint8)
    if kind == "pinned":
        h_in, h_out, h_back = h_in.pin_memory(), h_out.pin_memory(), h_back.pin_memory()
    else:
        h_out.zero_(); h_back.zero_()                      # touch the pages
    h_in.copy_(src)
    out_n, got, full = ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_size_t()
    tc, td = [], []
    for i in range(4):
        t0 = time.perf_counter()
        rc = L.b200_deflate_compress_into(h_in.data_ptr(), n, 2, h_out.data_ptr(), cap, ctypes.byref(out_n))
        t1 = time.perf_counter()
        assert rc == 0, rc
        rc = L.b200_inflate(h_out.data_ptr(), out_n.value, h_back.data_ptr(), n, ctypes.byref(got), ctypes.byref(full), 0)
        t2 = time.perf_counter()
        assert rc == 0, rc
        if i:
            tc.append(t1 - t0); td.append(t2 - t1)
    res[kind] = {"compress_GBps": round(n / (sum(tc) / len(tc)) / 1e9, 2), "inflate_GBps": round(n / (sum(td) / len(td)) / 1e9, 2),
                 "ok": bool(got.value == n and torch.equal(h_back, h_in))}
print(json.dumps(res), flush=True)
