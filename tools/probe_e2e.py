"""e2e probe of the host-buffer C ABI (pinned and pageable buffers); knobs come from the environment of THIS process
(the per-process default context reads them once).  usage: python tools/probe_e2e.py [MiB]"""
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
for kind in ("pinned", "pageable"):
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
