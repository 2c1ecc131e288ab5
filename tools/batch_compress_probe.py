"""Development probe: throughput of b200_deflate_compress_batch_dev on many small files (1-64 KiB, log-uniform)
cut from the synthetic corpus, then b200_inflate_batch_dev of the result and a byte-exact comparison.

    gpurun -- 'python tools/batch_compress_probe.py [MiB]'
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import deflate_hpp_b200 as d  # noqa: E402

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n = mib << 20
ctx = d.Context(0)
src = torch.empty(n, dtype=torch.uint8, device="cuda")
ctx.corpus_generate_dev(src.data_ptr(), 20261018, 0, n // d.CHUNK)
rng = np.random.default_rng(7)
lens = []
tot = 0
while tot < n:
    k = int(np.exp(rng.uniform(np.log(1024), np.log(65536))))
    k = min(k, n - tot)
    lens.append(k)
    tot += k
lens = np.array(lens, dtype=np.uint64)
offs = np.concatenate([[0], np.cumsum(lens[:-1])]).astype(np.uint64)
nf = len(lens)
cap = int(sum(int(x) + 20 * ((int(x) + 65535) // 65536) + 16 for x in lens))
t = lambda a: torch.from_numpy(a.view(np.int64)).cuda()
d_off, d_len = t(offs), t(lens)
dst = torch.empty(cap, dtype=torch.uint8, device="cuda")
d_out_off = torch.zeros(nf + 1, dtype=torch.int64, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for level in (2, 3):
    for it in range(3):
        e0.record()
        total = ctx.compress_batch_dev(src.data_ptr(), d_off.data_ptr(), d_len.data_ptr(), nf, level, dst.data_ptr(), cap, d_out_off.data_ptr())
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    # inflate the batch back
    c_off = d_out_off[:-1].contiguous()
    c_len = (d_out_off[1:] - d_out_off[:-1]).contiguous()
    back = torch.zeros(n, dtype=torch.uint8, device="cuda")
    out_len = torch.zeros(nf, dtype=torch.int64, device="cuda")
    status = torch.full((nf,), -1, dtype=torch.int32, device="cuda")
    e0.record()
    ctx.inflate_batch_dev(dst.data_ptr(), c_off.data_ptr(), c_len.data_ptr(), back.data_ptr(), d_off.data_ptr(), d_len.data_ptr(),
                          out_len.data_ptr(), status.data_ptr(), nf)
    e1.record()
    torch.cuda.synchronize()
    ims = e0.elapsed_time(e1)
    ok = bool(int((status != 0).sum()) == 0 and torch.equal(back, src))
    print(f"level {level}: {nf} files, {n / 1e9:.2f} GB -> {total / 1e9:.2f} GB, compress {ms:.2f} ms = {n / ms / 1e6:.1f} GB/s, "
          f"batch inflate {ims:.2f} ms = {n / ims / 1e6:.1f} GB/s, round trip bit-exact: {ok}")
