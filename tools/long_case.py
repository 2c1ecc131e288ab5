"""one damaged long stream (case k of test_foreign_errors_match_sequential) through b200.decompress; the caller kills us on a hang"""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import deflate_hpp_b200 as b200  # noqa: E402
import test_gpu_foreign as G  # noqa: E402

k = int(sys.argv[1])
data = G.mixed(3_000_000, seed=13)
stream = bytearray(G.raw(data, 6))
rng = np.random.default_rng(5)
cases = [bytes(stream[:len(stream) // 2]), bytes(stream[:len(stream) - 3])]
for _ in range(6):
    s = bytearray(stream)
    p = int(rng.integers(1000, len(s) - 1000))
    s[p] ^= 1 << int(rng.integers(0, 8))
    cases.append(bytes(s))
s = cases[k]
print("case", k, "len", len(s), "launches before", b200.launch_count(), flush=True)
try:
    out = b200.decompress(s, out_size=len(data) + 4096)
    print("ok", len(out), out == data, flush=True)
except b200.B200Error as e:
    print("err", e.code, flush=True)
print("launches", b200.launch_count(), flush=True)
