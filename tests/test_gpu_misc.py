"""GPU tests: device corpus generator == host restatement; the header-only C++ drop-ins run the
reference's own test flow (test/libdeflate.cpp) against the GPU library."""
import os
import subprocess

import pytest

from conftest import GOLD, ROOT

pytestmark = pytest.mark.gpu


def test_device_corpus_matches_host(b200, oracle):
    import torch
    ctx = b200.Context(0)
    for first, n in ((0, 7), (1000003, 5)):
        buf = torch.zeros(n * b200.CHUNK, dtype=torch.uint8, device="cuda")
        ctx.corpus_generate_dev(buf.data_ptr(), 20261018, first, n)
        torch.cuda.synchronize()
        assert bytes(buf.cpu().numpy()) == oracle.corpus(20261018, first, n)


def test_cpp_dropin(tmp_path):
    exe = tmp_path / "test_dropin"
    subprocess.run(["g++", "-std=c++17", "-O1", "-o", str(exe), os.path.join(ROOT, "tests", "cpp", "test_dropin.cpp"),
                    "-lz", "-ldl"], check=True)
    env = dict(os.environ, B200_DEFLATE_LIB=os.path.join(ROOT, "deflate.hpp_b200", "libb200deflate.so"))
    r = subprocess.run([str(exe), GOLD], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr[-4000:]
    assert "[FAIL]" not in r.stderr and "all passed" in r.stderr


def test_host_pipelines_with_small_slices(b200):
    """The pipelined host-buffer calls with slices small enough that a 9 MiB input goes through every branch of the slice
    schedules (compress: full slices, then B/2, B/4, B/4; inflate: graduated slices, control words through the mailbox
    kernels): same bytes as one device call, zlib decodes them, the way back is bit-exact, pinned and pageable."""
    import sys
    script = r"""
import sys, zlib
sys.path.insert(0, %r); sys.path.insert(0, %r)
import torch
import deflate_hpp_b200 as b200
import datagen
data = datagen.text_like(5 << 20, seed=21) + datagen.random_bytes(1 << 20, seed=22) + datagen.image_like((3 << 20) + 12345, seed=23)
n = len(data)
for level in (2, 3, 0):
    c = b200.compress(data, level)                       # pageable memory, host pipeline
    o = zlib.decompressobj(-15)
    assert o.decompress(c) == data and o.eof
    assert b200.decompress(c, out_size=n) == data        # pipelined host inflate (input >= 3 first slices)
    assert b200.decompress(c) == data
    # the same through pinned buffers and against ONE device call
    import ctypes
    L = b200.lib()
    h_in = torch.frombuffer(bytearray(data), dtype=torch.uint8).pin_memory()
    cap = b200.deflate_bound(n)
    h_out = torch.empty(cap, dtype=torch.uint8).pin_memory()
    out_n = ctypes.c_size_t()
    assert L.b200_deflate_compress_into(h_in.data_ptr(), n, level, h_out.data_ptr(), cap, ctypes.byref(out_n)) == 0
    assert bytes(h_out[:out_n.value].numpy()) == c
    ctx = b200.Context(0)
    d_in = h_in.cuda()
    d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
    cn = ctx.compress_dev(d_in.data_ptr(), n, level, d_out.data_ptr(), cap)
    assert bytes(d_out[:cn].cpu().numpy()) == c
    h_back = torch.empty(n, dtype=torch.uint8).pin_memory()
    got, full = ctypes.c_size_t(), ctypes.c_size_t()
    assert L.b200_inflate(h_out.data_ptr(), out_n.value, h_back.data_ptr(), n, ctypes.byref(got), ctypes.byref(full), 0) == 0
    assert got.value == n and bytes(h_back.numpy()) == data
print("ok", b200.launch_count())
""" % (ROOT, os.path.join(ROOT, "tests"))
    env = dict(os.environ, B200_HOST_SLICE_CHUNKS="16", B200_HOST_INFLATE_SLICE=str(1 << 20))
    r = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout[-2000:] + r.stderr[-4000:]


def test_publish_dev(b200):
    """b200_publish_dev: a few 64-bit words from device memory into pinned host memory by a kernel (what the multi-GPU gather
    uses for the sizes of a round instead of a copy that would queue behind the peer copies)."""
    import torch
    ctx = b200.Context(0)
    src = torch.arange(1, 33, dtype=torch.int64, device="cuda") * 0x0101010101
    dst = torch.zeros(32, dtype=torch.int64).pin_memory()
    for n in (1, 8, 32):
        dst.zero_()
        ctx.publish_dev(dst.data_ptr(), src.data_ptr(), n)
        torch.cuda.synchronize()
        assert torch.equal(dst[:n], src[:n].cpu()) and int(dst[n:].abs().sum()) == 0
    with pytest.raises(b200.B200Error):
        ctx.publish_dev(dst.data_ptr(), src.data_ptr(), 33)
    pageable = torch.zeros(4, dtype=torch.int64)
    with pytest.raises(b200.B200Error):                      # not pinned: the device cannot write it
        ctx.publish_dev(pageable.data_ptr(), src.data_ptr(), 4)
