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
