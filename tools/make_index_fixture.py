"""Generates tests/golden/b200_indexed.deflate: a GPU-compressed stream (fast level) of three full 64 KiB chunks
plus a partial one, whose full chunks carry the segment index, and b200_indexed_short.deflate: one short chunk of ten
segments with an index of its own size.  tests/test_oracle.py::test_segment_index_fixture
checks on the CPU that the index words equal the bit lengths an independent decoder observes.

    gpurun -- 'python tools/make_index_fixture.py'    (writes gpurun_out/b200_indexed.deflate; copy it to tests/golden/)
"""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import datagen  # noqa: E402
import deflate_hpp_b200 as d  # noqa: E402

data = datagen.text_like(65536, seed=51) + datagen.image_like(65536, seed=52) + datagen.text_like(65536 + 3000, seed=53)
c = d.compress(data, d.LEVEL_FAST)
assert d.decompress(c) == data
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "b200_indexed.deflate"), "wb").write(c)
print(len(data), "->", len(c))
# a stream that is ONE short chunk of 10 segments: its index has 40 groups (tests/golden/b200_indexed_short.deflate,
# test_segment_index_short_fixture)
short = datagen.text_like(40000, seed=61)
cs = d.compress(short, d.LEVEL_FAST)
assert d.decompress(cs) == short
open(os.path.join(ROOT, "gpurun_out", "b200_indexed_short.deflate"), "wb").write(cs)
print(len(short), "->", len(cs))
