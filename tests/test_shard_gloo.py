"""CPU tests of the multi-GPU host logic (gloo, world_size 2 and 3): chunk partition, flags and the
variable-size gather used by bench.py --gpus N.  The byte payloads are real raw-deflate shards made
with zlib full-flush framing, so the joined result can be checked as ONE valid stream."""
import os
import sys
import zlib

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))


def _shard_mod():
    import importlib.util
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    spec = importlib.util.spec_from_file_location("b200_shard", os.path.join(root, "deflate.hpp_b200", "shard.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_shard_range_tiles():
    sh = _shard_mod()
    for n in (0, 1, 7, 16, 16384, 262144 + 5):
        for world in (1, 2, 3, 4, 8):
            prev = 0
            sizes = []
            for r in range(world):
                lo, hi = sh.shard_range(n, r, world)
                assert lo == prev and hi >= lo
                prev = hi
                sizes.append(hi - lo)
            assert prev == n and max(sizes) - min(sizes) <= 1
    assert sh.gather_plan([5, 0, 7]) == [0, 5, 5]
    assert [sh.shard_flags(r, 3, 1) for r in range(3)] == [1, 1, 0]


def _payload(rank, world):
    """Rank's shard of a 3-chunk-per-rank message, as a raw deflate fragment (non-final except last)."""
    data = bytes([65 + rank]) * (1000 * (rank + 1)) + os.urandom(0)
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    frag = co.compress(data) + (co.flush(zlib.Z_FULL_FLUSH) if rank != world - 1 else co.flush())
    return data, frag


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sh = _shard_mod()
        data, frag = _payload(rank, world)
        local = torch.frombuffer(bytearray(frag) + bytearray(64), dtype=torch.uint8)   # capacity > n on purpose
        joined, sizes = sh.gather_bytes(local, len(frag), dst=0)
        if rank == 0:
            q.put((bytes(joined.numpy()), sizes))
        else:
            assert joined is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gather_bytes_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000 + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    joined, sizes = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    frags = [_payload(r, world) for r in range(world)]
    assert sizes == [len(f[1]) for f in frags]
    assert joined == b"".join(f[1] for f in frags)
    # the rank-ordered concatenation is one valid raw deflate stream of the rank-ordered data
    o = zlib.decompressobj(-15)
    assert o.decompress(joined) == b"".join(f[0] for f in frags) and o.eof


def _pipelined_worker(rank, world, port, q, pre_exchanged):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sh = _shard_mod()
        rounds = 3
        recv = torch.zeros(1 << 16, dtype=torch.uint8) if rank == 0 else None
        pg = sh.PipelinedGather(recv, dst=0)
        for k in range(rounds):
            frag = bytes([97 + rank + 3 * k]) * (50 * (rank + 1) + 7 * k)
            local = torch.frombuffer(bytearray(frag) + bytearray(16), dtype=torch.uint8)
            if pre_exchanged:
                # what bench.py does: the sizes travel through an all_gather the caller issued itself
                mine = torch.tensor([len(frag)], dtype=torch.int64)
                allsz = torch.zeros(world, dtype=torch.int64)
                dist.all_gather_into_tensor(allsz, mine)
                pg.post_round(local, None, sizes=[int(x) for x in allsz.tolist()])
            else:
                pg.post_round(local, len(frag))
        total = pg.finish()
        if rank == 0:
            q.put(bytes(recv[:total].numpy()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("pre_exchanged", [False, True])
def test_pipelined_gather_gloo(pre_exchanged):
    """Block-cyclic rounds: after every round each rank's slice lands at its final offset on rank 0, round by
    round and rank by rank, with the sizes exchanged by the gatherer or by the caller."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000 + int(pre_exchanged)
    procs = [ctx.Process(target=_pipelined_worker, args=(r, world, port, q, pre_exchanged)) for r in range(world)]
    for p in procs:
        p.start()
    joined = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    expect = b"".join(bytes([97 + r + 3 * k]) * (50 * (r + 1) + 7 * k) for k in range(3) for r in range(world))
    assert joined == expect


# ---- ShardedDeflate (the product-level multi-GPU call) with a CPU stand-in codec ---------------------
SEP_TAIL = b"\x00\x00\xff\xff\x00\x00\x00\xff\xff"


class _ZlibChunkCodec:
    """Same stream STRUCTURE as the GPU compressor (independent 64 KiB chunks, each followed by the doubled empty
    stored block, BFINAL on the last), produced with zlib so that the host logic can be tested without a GPU."""
    CH = 65536

    def deflate_bound(self, n):
        return n + 64 * ((n + self.CH - 1) // self.CH) + 64

    def compress_dev(self, src, level, out, flags):
        data = bytes(src.numpy())
        parts = []
        nch = (len(data) + self.CH - 1) // self.CH
        for i in range(nch):
            co = zlib.compressobj(6, zlib.DEFLATED, -15)
            last = i == nch - 1 and not (flags & 1)
            body = co.compress(data[i * self.CH:(i + 1) * self.CH])
            parts.append(body + (co.flush() if last else co.flush(zlib.Z_FULL_FLUSH) + b"\x00\x00\x00\xff\xff"))
        blob = b"".join(parts)
        out[:len(blob)] = torch.frombuffer(bytearray(blob), dtype=torch.uint8)
        return len(blob)

    def inflate_shard_dev(self, window, lo, hi, first_is_start, ends_stream, out):
        w = bytes(window.numpy())
        starts = [0] if first_is_start else []
        p = w.find(SEP_TAIL)
        while p >= 0:
            s = p + len(SEP_TAIL)
            if s < len(w) and (s >= lo or first_is_start) and s > 0:
                starts.append(s)
            p = w.find(SEP_TAIL, p + 1)
        starts = sorted(set(starts))
        mine = [s for s in starts if s < hi]
        bounds = starts + [len(w)]
        pos = 0
        for s in mine:
            e = bounds[bounds.index(s) + 1]
            piece = zlib.decompressobj(-15).decompress(w[s:e])
            out[pos:pos + len(piece)] = torch.frombuffer(bytearray(piece), dtype=torch.uint8)
            pos += len(piece)
        nxt = bounds[len(mine)] if len(mine) < len(bounds) else len(w)
        return pos, len(mine), nxt


def _corpus_chunk(i):
    import numpy as np
    rng = np.random.default_rng(1000 + i)
    if i % 3 == 2:
        return rng.integers(0, 256, 65536, dtype=np.uint8).tobytes()
    return (bytes([97 + i % 26]) * 61 + bytes(rng.integers(0, 256, 3, dtype=np.uint8))) * 1024


def _sharded_worker(rank, world, port, q, chunks_per_rank, plan):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sh = _shard_mod()
        codec = _ZlibChunkCodec()
        sd = sh.ShardedDeflate(codec, "cpu", chunks_per_rank, plan=plan, bound=codec.deflate_bound)
        src = bytearray()
        for first, nch, local_first in sd.layout:
            assert local_first * 65536 == len(src)
            for cidx in range(first, first + nch):
                src += _corpus_chunk(cidx)
        n = sd.compress(torch.frombuffer(src, dtype=torch.uint8), 2)
        out = torch.zeros(chunks_per_rank * world * 65536, dtype=torch.uint8)
        out_n, nch, out_first = sd.inflate(n, out)
        q.put((rank, bytes(sd.joined_buf[:n].numpy()) if rank == 0 else None, out_n, nch, out_first, bytes(out[:out_n].numpy())))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,chunks_per_rank,plan", [(2, 4, [2, 1, 1]), (3, 3, [1, 1, 1]), (2, 5, None)])
def test_sharded_deflate_gloo(world, chunks_per_rank, plan):
    """compress: block-cyclic rounds land in order on rank 0 as ONE valid stream; inflate: byte-range windows, each
    rank decodes exactly the chunks that start in its range, together they cover the stream once, in order."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + os.getpid() % 2000 + world * 7 + chunks_per_rank
    procs = [ctx.Process(target=_sharded_worker, args=(r, world, port, q, chunks_per_rank, plan)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    total = b"".join(_corpus_chunk(i) for i in range(world * chunks_per_rank))
    joined = got[0][1]
    o = zlib.decompressobj(-15)
    assert o.decompress(joined) == total and o.eof
    pos = 0
    for rank, _, out_n, nch, out_first, data in got:
        assert out_first == pos and out_n == nch * 65536
        assert data == total[pos:pos + out_n]
        pos += out_n
    assert pos == len(total)


def test_round_plan():
    sh = _shard_mod()
    for c in (1, 5, 1024, 4096, 4097, 16384, 32768, 32771, 131072):
        plan = sh.round_plan(c)
        assert sum(plan) == c and all(x > 0 for x in plan)
        if c >= 16384:
            assert plan[-1] * 16 <= c and plan[-1] >= 1024
        for world in (1, 2, 8):
            seen = []
            for r in range(world):
                seen += [(f, f + s) for f, s, _ in sh.round_layout(plan, r, world)]
            seen.sort()
            assert seen[0][0] == 0 and seen[-1][1] == world * c
            assert all(a[1] == b[0] for a, b in zip(seen, seen[1:]))


def test_round_plan_by_world(monkeypatch):
    """round_plan(world): the sizes add up, every round keeps the matcher busy (>= 1 024 chunks), rounds shrink, and none
    shrinks faster than the incast into rank 0 allows -- a round's slices take (world - 1) x 0.07 of the time the round took
    to compress to arrive, so the next round may not be smaller than that fraction (minus the remainder that moved to avoid
    a small trailing batch); no round sits just above a multiple of the 16 384-chunk compress batch; the layouts of all
    ranks tile the corpus.  B200_ROUND_PLAN=fixed gives the 8/16, 5/16, 2/16, 1/16 plan."""
    sh = _shard_mod()
    for world in (2, 3, 4, 8):
        for total_gib in (1, 4, 16):
            c = total_gib * 16384 // world
            plan = sh.round_plan(c, world=world)
            assert sum(plan) == c and all(x > 0 for x in plan) and len(plan) <= 6
            if len(plan) > 1:
                assert plan[-1] >= 1024
                assert all(a >= b for a, b in zip(plan, plan[1:]))
                rho = max(0.2, 0.07 * (world - 1))
                for a, b in zip(plan, plan[1:]):
                    assert b >= rho * a * 0.75 - 1, (world, plan)
                for x in plan[:-1]:
                    assert not (x > sh.BATCH_CHUNKS and 0 < x % sh.BATCH_CHUNKS < sh.BATCH_CHUNKS // 4), (world, plan)
            seen = []
            for r in range(world):
                seen += [(f, f + s) for f, s, _ in sh.round_layout(plan, r, world)]
            seen.sort()
            assert seen[0][0] == 0 and seen[-1][1] == world * c
            assert all(a[1] == b[0] for a, b in zip(seen, seen[1:]))
    assert sh.round_plan(32768, world=8) == [16384, 10041, 4257, 2086]
    assert sh.round_plan(500, world=8) == [500]
    monkeypatch.setenv("B200_ROUND_PLAN", "fixed")
    assert sh.round_plan(32768, world=8) == [16384, 10240, 4096, 2048]
