"""Multi-GPU sharding of the compress path: one process per GPU (torch.distributed), no data-path
collective -- chunks are independent -- and ONE exchange step, the variable-size gather of the
compressed bytes to a destination rank.

    rank r owns the contiguous chunk range shard_range(n_chunks, r, world);
    every rank but the last compresses with F_NOT_LAST, so the rank outputs concatenate, in rank
    order, into one valid raw DEFLATE stream (see include/b200_deflate.h);
    gather_bytes() = all_gather of the byte counts + point-to-point send/recv at the scanned offsets
    (NCCL has no gatherv).

Works with any torch.distributed backend: NCCL on GPUs, gloo on CPU tensors (tests/test_shard_gloo.py).
"""
from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_range(n_chunks: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced chunk range [lo, hi) of `rank`; ranges tile [0, n_chunks) in rank order."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(n_chunks, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_flags(rank: int, world: int, not_last_flag: int) -> int:
    """Compressor flags for this rank: only the last rank's last chunk carries BFINAL."""
    return 0 if rank == world - 1 else not_last_flag


def gather_plan(sizes: List[int]) -> List[int]:
    """Exclusive scan of per-rank byte counts -> offset of each rank's bytes in the joined stream."""
    out, acc = [], 0
    for s in sizes:
        out.append(acc)
        acc += int(s)
    return out


def gather_bytes(local: torch.Tensor, n: int, dst: int = 0, recv_buf: torch.Tensor = None):
    """Gather local[:n] (uint8) from every rank to `dst`, concatenated in rank order.

    Returns (joined_view_or_None, sizes).  On `dst` the result lives in recv_buf (allocated if None)
    and the local shard is copied into place; other ranks return None.
    """
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = local.device
    mine = torch.tensor([n], dtype=torch.int64, device=dev)
    allsz = torch.zeros(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(allsz, mine)
    sizes = [int(x) for x in allsz.tolist()]
    offs = gather_plan(sizes)
    total = offs[-1] + sizes[-1]
    if rank == dst:
        if recv_buf is None or recv_buf.numel() < total:
            recv_buf = torch.empty(total, dtype=torch.uint8, device=dev)
        ops = [dist.P2POp(dist.irecv, recv_buf[offs[r]:offs[r] + sizes[r]], r)
               for r in range(world) if r != dst and sizes[r]]
        recv_buf[offs[dst]:offs[dst] + sizes[dst]].copy_(local[:n])
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return recv_buf[:total], sizes
    if n:
        for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, local[:n], dst)]):
            w.wait()
    return None, sizes


# ---- block-cyclic sharding with a pipelined gather --------------------------------------------------
# A contiguous partition forces the destination to wait for every rank's TOTAL size before it knows
# where rank r's bytes go.  With a block-cyclic partition -- the corpus is cut into slices of
# `slice_chunks` chunks, slice s belongs to rank s % world -- the final offset of a slice depends only
# on slices that precede it in stream order, so after each round (one slice per rank) the ranks
# all_gather that round's sizes and ship the compressed slice straight to its final place on the
# destination while the next round is already being compressed.

def slice_owner(s: int, world: int) -> int:
    return s % world


def slice_first_chunk(round_idx: int, rank: int, world: int, slice_chunks: int) -> int:
    """Global index of the first chunk of the slice `rank` compresses in round `round_idx`."""
    return (round_idx * world + rank) * slice_chunks


def symmetric_buffer(nbytes: int, device):
    """A uint8 buffer of the same size on every rank whose rank-r instance every peer can address
    (torch symmetric memory: CUDA VMM + NVLink peer mapping).  Returns (local tensor, handle) or
    (None, None) when symmetric memory is not usable here (CPU/gloo, older drivers)."""
    try:
        import torch.distributed._symmetric_memory as symm_mem
        buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=device)
        hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
        return buf, hdl
    except Exception:  # noqa: BLE001
        return None, None


class PipelinedGather:
    """Per round: exchange sizes, ship the compressed slice to its final offset, return without waiting.

    Two transports.  With a symmetric-memory handle every sender copies its slice straight into the
    destination's buffer over NVLink with a peer cudaMemcpyAsync: copy engines, no SMs, so it overlaps
    the next round's kernels even while the persistent matcher owns every SM.  Otherwise (gloo, or no
    symmetric memory) NCCL/gloo point-to-point send/recv."""

    def __init__(self, recv_buf: torch.Tensor = None, dst: int = 0, symm_handle=None):
        self.world, self.rank, self.dst = dist.get_world_size(), dist.get_rank(), dst
        self.recv_buf = recv_buf
        self.offset = 0          # bytes of the joined stream placed so far (all rounds before this one)
        self.pending = []
        self.peer_dst = None
        if symm_handle is not None:
            n = recv_buf.numel()
            self.peer_dst = symm_handle.get_buffer(dst, (n,), torch.uint8)   # the destination's buffer, peer-mapped

    def post_round(self, local: torch.Tensor, n, sizes=None):
        """local[:n]: this rank's compressed slice of the current round.  `n` may be an int or a
        one-element int64 tensor on the device (the compressor's d_out_n), in which case the only host
        synchronisation of the round is reading back the gathered sizes.  `sizes`: every rank's byte count of
        this round if the caller has exchanged them already (then no collective is issued here)."""
        dev = local.device
        if sizes is None:
            mine = n if torch.is_tensor(n) else torch.tensor([n], dtype=torch.int64, device=dev)
            allsz = torch.zeros(self.world, dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(allsz, mine)
            sizes = [int(x) for x in allsz.tolist()]
        n = sizes[self.rank]
        offs = [self.offset + o for o in gather_plan(sizes)]
        if self.peer_dst is not None:
            # one-sided: every rank (the destination too) writes its slice at its final offset
            if n:
                self.peer_dst[offs[self.rank]:offs[self.rank] + n].copy_(local[:n], non_blocking=True)
        elif self.rank == self.dst:
            ops = [dist.P2POp(dist.irecv, self.recv_buf[offs[r]:offs[r] + sizes[r]], r)
                   for r in range(self.world) if r != self.dst and sizes[r]]
            self.recv_buf[offs[self.dst]:offs[self.dst] + n].copy_(local[:n], non_blocking=True)
            if ops:
                self.pending += dist.batch_isend_irecv(ops)
        elif n:
            self.pending += dist.batch_isend_irecv([dist.P2POp(dist.isend, local[:n], self.dst)])
        self.offset = offs[-1] + sizes[-1]
        return sizes

    def finish(self) -> int:
        for w in self.pending:
            w.wait()
        self.pending = []
        if self.peer_dst is not None:
            # one-sided writes: the destination may only read once every writer's copies have landed
            torch.cuda.current_stream().synchronize()
            dist.barrier()
        total, self.offset = self.offset, 0
        return total
