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
import os
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
        self.symm = symm_handle
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
            # one-sided writes: the destination may only read once every writer's copies have landed.  A
            # symmetric-memory barrier is a few flag writes over NVLink, ordered on the current stream behind the
            # copies -- no host synchronisation, no NCCL kernel (B200_GATHER_HOST_BARRIER=1: the round-1 way)
            import os
            if self.symm is not None and hasattr(self.symm, "barrier") and os.environ.get("B200_GATHER_HOST_BARRIER", "0") != "1":
                self.symm.barrier()
            else:
                torch.cuda.current_stream().synchronize()
                dist.barrier()
        total, self.offset = self.offset, 0
        return total


# ---- the product-level multi-GPU calls --------------------------------------------------------------
# BASELINE config 5: one large input partitioned across the GPUs of a box, every GPU's compressed bytes
# gathered to rank 0 for concatenation; and the way back: the joined stream on rank 0 is cut into byte
# ranges, every rank pulls its range and inflates the chunks that start inside it.

CHUNK = 65536
WINDOW_SLACK = CHUNK + 4096          # a chunk that starts just below a range's end reaches at most this far past it
                                     # (stored chunk: 64 KiB + 2 block headers + separator; the segment index only
                                     # goes in front of chunks that save at least 1280 bytes)


BATCH_CHUNKS = 16384          # chunks per batch of one compress call (b200_ctx::batch_chunks)


def round_plan(chunks_per_rank: int, min_round: int = 1024, world: int = 0) -> List[int]:
    """Chunks per round for one rank.  Two costs pull against each other: what a step cannot hide is the copy of the LAST
    round (so it should be small), and every round costs the tail of the persistent matcher, ~0.2 ms in which the SMs run
    dry one after the other (so there should be few).  A third one shows at 8 GPUs: all slices of a round go into ONE GPU,
    whose NVLink ingest (~750 GB/s) takes (world - 1) x 0.08 of the time the round took to compress -- 0.56 at 8 GPUs --
    so a round may not be smaller than that fraction of the round before it, or its copy is still waiting for the link
    when its compression is done (8/16, 5/16, 2/16, 1/16 at 8 GPUs: 1.3 + 0.2 + 0.9 ms exposed, model and measurement
    agree).  With `world` given: geometric rounds, ratio max(0.2, 0.07 x (world - 1)), as many (<= 6) as keep the last one at
    `min_round` chunks (below 64 MiB the matcher starves).  Without: 8/16, 5/16, 2/16, 1/16 of the
    shard; smaller shards get four equal rounds, or one.  (Measured at 8 GPUs, 2 GiB per rank: six rounds 4-4-4-2-1-1 lost
    1.2 ms of 26.8 to tails.)"""
    c = int(chunks_per_rank)
    if c <= 0:
        return []
    if world >= 2 and os.environ.get("B200_ROUND_PLAN", "geo") == "geo":
        rho = max(0.2, 0.07 * (world - 1))
        for k in range(6, 1, -1):
            w = [rho ** i for i in range(k)]
            plan = [int(c * x / sum(w)) for x in w]
            if plan[-1] >= min_round:
                plan[0] += c - sum(plan)
                # a compress call works in batches of 16 384 chunks (the size its scratch is laid out for): a round just
                # above a multiple of that would end in a small batch of its own -- one more matcher tail -- so the
                # excess moves into the next round
                for i in range(k - 1):
                    rem = plan[i] % BATCH_CHUNKS
                    if plan[i] > BATCH_CHUNKS and rem < BATCH_CHUNKS // 4:
                        plan[i] -= rem
                        plan[i + 1] += rem
                return plan
        return [c]
    if c // 16 >= min_round:
        u = c // 16
        plan = [8 * u, 5 * u, 2 * u, u]
        plan[0] += c - 16 * u
        return plan
    if c // 4 >= min_round:
        u = c // 4
        return [u + (c - 4 * u), u, u, u]
    return [c]


def round_layout(plan: List[int], rank: int, world: int) -> List[Tuple[int, int, int]]:
    """For every round k: (global index of the first chunk this rank compresses, chunks, first chunk inside the
    rank's local buffer).  Round k covers global chunks [world * P_k, world * (P_k + S_k)), P_k = sum of the rounds
    before it; rank r takes [world * P_k + r * S_k, + S_k).  Concatenating rounds in order, ranks in order inside a
    round, gives the corpus in order."""
    out, before = [], 0
    for s in plan:
        out.append((world * before + rank * s, s, before))
        before += s
    return out


class ShardedDeflate:
    """deflate::compress / inflate::decompress of ONE stream over the GPUs of one box (one process per GPU,
    torch.distributed; BASELINE config 5).

    compress(): every rank compresses its block-cyclic share round by round (F_NOT_LAST everywhere but on the very
    last slice), the 8-byte sizes of a round travel through an all_gather that is enqueued between two rounds on the
    compute stream, and every rank writes its compressed slice at its final offset of the joined stream in rank
    `dst`'s memory -- NVLink peer copies on the copy engines (symmetric memory) while the next round is being
    compressed, NCCL send/recv where symmetric memory is not available.  The result on `dst` is one valid raw
    DEFLATE stream, byte-identical to what one GPU produces for the whole input when the slices are whole chunks.

    inflate(): the joined stream on `dst` is cut into `world` byte ranges; every rank pulls its range (plus slack for
    the chunk that straddles the range's end) over NVLink and decodes the chunks that START inside the range
    (b200_inflate_shard_dev).  Outputs stay sharded: rank r holds the stream's bytes [out_first, out_first + out_n).

    `codec` needs compress_dev / inflate_shard_dev / deflate_bound with the signatures of api.Context (tests inject a
    CPU stand-in; the product passes the Context of the rank's GPU).
    """

    def __init__(self, codec, device, chunks_per_rank: int, dst: int = 0, plan: List[int] = None,
                 not_last_flag: int = 1, bound=None, transport: str = "auto"):
        self.codec, self.device, self.dst = codec, device, dst
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.plan = list(plan) if plan else round_plan(chunks_per_rank, world=self.world)
        assert sum(self.plan) == chunks_per_rank
        self.layout = round_layout(self.plan, self.rank, self.world)
        self.not_last = not_last_flag
        self.bound = bound
        self.cuda = torch.device(device).type == "cuda"
        rounds = len(self.plan)
        self.dst_cap = [bound(s * CHUNK) for s in self.plan]
        self.dst_off = [sum(self.dst_cap[:k]) for k in range(rounds)]
        self.local = torch.empty(sum(self.dst_cap) + 4096, dtype=torch.uint8, device=device)
        total_cap = bound(chunks_per_rank * CHUNK) * self.world
        self.symm = None
        self.joined_buf = None
        if self.cuda and transport in ("auto", "p2p_copy"):
            self.joined_buf, self.symm = symmetric_buffer(total_cap, device)
        if self.symm is None:
            self.joined_buf = torch.empty(total_cap, dtype=torch.uint8, device=device) if self.rank == dst else None
        self.transport = ("nvlink peer copy (symmetric memory, copy engines)" if self.symm is not None
                          else "send/recv (%s)" % dist.get_backend())
        self.gather = PipelinedGather(self.joined_buf, dst=dst, symm_handle=self.symm)
        self.sizes_dev = torch.zeros(rounds, dtype=torch.int64, device=device)
        self.allsz_dev = torch.zeros(rounds, self.world, dtype=torch.int64, device=device)
        self.allsz_host = torch.zeros(rounds, self.world, dtype=torch.int64)
        if self.cuda:
            self.allsz_host = self.allsz_host.pin_memory()
            self.side = torch.cuda.Stream(device=device)
            self.round_done = [torch.cuda.Event() for _ in range(rounds)]
        self.joined_n = 0

    # -- compress ------------------------------------------------------------------------------------
    def compress(self, src: torch.Tensor, level: int) -> int:
        """src: this rank's slices back to back in round order (round k at chunk offset layout[k][2]).  Returns the size of the
        joined stream (every rank gets it); the bytes are in self.joined_buf on rank `dst`."""
        rounds = len(self.plan)
        last_rank = self.rank == self.world - 1
        if self.cuda:
            main = torch.cuda.current_stream()
            st = main.cuda_stream
        for k, (_, nchunks, first) in enumerate(self.layout):
            flags = 0 if (k == rounds - 1 and last_rank) else self.not_last      # only the stream's very last chunk is final
            out = self.local[self.dst_off[k]:self.dst_off[k] + self.dst_cap[k]]
            if self.cuda:
                self.codec.compress_dev(src.data_ptr() + first * CHUNK, nchunks * CHUNK, level, out.data_ptr(), self.dst_cap[k],
                                        flags=flags, stream=st, d_out_n=self.sizes_dev[k:k + 1].data_ptr(), sync=False)
            else:
                self.sizes_dev[k] = self.codec.compress_dev(src[first * CHUNK:(first + nchunks) * CHUNK], level, out, flags)
            # the sizes of this round: on the compute stream, BETWEEN two rounds (a collective posted on a side stream
            # waits for an SM until the persistent matcher of the next round ends)
            dist.all_gather_into_tensor(self.allsz_dev[k], self.sizes_dev[k:k + 1])
            if self.cuda and self.world <= 32 and hasattr(self.codec, "publish_dev"):
                # to the host by a kernel: a copy of a few bytes would queue on a copy engine behind the peer copies of the
                # round before -- and the compute stream, on which the next round is about to start, with it
                self.codec.publish_dev(self.allsz_host[k].data_ptr(), self.allsz_dev[k].data_ptr(), self.world, stream=st)
            else:
                self.allsz_host[k].copy_(self.allsz_dev[k], non_blocking=True)
            if self.cuda:
                self.round_done[k].record(main)
        # trail the rounds on a side stream: each slice goes to its final offset as soon as its round's sizes are known
        if self.cuda:
            with torch.cuda.stream(self.side):
                for k in range(rounds):
                    self.round_done[k].synchronize()           # host: this round's sizes are in pinned memory
                    self.side.wait_event(self.round_done[k])
                    self.gather.post_round(self.local[self.dst_off[k]:self.dst_off[k] + self.dst_cap[k]], None,
                                           sizes=[int(x) for x in self.allsz_host[k].tolist()])
                self.joined_n = self.gather.finish()
            main.wait_stream(self.side)
        else:
            for k in range(rounds):
                self.gather.post_round(self.local[self.dst_off[k]:self.dst_off[k] + self.dst_cap[k]], None,
                                       sizes=[int(x) for x in self.allsz_host[k].tolist()])
            self.joined_n = self.gather.finish()
        return self.joined_n

    def local_compressed_bytes(self) -> int:
        return int(self.allsz_host[:, self.rank].sum())

    # -- inflate -------------------------------------------------------------------------------------
    def byte_range(self, n: int, rank: int = None) -> Tuple[int, int]:
        r = self.rank if rank is None else rank
        return (n * r) // self.world, (n * (r + 1)) // self.world

    def inflate(self, joined_n: int, out: torch.Tensor, window: torch.Tensor = None, pieces: int = 4):
        """Decode this rank's share of the joined stream (joined_n bytes in rank dst's joined_buf) into `out`.
        Returns (out_n, n_chunks, out_first): this rank produced the stream's bytes [out_first, out_first + out_n).

        With symmetric memory the byte range is pulled in `pieces` parts on a side stream and part k is decoded while
        part k + 1 is still on the wire: all ranks pull from ONE GPU, whose NVLink egress (the whole joined stream but
        its own share, ~10 ms for 16 GiB of input at 8 GPUs) would otherwise sit in front of every rank's decode."""
        lo, hi = self.byte_range(joined_n)
        ws = 0 if self.rank == 0 else max(0, (lo - 16) & ~15)
        we = min(joined_n, hi + WINDOW_SLACK)
        if window is None or window.numel() < we - ws:
            window = torch.empty(max(we - ws, 16), dtype=torch.uint8, device=self.device)
        if self.symm is not None and self.cuda and pieces > 1 and hi - lo >= pieces * (8 << 20):
            out_n, nch = self._inflate_pipelined(joined_n, out, window, lo, hi, ws, pieces)
        else:
            self.pull(window, ws, we, joined_n)
            ends = we == joined_n
            if self.cuda:
                st = torch.cuda.current_stream().cuda_stream
                out_n, nch, _ = self.codec.inflate_shard_dev(window.data_ptr(), we - ws, lo - ws, hi - ws, self.rank == 0, ends,
                                                             out.data_ptr(), out.numel(), stream=st)
            else:
                out_n, nch, _ = self.codec.inflate_shard_dev(window[:we - ws], lo - ws, hi - ws, self.rank == 0, ends, out)
        mine = torch.tensor([out_n, nch], dtype=torch.int64, device=self.device)
        allv = torch.zeros(self.world * 2, dtype=torch.int64, device=self.device)
        dist.all_gather_into_tensor(allv, mine)
        self.inflate_sizes = [int(x) for x in allv[0::2].tolist()]
        out_first = sum(self.inflate_sizes[:self.rank])
        return out_n, nch, out_first

    def _inflate_pipelined(self, joined_n, out, window, lo, hi, ws, pieces):
        main = torch.cuda.current_stream()
        peer = self.symm.get_buffer(self.dst, (self.joined_buf.numel(),), torch.uint8)
        self.symm.barrier()                              # dst's buffer is complete (and nobody is still writing it)
        bounds = [lo + (hi - lo) * i // pieces for i in range(pieces + 1)]
        reach = [min(joined_n, b + WINDOW_SLACK) for b in bounds]            # part k needs the stream up to reach[k + 1]
        start = torch.cuda.Event()
        start.record(main)
        landed = []
        with torch.cuda.stream(self.side):
            self.side.wait_event(start)
            for k in range(pieces):
                a = ws if k == 0 else reach[k]
                b = reach[k + 1]
                if b > a:
                    window[a - ws:b - ws].copy_(peer[a:b], non_blocking=True)
                e = torch.cuda.Event()
                e.record(self.side)
                landed.append(e)
        out_n = nch = 0
        for k in range(pieces):
            main.wait_event(landed[k])
            n_here = reach[k + 1] - ws
            ends = k == pieces - 1 and reach[k + 1] == joined_n
            got, c, _ = self.codec.inflate_shard_dev(window.data_ptr(), n_here, bounds[k] - ws, bounds[k + 1] - ws,
                                                     self.rank == 0 and k == 0, ends, out.data_ptr() + out_n, out.numel() - out_n,
                                                     stream=main.cuda_stream)
            out_n += got
            nch += c
        return out_n, nch

    def pull(self, window: torch.Tensor, ws: int, we: int, joined_n: int):
        """window[0 : we - ws) <- joined stream bytes [ws, we) from rank dst."""
        n = we - ws
        if n <= 0:
            return
        if self.symm is not None:
            peer = self.symm.get_buffer(self.dst, (self.joined_buf.numel(),), torch.uint8)
            self.symm.barrier()                          # dst's buffer is complete (and nobody is still writing it)
            window[:n].copy_(peer[ws:we], non_blocking=True)
            return
        # no peer mapping: dst sends every rank its window
        if self.rank == self.dst:
            ops = []
            for r in range(self.world):
                if r == self.dst:
                    continue
                rlo, rhi = self.byte_range(joined_n, r)
                rws = 0 if r == 0 else max(0, (rlo - 16) & ~15)
                rwe = min(joined_n, rhi + WINDOW_SLACK)
                if rwe > rws:
                    ops.append(dist.P2POp(dist.isend, self.joined_buf[rws:rwe], r))
            window[:n].copy_(self.joined_buf[ws:we])
            if ops:
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
        else:
            for w in dist.batch_isend_irecv([dist.P2POp(dist.irecv, window[:n], self.dst)]):
                w.wait()
