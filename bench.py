#!/usr/bin/env python
"""bench.py -- headline benchmark of the DEFLATE hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--mib M] [--level 2|3]

Workload (config.workload): BASELINE.json configs[2], the 1 GiB synthetic mixed-entropy corpus
(text-like + image-like + random, 64 KiB chunks, definition in oracle/corpus_oracle.c / DESIGN.md),
fast level, generated on the device.  One step = one pass of the whole compress path over the corpus
(K1 lz77 -> K2 huffman -> K3 scan -> K4 encode, per batch).  With N > 1 (torchrun, one rank per GPU)
every rank owns its own 1 GiB shard of one N GiB corpus (weak scaling), compresses it with
B200_F_NOT_LAST on all but the last rank, and the compressed bytes are gathered to rank 0 over NCCL
inside the timed region -- the only exchange step this path has.

Printed JSON (one line, rank 0):
  value      compress throughput, input GB/s, inputs resident in HBM, CUDA events, max over ranks
  decompress output GB/s of b200_inflate_dev on the stream just produced (same timing rules)
  e2e        the same compress metric through the host-buffer C-ABI call (b200_deflate_compress_into)
             with pinned HOST buffers: H2D + kernels + D2H inside the timed region
  roofline   dominant kernel vs measured HBM peak (MEASURED_PEAKS.json), algorithmic bytes = (1 + r)
             bytes per input byte; per-kernel device time from CUDA events recorded by
             the library around every launch on its stream
  cpu_baseline  the UNMODIFIED reference (oracle/_ref) on the host cores over a bounded sample
`--impl reference` times only that CPU arm and prints the same line shape with "impl": "reference".
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 20261018
CHUNK = 65536
METRIC = "compress_input_GBps_fast_level"
UNIT = "GB/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mib", type=int, default=1024, help="corpus MiB per GPU")
    ap.add_argument("--level", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--ref-budget-s", type=float, default=90.0, help="CPU seconds (wall) the reference arm may use")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: oracle/_ref (the unmodified reference) -- the only place bench.py
# touches oracle/, and only as the baseline being reported, never on the product path
class RefLib:
    def __init__(self):
        p = os.path.join(ROOT, "oracle", "_ref", "libref_deflate.so")
        if not os.path.exists(p):
            raise FileNotFoundError(p)
        self.lib = ctypes.CDLL(p)
        self.lib.ref_quiet(1)
        self.lib.ref_compress_slices_mt.restype = ctypes.c_double
        self.lib.ref_compress_slices_mt.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                                    ctypes.c_int, ctypes.c_int, ctypes.c_void_p]

    def compress_mt(self, buf_addr, nbytes, slice_bytes, level, threads):
        import numpy as np
        count = (nbytes + slice_bytes - 1) // slice_bytes
        offs = (np.arange(count, dtype=np.uint64) * slice_bytes)
        lens = np.minimum(slice_bytes, nbytes - offs).astype(np.uint64)
        sizes = np.zeros(count, dtype=np.int64)
        secs = self.lib.ref_compress_slices_mt(buf_addr, offs.ctypes.data, lens.ctypes.data, count, level, threads,
                                               sizes.ctypes.data)
        return secs, int(sizes.sum()), int((sizes < 0).sum())


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def host_corpus(nchunks, first_chunk=0):
    """Sample of the workload generated on the host by the oracle's generator (same bytes as the device)."""
    so = os.path.join(ROOT, "oracle", "liboracle.so")
    if not os.path.exists(so):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), so], check=True, capture_output=True)
    o = ctypes.CDLL(so)
    o.oracle_corpus_generate.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_void_p]
    import numpy as np
    buf = np.empty(nchunks * CHUNK, dtype=np.uint8)
    o.oracle_corpus_generate(SEED, first_chunk, nchunks, buf.ctypes.data)
    return buf


def cpu_reference_run(level, steps, warmup, budget_s=12.0):
    """Times deflate::compress of the unmodified reference with all host threads on a bounded sample of
    the corpus.  Returns dict(value GB/s, cores, sample, ratio, ms_per_step)."""
    ref = RefLib()
    cores = host_cores()
    # calibrate on 3 chunks (one of each kind), single thread
    cal = host_corpus(3)
    secs, _, _ = ref.compress_mt(cal.ctypes.data, cal.size, cal.size, level, 1)
    per_core = cal.size / max(secs, 1e-6)
    total_steps = max(1, steps + warmup)
    want = per_core * cores * budget_s / total_steps
    nchunks = int(max(3 * cores, min(16384, want // CHUNK)))
    nchunks -= nchunks % 3                                   # whole T/I/R triples
    nchunks = max(nchunks, 3)
    buf = host_corpus(nchunks)
    slice_bytes = 3 * CHUNK                                  # one triple per task
    times = []
    comp = 0
    for i in range(total_steps):
        secs, comp, bad = ref.compress_mt(buf.ctypes.data, buf.size, slice_bytes, level, cores)
        if i >= warmup:
            times.append(secs)
    t = sum(times) / len(times)
    return {
        "value": buf.size / t / 1e9, "unit": UNIT, "cores": cores, "kind": "reference",
        "sample": f"first {nchunks} chunks ({buf.size / 2**20:.0f} MiB) of the same corpus, level {level}, "
                  f"{cores} threads x 192 KiB slices, {len(times)} timed passes",
        "ratio": comp / buf.size, "ms_per_step": t * 1e3, "bytes": int(buf.size),
    }


# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(kernel):
    """dram bytes per launch of `kernel` from the committed ncu capture summary, or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(kernel)
        except Exception:  # noqa: BLE001
            return None
    return None


def main():
    args = parse()
    # stdout must carry exactly one JSON line: park the real stdout and point fd 1 at stderr so that
    # library chatter (e.g. "NCCL version ...") cannot precede it
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        try:
            r = cpu_reference_run(args.level, args.steps, args.warmup, budget_s=args.ref_budget_s)
        except FileNotFoundError as e:
            emit({"impl": "reference", "unavailable": f"oracle/_ref not built: {e}"})
            return 0
        line = {
            "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "1 GiB synthetic mixed-entropy corpus (T/I/R 64 KiB chunks), fast level -- bounded sample",
                       "level": args.level, "sample_bytes": r["bytes"]},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "ratio": {"reference_sample": r["ratio"]},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        emit(line)
        return 0

    import torch
    import deflate_hpp_b200 as d

    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    ctx = d.Context(local_rank)
    st = torch.cuda.current_stream().cuda_stream

    nchunks = args.mib * 16
    n = nchunks * CHUNK
    import importlib
    shard = importlib.import_module("deflate_hpp_b200.shard")
    src = torch.empty(n, dtype=torch.uint8, device=dev)
    cap = d.deflate_bound(n)
    dst = torch.empty(cap + 4096, dtype=torch.uint8, device=dev)
    if world == 1:
        ctx.corpus_generate_dev(src.data_ptr(), SEED, 0, nchunks, stream=st)
        rounds = 1
        slice_chunks = nchunks
    else:
        # block-cyclic shards of one (world x mib) MiB corpus: slice s (slice_chunks chunks) belongs to
        # rank s % world, so compressed slices can be shipped to their final offset round by round
        rounds = int(os.environ.get("B200_GATHER_ROUNDS", "4"))
        if nchunks % rounds or nchunks // rounds < 1024:
            rounds = 1
        slice_chunks = nchunks // rounds
        # B200_GATHER_SPLIT="5,5,4,2": uneven rounds (weights) -- the copy of the LAST round is what the step cannot
        # hide, so it should be small; round k then covers chunks [world * P_k + rank * S_k, + S_k) of the corpus
        round_chunks = [slice_chunks] * rounds
        split = os.environ.get("B200_GATHER_SPLIT", "")
        if split and os.environ.get("B200_GATHER", "p2p_copy") != "fused":
            w = [int(x) for x in split.split(",") if x.strip()]
            if w and all(x > 0 for x in w) and (nchunks % sum(w)) == 0 and min(w) * (nchunks // sum(w)) >= 1024:
                round_chunks = [x * (nchunks // sum(w)) for x in w]
                rounds = len(round_chunks)
        round_first = [sum(round_chunks[:k]) for k in range(rounds)]          # P_k, in chunks (also the offset inside src)
        for k in range(rounds):
            ctx.corpus_generate_dev(src.data_ptr() + round_first[k] * CHUNK, SEED,
                                    world * round_first[k] + rank * round_chunks[k], round_chunks[k], stream=st)
    gather_buf, symm = None, None
    transport = None
    if world > 1:
        gmode = os.environ.get("B200_GATHER", "p2p_copy")
        if gmode in ("p2p_copy", "fused"):
            gather_buf, symm = shard.symmetric_buffer(cap * world, dev)      # rank 0's instance receives
        if symm is None:
            gmode = "nccl"
        transport = {"fused": "fused into the encode kernel: stores to rank 0's memory over NVLink (symmetric memory)",
                     "p2p_copy": "nvlink peer copy (symmetric memory, copy engines)", "nccl": "nccl send/recv"}[gmode]
        if symm is None:
            gather_buf = torch.empty(cap * world, dtype=torch.uint8, device=dev) if rank == 0 else None
    pg = shard.PipelinedGather(gather_buf, dst=0, symm_handle=symm) if world > 1 else None
    if world > 1:
        side = torch.cuda.Stream(device=dev)
        sizes_dev = torch.zeros(rounds, dtype=torch.int64, device=dev)
        round_done = [torch.cuda.Event() for _ in range(rounds)]
        inline_sizes = os.environ.get("B200_GATHER_INLINE_SIZES", "1") != "0"
        allsz_dev = torch.zeros(rounds, world, dtype=torch.int64, device=dev)
        allsz_host = torch.zeros(rounds, world, dtype=torch.int64).pin_memory()
    slice_bytes = slice_chunks * CHUNK
    slice_cap = d.deflate_bound(slice_bytes)
    if world > 1:
        dst_cap = [d.deflate_bound(c * CHUNK) for c in round_chunks]
        dst_off = [sum(dst_cap[:k]) for k in range(rounds)]

    if world > 1 and gmode == "fused":
        peer0 = symm.get_buffer(0, (gather_buf.numel(),), torch.uint8)     # rank 0's buffer as seen from here
        allsz = torch.zeros(rounds, world, dtype=torch.int64, device=dev)
        zero = torch.zeros((), dtype=torch.int64, device=dev)
        token = torch.zeros(1, dtype=torch.int32, device=dev)

    def step_fused():
        # per round: K1..K3, 8-byte all_gather of the shard sizes, K4 writes straight into rank 0's memory
        # at the shard's final offset.  Everything is stream-ordered; the host never waits inside a step.
        running = zero
        keep = []
        for k in range(rounds):
            last = (k == rounds - 1) and (rank == world - 1)
            ctx.compress_stage1_dev(src.data_ptr() + k * slice_bytes, slice_bytes, args.level,
                                    sizes_dev[k:k + 1].data_ptr(), flags=0 if last else d.F_NOT_LAST, stream=st)
            dist.all_gather_into_tensor(allsz[k], sizes_dev[k:k + 1])
            base = running + allsz[k, :rank].sum()
            running = running + allsz[k].sum()
            keep.append(base)
            ctx.compress_stage2_dev(src.data_ptr() + k * slice_bytes, slice_bytes, peer0.data_ptr(),
                                    d_base=base.data_ptr(), stream=st)
        dist.all_reduce(token)          # stream-ordered barrier: every rank's stores have been issued and retired
        step.keep = keep
        step.running = running
        return sizes_dev

    def step():
        if world == 1:
            return ctx.compress_dev(src.data_ptr(), n, args.level, dst.data_ptr(), cap, flags=0, stream=st)
        if gmode == "fused":
            step_fused()
            return None
        # enqueue every round's kernels first (no host sync: sizes stay on the device).  The 8-byte all_gather of
        # a round's sizes is enqueued on the SAME stream, between two rounds: the persistent matcher owns every SM
        # while it runs, and a collective posted on a side stream waits for an SM until that kernel ends
        main = torch.cuda.current_stream()
        for k in range(rounds):
            last = (k == rounds - 1) and (rank == world - 1)       # only the stream's very last chunk is final
            ctx.compress_dev(src.data_ptr() + round_first[k] * CHUNK, round_chunks[k] * CHUNK, args.level,
                             dst.data_ptr() + dst_off[k], dst_cap[k], flags=0 if last else d.F_NOT_LAST, stream=st,
                             d_out_n=sizes_dev[k:k + 1].data_ptr(), sync=False)
            if inline_sizes:
                dist.all_gather_into_tensor(allsz_dev[k], sizes_dev[k:k + 1])
                allsz_host[k].copy_(allsz_dev[k], non_blocking=True)
            round_done[k].record(main)
        # ... and trail them on a side stream: per round the sends/receives (peer copies) at final offsets
        with torch.cuda.stream(side):
            total = 0
            for k in range(rounds):
                if inline_sizes:
                    round_done[k].synchronize()                    # host: this round's sizes are in pinned memory
                    side.wait_event(round_done[k])
                    sz = pg.post_round(dst[dst_off[k]:dst_off[k] + dst_cap[k]], None, sizes=[int(x) for x in allsz_host[k].tolist()])
                else:
                    side.wait_event(round_done[k])
                    sz = pg.post_round(dst[dst_off[k]:dst_off[k] + dst_cap[k]], sizes_dev[k:k + 1])
                total += sz[rank]
            step.joined = pg.finish()
        main.wait_stream(side)
        return total

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    launches0 = None
    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(args.warmup):
        cn = step()
    barrier()
    if sampler:
        sampler.start()
    ctx.profile(True)
    launches0 = d.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        cn = step()
    e1.record()
    barrier()
    launches = d.launch_count() - launches0
    ctx.profile(False)
    kern = ctx.profile_read()
    ms = e0.elapsed_time(e1)
    if world > 1 and gmode == "fused":
        cn = int(sizes_dev.sum().item())
        step.joined = int(step.running.item())
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        tot = torch.tensor([cn], dtype=torch.int64, device=dev)
        dist.all_reduce(tot)
        total_comp = int(tot.item())
    else:
        total_comp = cn
    ms_per_step = ms / args.steps
    total_in = n * world
    value = total_in / (ms_per_step * 1e-3) / 1e9
    ratio = total_comp / total_in

    # ---- decompress (same stream, device resident) ----
    back = torch.empty(n, dtype=torch.uint8, device=dev)
    dec = None
    if world == 1:
        for _ in range(min(args.warmup, 3)):
            ctx.inflate_dev(dst.data_ptr(), cn, back.data_ptr(), n, stream=st)
        torch.cuda.synchronize()
        ctx.profile(True)
        dsteps = max(1, min(args.steps, 10))
        e0.record()
        for _ in range(dsteps):
            w, full = ctx.inflate_dev(dst.data_ptr(), cn, back.data_ptr(), n, stream=st)
        e1.record()
        torch.cuda.synchronize()
        ctx.profile(False)
        dk = ctx.profile_read()
        dms = e0.elapsed_time(e1) / dsteps
        ok = bool(full == n and torch.equal(back, src))
        dec = {"value": n / (dms * 1e-3) / 1e9, "unit": "GB/s (output bytes)", "ms_per_step": dms, "steps": dsteps,
               "round_trip_bit_exact": ok,
               "kernels_ms_per_step": {k: v[0] / dsteps for k, v in dk.items()}}
        if dk:
            # roofline of the inflater's dominant kernel: algorithmic bytes = N_comp + N_out
            ddom = max(dk, key=lambda k: dk[k][0])
            dl = max(1, dk[ddom][1] // dsteps)
            d_alg = float(cn + n) / dl
            d_avg_ms = dk[ddom][0] / dk[ddom][1]
            dpeak, dpeak_src = measured_peak()
            dec["roofline"] = {"bound": "hbm", "kernel": ddom, "achieved": d_alg / (d_avg_ms * 1e-3) / 1e9, "peak": dpeak,
                               "unit": "GB/s", "frac": d_alg / (d_avg_ms * 1e-3) / 1e9 / dpeak, "peak_source": dpeak_src,
                               "traffic": ncu_traffic(ddom), "algorithmic_bytes_per_launch": d_alg,
                               "avg_launch_ms": d_avg_ms, "launches_per_step": dl,
                               "kernel_share_of_step": dk[ddom][0] / dsteps / dms,
                               "whole_path_frac": float(cn + n) / (dms * 1e-3) / 1e9 / dpeak}
    clocks = sampler.stop() if sampler else None

    # ---- N > 1: the bytes gathered on rank 0 must be ONE valid stream of the whole corpus ----
    gathered_ok = None
    if world > 1 and rank == 0:
        total_n = n * world
        joined = step.joined
        whole = torch.empty(total_n, dtype=torch.uint8, device=dev)
        w, full = ctx.inflate_dev(gather_buf.data_ptr(), joined, whole.data_ptr(), total_n, stream=st)
        expect = torch.empty(total_n, dtype=torch.uint8, device=dev)
        ctx.corpus_generate_dev(expect.data_ptr(), SEED, 0, nchunks * world, stream=st)
        torch.cuda.synchronize()
        gathered_ok = bool(full == total_n and torch.equal(whole, expect))
        del whole, expect

    # ---- roofline of the dominant kernel ----
    peak, peak_src = measured_peak()
    per_step = {k: (v[0] / args.steps, v[1] // args.steps) for k, v in kern.items()}
    dom = max(per_step, key=lambda k: per_step[k][0]) if per_step else None
    roof = None
    if dom:
        dom_ms, dom_launches = per_step[dom]
        alg_bytes_step = n * (1.0 + cn / n)                     # (1 + r) bytes per input byte, this rank
        alg_per_launch = alg_bytes_step / max(1, dom_launches)
        avg_launch_ms = dom_ms / max(1, dom_launches)
        achieved = alg_per_launch / (avg_launch_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "peak_source": peak_src,
                "traffic": ncu_traffic({1: "lz77_literal_kernel", 2: "lz77_fast_kernel", 3: "lz77_better_kernel"}.get(args.level, dom)
                                       if dom == "lz77_kernel" else dom),
                "algorithmic_bytes_per_launch": alg_per_launch, "avg_launch_ms": avg_launch_ms,
                "launches_per_step": dom_launches,
                "kernel_share_of_step": dom_ms / ms_per_step,
                "kernels_ms_per_step": {k: v[0] for k, v in per_step.items()},
                "whole_path_frac": (alg_bytes_step / (ms_per_step * 1e-3) / 1e9) / peak}

    # ---- e2e: host-buffer C-ABI call, pinned host memory, H2D + D2H inside the timed region ----
    e2e = None
    if not args.no_e2e:
        L = d.lib()
        h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
        h_in.copy_(src)
        h_out = torch.empty(cap, dtype=torch.uint8).pin_memory()
        out_n = ctypes.c_size_t()
        esteps = max(1, min(args.steps, 5))
        ewarm = max(1, min(args.warmup, 2))
        barrier()
        times = []
        for i in range(ewarm + esteps):
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            rc = L.b200_deflate_compress_into(h_in.data_ptr(), n, args.level, h_out.data_ptr(), cap, ctypes.byref(out_n))
            t1 = time.perf_counter()
            if rc:
                raise d.B200Error(rc, "b200_deflate_compress_into")
            if i >= ewarm:
                times.append(t1 - t0)
        et = sum(times) / len(times)
        if world > 1:
            t = torch.tensor([et], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            et = float(t.item())
        e2e = {"value": total_in / et / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(n), "d2h_bytes_per_step": int(out_n.value),
               "ms_per_step": et * 1e3, "steps": esteps,
               "api": "b200_deflate_compress_into(host in, host out) -- pinned host buffers, per rank"}

    # inflate through the host-buffer C-ABI call as well (not pipelined: H2D of the stream, kernels, D2H of the output)
    if dec is not None and not args.no_e2e:
        L = d.lib()
        h_comp = torch.empty(cn, dtype=torch.uint8).pin_memory()
        h_comp.copy_(dst[:cn])
        h_back = torch.empty(n, dtype=torch.uint8).pin_memory()
        got, full_n = ctypes.c_size_t(), ctypes.c_size_t()
        times = []
        for i in range(3):
            t0 = time.perf_counter()
            rc = L.b200_inflate(h_comp.data_ptr(), cn, h_back.data_ptr(), n, ctypes.byref(got), ctypes.byref(full_n), 0)
            t1 = time.perf_counter()
            if rc:
                raise d.B200Error(rc, "b200_inflate")
            if i:
                times.append(t1 - t0)
        et = sum(times) / len(times)
        dec["e2e"] = {"value": n / et / 1e9, "unit": "GB/s (output bytes)", "h2d_bytes_per_step": int(cn), "d2h_bytes_per_step": int(n),
                      "ms_per_step": et * 1e3, "bit_exact": bool(got.value == n and torch.equal(h_back, src.cpu())),
                      "api": "b200_inflate(host in, host out) -- pinned host buffers"}
        del h_comp, h_back

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            r = cpu_reference_run(args.level, 1, 0, budget_s=15.0)
            cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
            cpu["ratio"] = r["ratio"]
        except Exception as e:  # noqa: BLE001
            cpu = {"value": None, "unit": UNIT, "cores": host_cores(), "kind": "reference", "sample": f"unavailable: {e}"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"{args.mib} MiB per GPU of the synthetic mixed-entropy corpus (T/I/R, 64 KiB chunks), "
                                   f"level {args.level} ({'fast' if args.level == 2 else 'better' if args.level == 3 else args.level}), "
                                   f"device-resident; N>1: per-rank shards + NCCL gather of compressed bytes to rank 0",
                       "bytes_per_gpu": int(n), "chunks_per_gpu": nchunks, "l2_hygiene": "inputs (>= 1 GiB) larger than the 126 MB L2",
                       "seed": SEED},
            "ratio": {"b200": ratio, "reference_sample": cpu.get("ratio") if cpu else None},
            "gathered_stream_bit_exact": gathered_ok, "gather_transport": transport,
            "decompress": dec, "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks,
            "gpu_launches": int(launches),
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
