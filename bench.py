#!/usr/bin/env python
"""bench.py -- headline benchmark of the DEFLATE / INFLATE hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--mib M] [--level 2|3]

N = 1 (config.workload = BASELINE.json configs[2], the 1 GiB synthetic mixed-entropy corpus, fast level):
  value / roofline / cpu_baseline / e2e   compress, fast level (K1 lz77 -> K2 huffman -> K3 scan -> K4 encode)
  decompress     b200_inflate_dev of the stream just produced (+ roofline, e2e, cpu_baseline = reference inflater)
  better         level 3 on the same corpus and on the config-2 stand-in (large.bmp), ratio vs the sampled reference
  batch_inflate  config 4: 100 000 independent small streams (zlib / reference producers), one launch
  foreign_inflate  one single zlib-6 stream of the whole corpus (a stream this library did not write) and a
                 reference-level-3 stream
  e2e_dropin     deflate::compress / inflate::decompress through the C++ drop-in headers on pageable memory
  config5        16 GiB at both levels on this one GPU (the N = 1 point of the strong-scaling curve)
N > 1 (torchrun, one rank per GPU): BASELINE config 5 -- 16 GiB in total, STRONG scaling, sharded block-cyclically,
  compressed bytes gathered to rank 0 inside the timed region (deflate.hpp_b200/shard.py: ShardedDeflate), levels 2
  and 3, and the way back (sharded inflate of the joined stream).

Every sub-object carries the metric's own `roofline` (dominant kernel, algorithmic bytes / event-timed launch
duration / measured HBM peak) and `cpu_baseline` (the UNMODIFIED reference, oracle/_ref, on the host cores over a
bounded sample).  `--impl reference` times only that CPU arm and prints the same line shape with "impl": "reference".
"""
import argparse
import ctypes
import json
import os
import statistics
import struct
import subprocess
import sys
import threading
import time
import zlib
from concurrent.futures import ThreadPoolExecutor

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
    ap.add_argument("--mib", type=int, default=1024, help="corpus MiB (N = 1 headline workload)")
    ap.add_argument("--total-gib", type=int, default=16, help="config 5: corpus GiB in total (N > 1, and the N = 1 config5 object)")
    ap.add_argument("--level", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--only", default="", help="comma list of extra sections to run (config1,better,batch,foreign,dropin,config5); default all")
    ap.add_argument("--skip", default="", help="comma list of extra sections to skip")
    ap.add_argument("--ref-budget-s", type=float, default=90.0, help="CPU seconds (wall) the reference arm may use")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu baselines: oracle/_ref (the unmodified reference) -- the only place bench.py
# touches oracle/, and only as the baseline being reported, never on the product path
class RefLib:
    def __init__(self):
        p = os.path.join(ROOT, "oracle", "_ref", "libref_deflate.so")
        if not os.path.exists(p):
            raise FileNotFoundError(p)
        self.lib = L = ctypes.CDLL(p)
        L.ref_quiet(1)
        L.ref_compress_slices_mt.restype = ctypes.c_double
        L.ref_compress_slices_mt.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                             ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        L.ref_inflate_slices_mt.restype = ctypes.c_double
        L.ref_inflate_slices_mt.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p,
                                            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        L.ref_compress.restype = ctypes.c_longlong
        L.ref_compress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t]
        L.ref_inflate.restype = ctypes.c_longlong
        L.ref_inflate.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]

    def compress_mt(self, buf_addr, nbytes, slice_bytes, level, threads):
        import numpy as np
        count = (nbytes + slice_bytes - 1) // slice_bytes
        offs = (np.arange(count, dtype=np.uint64) * slice_bytes)
        lens = np.minimum(slice_bytes, nbytes - offs).astype(np.uint64)
        sizes = np.zeros(count, dtype=np.int64)
        secs = self.lib.ref_compress_slices_mt(buf_addr, offs.ctypes.data, lens.ctypes.data, count, level, threads,
                                               sizes.ctypes.data)
        return secs, int(sizes.sum()), int((sizes < 0).sum())

    def compress(self, data, level):
        """deflate::compress(char*, n, level) of the reference -> bytes (single thread; ctypes drops the GIL)."""
        import numpy as np
        a = np.frombuffer(data, dtype=np.uint8)
        out = np.empty(a.size * 2 + 1024, dtype=np.uint8)
        n = self.lib.ref_compress(a.ctypes.data, a.size, level, out.ctypes.data, out.size)
        if n < 0 or n > out.size:
            raise RuntimeError("reference compress failed")
        return out[:n].tobytes()

    def inflate(self, data, cap):
        import numpy as np
        a = np.frombuffer(data, dtype=np.uint8)
        out = np.empty(cap + 16, dtype=np.uint8)
        n = self.lib.ref_inflate(a.ctypes.data, a.size, out.ctypes.data, cap)
        return None if n < 0 else out[:min(n, cap)].tobytes()

    def inflate_mt(self, streams, out_sizes, threads, repeat=1):
        """inflate::decompress(void*, n, void*, cap) of the reference over independent streams, `threads` host threads.
        Returns (seconds, output bytes, failures)."""
        import numpy as np
        blob = np.frombuffer(b"".join(streams), dtype=np.uint8)
        lens = np.array([len(s) for s in streams], dtype=np.uint64)
        offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint64)
        caps = np.array(out_sizes, dtype=np.uint64)
        ooff = np.concatenate([[0], np.cumsum(caps)[:-1]]).astype(np.uint64)
        total_out = int(caps.sum())
        out = np.empty(total_out * repeat + 64, dtype=np.uint8)
        offs_r = np.tile(offs, repeat)
        lens_r = np.tile(lens, repeat)
        caps_r = np.tile(caps, repeat)
        ooff_r = np.concatenate([ooff + k * total_out for k in range(repeat)]).astype(np.uint64)
        sizes = np.zeros(len(offs_r), dtype=np.int64)
        secs = self.lib.ref_inflate_slices_mt(blob.ctypes.data, offs_r.ctypes.data, lens_r.ctypes.data, len(offs_r), out.ctypes.data,
                                              ooff_r.ctypes.data, caps_r.ctypes.data, threads, sizes.ctypes.data)
        bad = int((sizes != caps_r.astype(np.int64)).sum())
        return secs, total_out * repeat, bad


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


_oracle_lib = None


def host_corpus(nchunks, first_chunk=0):
    """Sample of the workload generated on the host by the oracle's generator (same bytes as the device)."""
    global _oracle_lib
    if _oracle_lib is None:
        so = os.path.join(ROOT, "oracle", "liboracle.so")
        if not os.path.exists(so):
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), so], check=True, capture_output=True)
        _oracle_lib = ctypes.CDLL(so)
        _oracle_lib.oracle_corpus_generate.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_void_p]
    import numpy as np
    buf = np.empty(nchunks * CHUNK, dtype=np.uint8)
    _oracle_lib.oracle_corpus_generate(SEED, first_chunk, nchunks, buf.ctypes.data)
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


def sampled_chunks(total_chunks, count):
    """`count` chunk indices spread evenly over the corpus, kinds (index mod 3) balanced."""
    count = max(3, count - count % 3)
    stride = max(1, total_chunks // count)
    stride += (1 - stride % 3) % 3                          # stride = 1 mod 3: consecutive picks cycle T, I, R
    return [min(total_chunks - 1, i * stride) for i in range(count)]


def cpu_reference_level3(total_chunks, budget_s=10.0):
    """Reference level 3 (deflate.hpp:268-304) is O(n^2) per 32 KB chunk, ~1.1 s each: it is timed on a SAMPLE of 64 KiB
    chunks spread over the corpus, every chunk as its two 32 KB halves (the reference's own
    chunk size), all cores.  Returns the dict and the sampled bytes (so that the GPU can compress the very same bytes)."""
    import numpy as np
    ref = RefLib()
    cores = host_cores()
    per_core = max(1, int(budget_s / 2.4))                  # 2 halves x ~1.1-1.2 s
    picks = sampled_chunks(total_chunks, cores * per_core)
    buf = np.concatenate([host_corpus(1, c) for c in picks])
    secs, comp, bad = ref.compress_mt(buf.ctypes.data, buf.size, 32768, 3, cores)
    return {
        "value": buf.size / secs / 1e9, "unit": UNIT, "cores": cores, "kind": "reference",
        "sample": f"{len(picks)} chunks of 64 KiB spread evenly over the corpus (kinds balanced), as {2 * len(picks)} x 32 KB "
                  f"(the reference's chunk size), level 3, {cores} threads, one pass",
        "ratio": comp / buf.size, "seconds": secs, "bytes": int(buf.size), "failed_slices": bad,
    }, buf


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
    """dram bytes per launch of `kernel` from the committed ncu capture summary (profiles/traffic.json; the capture
    is named there), or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(kernel)
        except Exception:  # noqa: BLE001
            return None
    return None


LZ_NAME = {1: "lz77_literal_kernel", 2: "lz77_fast_kernel", 3: "lz77_better_kernel"}


def roofline(kern, steps, alg_bytes_step, ms_per_step, level=None):
    """Dominant kernel of a profiled region: algorithmic bytes per launch / its average event-timed launch duration
    against the measured HBM peak.  kern: {name: (total ms, launches)} over `steps` steps."""
    if not kern:
        return None
    per_step = {k: (v[0] / steps, max(1, v[1] // steps)) for k, v in kern.items()}
    dom = max(per_step, key=lambda k: per_step[k][0])
    dom_ms, dom_launches = per_step[dom]
    peak, peak_src = measured_peak()
    alg = alg_bytes_step / dom_launches
    avg_ms = dom_ms / dom_launches
    achieved = alg / (avg_ms * 1e-3) / 1e9
    name = LZ_NAME.get(level, dom) if dom == "lz77_kernel" else dom
    return {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "peak_source": peak_src, "traffic": ncu_traffic(name), "algorithmic_bytes_per_launch": alg,
            "avg_launch_ms": avg_ms, "launches_per_step": dom_launches, "kernel_share_of_step": dom_ms / ms_per_step,
            "kernels_ms_per_step": {k: v[0] for k, v in per_step.items()},
            "whole_path_frac": (alg_bytes_step / (ms_per_step * 1e-3) / 1e9) / peak}


def large_bmp_standin():
    """BASELINE config 2: large.bmp is missing from the reference checkout; the stand-in is test.bmp (85 x 85 x 24 bpp,
    pixel data at offset 138) nearest-neighbour upscaled x24 -> 2040 x 2040 x 24 bpp (~12.5 MB), same header layout
    (the same generator as tests/test_gpu_configs.py)."""
    import numpy as np
    src = open(os.path.join(ROOT, "tests", "golden", "test.bmp"), "rb").read()
    off = struct.unpack_from("<I", src, 10)[0]
    w, h = struct.unpack_from("<ii", src, 18)
    stride = (w * 3 + 3) & ~3
    px = np.frombuffer(src, dtype=np.uint8, count=stride * abs(h), offset=off).reshape(abs(h), stride)[:, :w * 3]
    px = px.reshape(abs(h), w, 3)
    k = 24
    big = np.repeat(np.repeat(px, k, axis=0), k, axis=1)
    W, H = w * k, abs(h) * k
    bstride = (W * 3 + 3) & ~3
    rows = np.zeros((H, bstride), dtype=np.uint8)
    rows[:, :W * 3] = big.reshape(H, W * 3)
    hdr = bytearray(src[:off])
    struct.pack_into("<I", hdr, 2, off + rows.size)
    struct.pack_into("<ii", hdr, 18, W, H if h > 0 else -H)
    struct.pack_into("<I", hdr, 34, rows.size)
    return bytes(hdr) + rows.tobytes()


def photo_bmp_standin():
    """BASELINE config 2, second stand-in: a 2 048 x 2 048 x 24 bpp synthetic "photo" (the gradient + noise
    of the corpus's image kind) behind test.bmp's header; the same generator as tests/test_gpu_configs.py."""
    import numpy as np
    src = open(os.path.join(ROOT, "tests", "golden", "test.bmp"), "rb").read()
    off = struct.unpack_from("<I", src, 10)[0]
    W = H = 2048
    x = np.arange(W, dtype=np.uint32)[None, :]
    y = np.arange(H, dtype=np.uint32)[:, None]
    v = ((x >> 2) + (y >> 2)) & 255
    noise = np.random.default_rng(7).integers(0, 3, size=(H, W), dtype=np.uint32)
    px = np.stack([(v + noise) & 255, (2 * v) & 255, 255 - v], axis=2).astype(np.uint8)
    hdr = bytearray(src[:off])
    struct.pack_into("<I", hdr, 2, off + px.size)
    struct.pack_into("<ii", hdr, 18, W, H)
    struct.pack_into("<I", hdr, 34, px.size)
    return bytes(hdr) + px.tobytes()


def gpu_numa_bind(local_rank):
    """Prefer host memory on the NUMA node the rank's GPU hangs off (pinned staging then sits next to the GPU's PCIe
    root) and, where the cpuset allows it, run on that node's cores.  Best effort; returns a description."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        if all(hasattr(pr, a) for a in ("pci_domain_id", "pci_bus_id", "pci_device_id")):
            bus = "%04x:%02x:%02x.0" % (int(pr.pci_domain_id), int(pr.pci_bus_id), int(pr.pci_device_id))
        else:
            bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local_rank)],
                                 capture_output=True, text=True).stdout.strip()
        bus = bus.lower()
        if bus.count(":") == 2 and len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return f"gpu {local_rank}: numa_node unknown"
        cpus = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
        want = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            want |= set(range(int(a), int(b or a) + 1))
        have = os.sched_getaffinity(0)
        both = want & have
        desc = f"gpu {local_rank} -> numa node {node}"
        if both:
            os.sched_setaffinity(0, both)
            desc += f", {len(both)} cpus"
        # memory policy: MPOL_PREFERRED (1) on that node, via the raw syscall (no libnuma in the image)
        libc = ctypes.CDLL(None, use_errno=True)
        mask = ctypes.c_ulong(1 << node)
        SYS_set_mempolicy = 238                                  # x86_64
        r = libc.syscall(SYS_set_mempolicy, 1, ctypes.byref(mask), ctypes.c_ulong(64))
        desc += ", mempolicy preferred" if r == 0 else ", mempolicy unchanged"
        return desc
    except Exception as e:  # noqa: BLE001
        return f"numa bind skipped: {e}"


# ---------------------------------------------------------------------------------------------------
class Bench:
    def __init__(self, args, torch, d, ctx, dev, rank, world, dist):
        self.args, self.torch, self.d, self.ctx, self.dev = args, torch, d, ctx, dev
        self.rank, self.world, self.dist = rank, world, dist
        self.st = torch.cuda.current_stream().cuda_stream
        self.e0 = torch.cuda.Event(enable_timing=True)
        self.e1 = torch.cuda.Event(enable_timing=True)

    def timed(self, fn, steps, warmup, profile=True):
        """warmup untimed calls, then `steps` calls between two CUDA events on the launching stream.  Returns
        (ms per step, {kernel: (ms, launches)}, last result, launches)."""
        torch, ctx = self.torch, self.ctx
        r = None
        for _ in range(warmup):
            r = fn()
        torch.cuda.synchronize()
        if profile:
            ctx.profile(True)
        l0 = self.d.launch_count()
        self.e0.record()
        for _ in range(steps):
            r = fn()
        self.e1.record()
        torch.cuda.synchronize()
        launches = self.d.launch_count() - l0
        kern = {}
        if profile:
            ctx.profile(False)
            kern = ctx.profile_read()
        return self.e0.elapsed_time(self.e1) / steps, kern, r, launches

    # ---- compress + inflate of a device-resident buffer at one level ---------------------------------
    def codec_point(self, src, n, level, steps, warmup, dst=None, back=None, inflate_steps=None):
        torch, d, ctx, st = self.torch, self.d, self.ctx, self.st
        cap = d.deflate_bound(n)
        if dst is None:
            dst = torch.empty(cap + 4096, dtype=torch.uint8, device=self.dev)
        ms, kern, cn, launches = self.timed(lambda: ctx.compress_dev(src.data_ptr(), n, level, dst.data_ptr(), cap, flags=0, stream=st),
                                            steps, warmup)
        out = {"value": n / (ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms, "steps": steps, "warmup": warmup, "level": level,
               "bytes": int(n), "ratio": cn / n, "gpu_launches": int(launches),
               "roofline": roofline(kern, steps, n + cn, ms, level)}
        if back is None:
            back = torch.empty(n, dtype=torch.uint8, device=self.dev)
        isteps = inflate_steps or max(1, min(steps, 10))
        dms, dk, (w, full), _ = self.timed(lambda: ctx.inflate_dev(dst.data_ptr(), cn, back.data_ptr(), n, stream=st), isteps, min(warmup, 3))
        ok = bool(full == n and torch.equal(back[:n], src[:n]))
        out["decompress"] = {"value": n / (dms * 1e-3) / 1e9, "unit": "GB/s (output bytes)", "ms_per_step": dms, "steps": isteps,
                             "round_trip_bit_exact": ok, "roofline": roofline(dk, isteps, cn + n, dms)}
        return out, cn, dst, back


def reference_inflate_baseline(pieces, budget_s=8.0):
    """Reference inflate::decompress (inflate.hpp:338) on all host cores over independent zlib-6 streams of corpus slices."""
    ref = RefLib()
    cores = host_cores()
    with ThreadPoolExecutor(max_workers=cores) as ex:
        def comp(p):
            co = zlib.compressobj(6, zlib.DEFLATED, -15)
            return co.compress(p) + co.flush()
        streams = list(ex.map(comp, pieces))
    sizes = [len(p) for p in pieces]
    secs, outb, bad = ref.inflate_mt(streams, sizes, cores, repeat=1)          # calibrate (also warms the caches)
    rep = int(max(1, min(16, budget_s / max(secs, 1e-3))))
    secs, outb, bad = ref.inflate_mt(streams, sizes, cores, repeat=rep)
    return {"value": outb / secs / 1e9, "unit": "GB/s (output bytes)", "cores": cores, "kind": "reference",
            "sample": f"{len(pieces)} independent zlib-6 streams of {sizes[0] / 2**20:.0f} MiB corpus slices (T/I/R mixed), "
                      f"{cores} threads, decoded {rep}x, reference inflate::decompress(void*, n, void*, cap)",
            "failed_streams": bad, "seconds": secs}


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
            "scaling": "weak" if args.gpus == 1 else "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
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
    numa = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        numa = gpu_numa_bind(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    ctx = d.Context(local_rank)
    B = Bench(args, torch, d, ctx, dev, rank, world, dist)
    if world > 1:
        line = run_multi(B, numa)
        if rank == 0:
            emit(line)
        dist.barrier()
        dist.destroy_process_group()
        return 0
    emit(run_single(B))
    return 0


# ===================================================================================================
def run_single(B):
    args, torch, d, ctx, dev, st = B.args, B.torch, B.d, B.ctx, B.dev, B.st
    only = set(x for x in args.only.split(",") if x)
    skip = set(x for x in args.skip.split(",") if x)

    def want(name):
        return (not only or name in only) and name not in skip

    nchunks = args.mib * 16
    n = nchunks * CHUNK
    src = torch.empty(n, dtype=torch.uint8, device=dev)
    ctx.corpus_generate_dev(src.data_ptr(), SEED, 0, nchunks, stream=st)
    cap = d.deflate_bound(n)
    dst = torch.empty(cap + 4096, dtype=torch.uint8, device=dev)
    back = torch.empty(n, dtype=torch.uint8, device=dev)

    # the reference-level-3 stream of foreign_inflate takes ~10 s of one host core: start it now, in the background
    ref3, ref3_thread = {}, None
    if want("foreign") and not args.no_cpu_baseline:
        def make_ref3():
            try:
                data = host_corpus(4).tobytes()                     # T, I, R, T: 256 KiB
                t0 = time.perf_counter()
                ref3["stream"] = RefLib().compress(data, 3)
                ref3["seconds"] = time.perf_counter() - t0
                ref3["data"] = data
            except Exception as e:  # noqa: BLE001
                ref3["error"] = str(e)
        ref3_thread = threading.Thread(target=make_ref3, daemon=True)
        ref3_thread.start()

    sampler = ClockSampler(0)
    sampler.start()
    # ---- headline: config 3, args.level (fast) --------------------------------------------------------
    head, cn, dst, back = B.codec_point(src, n, args.level, args.steps, args.warmup, dst=dst, back=back)
    clocks = sampler.stop()
    dec = head.pop("decompress")
    launches = head["gpu_launches"]
    ms_per_step = head["ms_per_step"]

    # ---- e2e: host-buffer C-ABI calls, pinned host memory, H2D + D2H inside the timed region -----------
    e2e = None
    L = d.lib()
    if not args.no_e2e:
        h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
        h_in.copy_(src)
        h_out = torch.empty(cap, dtype=torch.uint8).pin_memory()
        out_n = ctypes.c_size_t()
        esteps, ewarm = max(1, min(args.steps, 5)), max(1, min(args.warmup, 2))
        times = []
        for i in range(ewarm + esteps):
            t0 = time.perf_counter()
            rc = L.b200_deflate_compress_into(h_in.data_ptr(), n, args.level, h_out.data_ptr(), cap, ctypes.byref(out_n))
            t1 = time.perf_counter()
            if rc:
                raise d.B200Error(rc, "b200_deflate_compress_into")
            if i >= ewarm:
                times.append(t1 - t0)
        et = sum(times) / len(times)
        e2e = {"value": n / et / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(n), "d2h_bytes_per_step": int(out_n.value),
               "ms_per_step": et * 1e3, "steps": esteps,
               "api": "b200_deflate_compress_into(host in, host out) -- pinned host buffers"}
        h_back = torch.empty(n, dtype=torch.uint8).pin_memory()
        got, full_n = ctypes.c_size_t(), ctypes.c_size_t()
        times = []
        for i in range(3):
            t0 = time.perf_counter()
            rc = L.b200_inflate(h_out.data_ptr(), out_n.value, h_back.data_ptr(), n, ctypes.byref(got), ctypes.byref(full_n), 0)
            t1 = time.perf_counter()
            if rc:
                raise d.B200Error(rc, "b200_inflate")
            if i:
                times.append(t1 - t0)
        et = sum(times) / len(times)
        dec["e2e"] = {"value": n / et / 1e9, "unit": "GB/s (output bytes)", "h2d_bytes_per_step": int(out_n.value), "d2h_bytes_per_step": int(n),
                      "ms_per_step": et * 1e3, "bit_exact": bool(got.value == n and torch.equal(h_back, h_in)),
                      "api": "b200_inflate(host in, host out) -- pinned host buffers"}
        del h_out, h_back

    # ---- cpu baselines of the headline pair -----------------------------------------------------------
    cpu = None
    if not args.no_cpu_baseline:
        try:
            r = cpu_reference_run(args.level, 1, 0, budget_s=12.0)
            cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
            cpu["ratio"] = r["ratio"]
        except Exception as e:  # noqa: BLE001
            cpu = {"value": None, "unit": UNIT, "cores": host_cores(), "kind": "reference", "sample": f"unavailable: {e}"}
        try:
            pieces = [host_corpus(48, 48 * i).tobytes() for i in range(host_cores())]     # 3 MiB slices, kinds mixed
            dec["cpu_baseline"] = reference_inflate_baseline(pieces)
        except Exception as e:  # noqa: BLE001
            dec["cpu_baseline"] = {"value": None, "kind": "reference", "cores": host_cores(), "sample": f"unavailable: {e}"}

    line = {
        "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"{args.mib} MiB of the synthetic mixed-entropy corpus (T/I/R, 64 KiB chunks; BASELINE configs[2]), "
                               f"level {args.level} ({'fast' if args.level == 2 else 'better' if args.level == 3 else args.level}), device-resident",
                   "bytes_per_gpu": int(n), "chunks_per_gpu": nchunks, "l2_hygiene": "inputs (>= 1 GiB) larger than the 126 MB L2",
                   "seed": SEED},
        "ratio": {"b200": head["ratio"], "reference_sample": cpu.get("ratio") if cpu else None},
        "decompress": dec, "roofline": head["roofline"], "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks,
        "gpu_launches": int(launches),
    }

    extras = (("config1", section_config1), ("better", section_better), ("batch_inflate", section_batch),
              ("foreign_inflate", section_foreign), ("e2e_dropin", section_dropin), ("config5", section_config5))
    shared = {"src": src, "n": n, "nchunks": nchunks, "dst": dst, "back": back, "cn": cn, "ref3": ref3,
              "ref3_thread": ref3_thread}
    for name, fn in extras:
        key = name.split("_")[0] if name != "e2e_dropin" else "dropin"
        if not want(key):
            continue
        t0 = time.perf_counter()
        try:
            line[name] = fn(B, shared)
        except Exception as e:  # noqa: BLE001
            import traceback
            traceback.print_exc()
            line[name] = {"error": f"{type(e).__name__}: {e}"}
        if isinstance(line[name], dict):
            line[name]["section_seconds"] = round(time.perf_counter() - t0, 2)
    return line


# ---- config 1: test.bmp through the host API (small-input latency) -------------------------------------------------------
def section_config1(B, S):
    """BASELINE configs[0]: deflate::compress (fast, better) of test.bmp + inflate::decompress round trip -- the reference's
    own CPU-runnable case.  At 21 KB nothing is throughput: what is reported is the latency of one call through the C ABI
    with pageable host buffers (context warm), next to the unmodified reference timed on one host core in this run."""
    d = B.d
    L = d.lib()
    data = open(os.path.join(ROOT, "tests", "golden", "test.bmp"), "rb").read()
    out = {"workload": f"tests/golden/test.bmp ({len(data)} bytes; byte-identical to the reference's fixture), one call at a time through "
                       "b200_deflate_compress_into / b200_inflate, pageable host memory, median of 200 calls"}
    src = ctypes.create_string_buffer(data, len(data))
    cap = d.deflate_bound(len(data))
    dst = ctypes.create_string_buffer(cap)
    back = ctypes.create_string_buffer(len(data) + 64)
    out_n, got, full = ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_size_t()
    try:
        ref = RefLib() if not B.args.no_cpu_baseline else None
    except Exception:  # noqa: BLE001
        ref = None
    for level, name in ((2, "fast"), (3, "better")):
        ts, ti = [], []
        for i in range(220):
            t0 = time.perf_counter()
            rc = L.b200_deflate_compress_into(src, len(data), level, dst, cap, ctypes.byref(out_n))
            t1 = time.perf_counter()
            if rc:
                raise d.B200Error(rc, "b200_deflate_compress_into")
            rc = L.b200_inflate(dst, out_n.value, back, len(data) + 64, ctypes.byref(got), ctypes.byref(full), 0)
            t2 = time.perf_counter()
            if rc:
                raise d.B200Error(rc, "b200_inflate")
            if i >= 20:
                ts.append(t1 - t0); ti.append(t2 - t1)
        ok = got.value == len(data) and back.raw[:len(data)] == data
        zok = zlib.decompressobj(-15).decompress(dst.raw[:out_n.value]) == data
        row = {"compress_us": statistics.median(ts) * 1e6, "inflate_us": statistics.median(ti) * 1e6, "compressed_bytes": int(out_n.value),
               "round_trip_bit_exact": bool(ok), "zlib_decodes_it": bool(zok)}
        if ref is not None:
            reps = 20 if level == 2 else 1
            t0 = time.perf_counter()
            for _ in range(reps):
                rstream = ref.compress(data, level)
            rc_t = (time.perf_counter() - t0) / reps
            t0 = time.perf_counter()
            for _ in range(20):
                rback = ref.inflate(rstream, len(data))
            ri_t = (time.perf_counter() - t0) / 20
            row["cpu_baseline"] = {"kind": "reference", "cores": 1, "compress_us": rc_t * 1e6, "inflate_us": ri_t * 1e6,
                                   "compressed_bytes": len(rstream), "reference_round_trip_ok": bool(rback == data),
                                   "sample": "the same file, deflate::compress(char*, n, level) / inflate::decompress(void*, n, void*, cap)"}
            row["ratio_vs_reference"] = {"b200_over_reference": out_n.value / len(rstream), "tolerance": 1.03}
        out[name] = row
    return out



# ---- better level -------------------------------------------------------------------------------------
def section_better(B, S):
    args, torch, d, ctx, dev, st = B.args, B.torch, B.d, B.ctx, B.dev, B.st
    src, n = S["src"], S["n"]
    steps = max(1, min(args.steps, 5))
    pt, cn3, _, _ = B.codec_point(src, n, 3, steps, 1, dst=S["dst"], back=S["back"], inflate_steps=2)
    out = {"metric": "compress_input_GBps_better_level", "config3": pt}
    if not args.no_cpu_baseline:
        try:
            cb, sample = cpu_reference_level3(S["nchunks"])
            # the same bytes through the GPU compressor (each 64 KiB chunk is independent there too)
            sdev = torch.from_numpy(sample).to(dev)
            sdst = torch.empty(d.deflate_bound(sample.size) + 64, dtype=torch.uint8, device=dev)
            ours = {lv: ctx.compress_dev(sdev.data_ptr(), sample.size, lv, sdst.data_ptr(), sdst.numel() - 64, stream=st) / sample.size
                    for lv in (2, 3)}
            pt["cpu_baseline"] = cb
            pt["ratio_vs_reference"] = {"b200_level3_same_sample": ours[3], "b200_level2_same_sample": ours[2],
                                        "reference_level3_sample": cb["ratio"], "b200_over_reference": ours[3] / cb["ratio"],
                                        "tolerance": 1.03}
        except Exception as e:  # noqa: BLE001
            pt["cpu_baseline"] = {"value": None, "kind": "reference", "sample": f"unavailable: {e}"}
    # config 2: the large.bmp stand-in, both levels
    bmp = large_bmp_standin()
    import numpy as np
    bsrc = torch.from_numpy(np.frombuffer(bmp, dtype=np.uint8).copy()).to(dev)
    c2 = {"workload": f"large.bmp stand-in (test.bmp upscaled x24, {len(bmp)} bytes; BASELINE configs[1])"}
    for lv in (2, 3):
        p, _, _, _ = B.codec_point(bsrc, len(bmp), lv, max(3, steps), 2)
        c2[f"level{lv}"] = p
    if not args.no_cpu_baseline:
        try:
            ref = RefLib()
            a = np.frombuffer(bmp, dtype=np.uint8)
            secs, comp2, _ = ref.compress_mt(a.ctypes.data, a.size, a.size, 2, 1)
            c2["level2"]["cpu_baseline"] = {"value": a.size / secs / 1e9, "unit": UNIT, "cores": 1, "kind": "reference",
                                            "sample": "the whole file, one call of deflate::compress(char*, n, 2)", "ratio": comp2 / a.size}
            c2["level2"]["ratio_vs_reference"] = {"b200": c2["level2"]["ratio"], "reference": comp2 / a.size,
                                                  "b200_over_reference": c2["level2"]["ratio"] / (comp2 / a.size), "tolerance": 1.03}
            picks = [0, 40, 150, 200, 300, 370]               # 32 KB pieces: header, flat and detailed regions
            pieces = np.concatenate([a[p * 32768:(p + 1) * 32768] for p in picks])
            secs, comp3, _ = ref.compress_mt(pieces.ctypes.data, pieces.size, 32768, 3, min(len(picks), host_cores()))
            pdev = torch.from_numpy(pieces.copy()).to(dev)
            pdst = torch.empty(d.deflate_bound(32768) * len(picks) + 64, dtype=torch.uint8, device=dev)
            ours = 0
            for i in range(len(picks)):
                ours += ctx.compress_dev(pdev.data_ptr() + i * 32768, 32768, 3, pdst.data_ptr(), d.deflate_bound(32768), stream=st)
            c2["level3"]["cpu_baseline"] = {"value": pieces.size / secs / 1e9, "unit": UNIT, "cores": min(len(picks), host_cores()),
                                            "kind": "reference", "sample": f"{len(picks)} pieces of 32 KB (indices {picks}), level 3",
                                            "ratio": comp3 / pieces.size}
            c2["level3"]["ratio_vs_reference"] = {"b200_same_pieces": ours / pieces.size, "reference": comp3 / pieces.size,
                                                  "b200_over_reference": ours / comp3, "tolerance": 1.03}
        except Exception as e:  # noqa: BLE001
            c2["cpu_baseline_error"] = str(e)
    out["config2"] = c2
    # config 2, second stand-in: the synthetic photo (device timing, ratio against zlib -6 and, at the fast level, the reference)
    photo = photo_bmp_standin()
    psrc = torch.from_numpy(np.frombuffer(photo, dtype=np.uint8).copy()).to(dev)
    c2p = {"workload": f"synthetic photo, 2048 x 2048 x 24 bpp ({len(photo)} bytes; BASELINE configs[1], stand-in B)",
           "zlib6_ratio": len(zlib.compress(photo, 6)) / len(photo)}
    for lv in (2, 3):
        p, _, _, _ = B.codec_point(psrc, len(photo), lv, max(3, steps), 2)
        c2p[f"level{lv}"] = p
    if not args.no_cpu_baseline:
        try:
            ref = RefLib()
            a = np.frombuffer(photo, dtype=np.uint8)
            secs, comp2, _ = ref.compress_mt(a.ctypes.data, a.size, a.size, 2, 1)
            c2p["level2"]["cpu_baseline"] = {"value": a.size / secs / 1e9, "unit": UNIT, "cores": 1, "kind": "reference",
                                             "sample": "the whole file, one call of deflate::compress(char*, n, 2)", "ratio": comp2 / a.size}
        except Exception as e:  # noqa: BLE001
            c2p["cpu_baseline_error"] = str(e)
    out["config2_photo"] = c2p
    return out


# ---- config 4: batch inflate ----------------------------------------------------------------------------
def section_batch(B, S):
    args, torch, d, ctx, dev, st = B.args, B.torch, B.d, B.ctx, B.dev, B.st
    import numpy as np
    rng = np.random.default_rng(11)
    ndistinct, copies = 2000, 50
    host = S["src"][:96 * CHUNK * 32].cpu().numpy()            # 192 MiB of corpus to cut pieces from
    ref = None
    if not args.no_cpu_baseline:
        try:
            ref = RefLib()
        except Exception:  # noqa: BLE001
            ref = None

    def runs(nbytes, seed):
        r = np.random.default_rng(seed)
        out = bytearray()
        while len(out) < nbytes:
            out += bytes([int(r.integers(0, 256))]) * int(r.integers(1, 700))
        return bytes(out[:nbytes])

    def z(data, level, strategy=zlib.Z_DEFAULT_STRATEGY, flush_every=None):
        co = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
        if flush_every is None:
            return co.compress(data) + co.flush()
        out = b""
        for i in range(0, len(data), flush_every):
            out += co.compress(data[i:i + flush_every]) + co.flush(zlib.Z_SYNC_FLUSH)
        return out + co.flush()
    producers = ["zlib1", "zlib6", "zlib9", "fixed", "stored", "huffman_only", "rle", "sync_flush_8k", "ref_l0", "ref_l3"]
    jobs = []
    for i in range(ndistinct):
        nb = int(np.exp(rng.uniform(np.log(1024), np.log(65536))))
        kind = i % 4
        if kind == 3:
            data = runs(nb, 5000 + i)
        else:
            c = 3 * int(rng.integers(0, host.size // CHUNK // 3 - 1)) + kind      # a chunk of kind T / I / R
            o = int(rng.integers(0, CHUNK - nb + 1))
            data = host[c * CHUNK + o:c * CHUNK + o + nb].tobytes()
        prod = producers[i % len(producers)]
        if prod == "ref_l3":
            data = data[:8192]                                                    # the reference's O(n^2) level
        if prod.startswith("ref_") and ref is None:
            prod = "zlib6"
        jobs.append((prod, data))

    def make(job):
        prod, data = job
        if prod == "zlib1": return z(data, 1)
        if prod == "zlib6": return z(data, 6)
        if prod == "zlib9": return z(data, 9)
        if prod == "fixed": return z(data, 6, zlib.Z_FIXED)
        if prod == "stored": return z(data, 0)
        if prod == "huffman_only": return z(data, 6, zlib.Z_HUFFMAN_ONLY)
        if prod == "rle": return z(data, 6, zlib.Z_RLE)
        if prod == "sync_flush_8k": return z(data, 6, flush_every=8192)
        if prod == "ref_l0": return ref.compress(data, 0)
        return ref.compress(data, 3)
    with ThreadPoolExecutor(max_workers=host_cores()) as ex:
        streams = list(ex.map(make, jobs))
    # golden: zlib's output for zlib's streams; for the reference compressor's own streams what the REFERENCE INFLATER makes
    # of them (the north star's second parity leg).  That is not always the input -- its level 3 mis-encodes runs longer
    # than 258 -- and zlib rejects some of them outright (a repeat code as the very first code length,
    # which the reference reads as "repeat 0", inflate.hpp:170): parity is with what the reference DECODES them to.
    expect, ref_streams_differ, ref_streams_zlib_rejects, ref_streams_dropped = [], 0, 0, 0
    for k, (s_, (prod, data)) in enumerate(zip(streams, jobs)):
        if prod.startswith("ref_"):
            e = ref.inflate(s_, len(data) + 4096)
            if e is None:                                          # the reference cannot read its own stream: not a parity case
                ref_streams_dropped += 1
                streams[k] = z(data, 6)
                e = data
            else:
                try:
                    if zlib.decompressobj(-15).decompress(s_) != e:
                        ref_streams_zlib_rejects += 1
                except zlib.error:
                    ref_streams_zlib_rejects += 1
            ref_streams_differ += e != data
        else:
            e = zlib.decompressobj(-15).decompress(s_)
        expect.append(e)
    for f in ("zlib.dat", "weird.dat"):                           # the reference's own fixtures (zlib framing: skip 2 bytes)
        raw = open(os.path.join(ROOT, "tests", "golden", f), "rb").read()[2:]
        streams.append(raw)
        expect.append(zlib.decompressobj(-15).decompress(raw))
    nd = len(streams)
    order = np.tile(np.arange(nd), copies)[:100000]
    rng.shuffle(order)
    slen = np.array([len(s) for s in streams], dtype=np.uint64)
    olen = np.array([len(e) for e in expect], dtype=np.uint64)
    in_len = slen[order]
    in_off = np.concatenate([[0], np.cumsum((in_len + 15) & ~np.uint64(15))[:-1]]).astype(np.uint64)
    out_cap = olen[order]
    out_off = np.concatenate([[0], np.cumsum((out_cap + 15) & ~np.uint64(15))[:-1]]).astype(np.uint64)
    total_in = int(in_off[-1] + in_len[-1])
    total_out_span = int(out_off[-1] + out_cap[-1])
    blob = np.zeros(total_in + 64, dtype=np.uint8)
    exp = np.zeros(total_out_span + 64, dtype=np.uint8)
    sarr = [np.frombuffer(s, dtype=np.uint8) for s in streams]
    earr = [np.frombuffer(e, dtype=np.uint8) for e in expect]
    for k, j in enumerate(order):
        blob[int(in_off[k]):int(in_off[k]) + len(sarr[j])] = sarr[j]
        exp[int(out_off[k]):int(out_off[k]) + len(earr[j])] = earr[j]
    t = lambda a: torch.from_numpy(a).to(dev)
    d_blob, d_exp = t(blob), t(exp)
    d_in_off, d_in_len, d_out_off, d_out_cap = t(in_off.view(np.int64)), t(in_len.view(np.int64)), t(out_off.view(np.int64)), t(out_cap.view(np.int64))
    d_out = torch.zeros(total_out_span + 64, dtype=torch.uint8, device=dev)
    d_out_len = torch.zeros(len(order), dtype=torch.int64, device=dev)
    d_status = torch.zeros(len(order), dtype=torch.int32, device=dev)
    ns = len(order)

    def step():
        ctx.inflate_batch_dev(d_blob.data_ptr(), d_in_off.data_ptr(), d_in_len.data_ptr(), d_out.data_ptr(), d_out_off.data_ptr(),
                              d_out_cap.data_ptr(), d_out_len.data_ptr(), d_status.data_ptr(), ns, stream=st)
    steps = max(1, min(args.steps, 10))
    ms, kern, _, launches = B.timed(step, steps, 2)
    nout = int(out_cap.sum())
    ncomp = int(in_len.sum())
    ok = bool(int(d_status.abs().sum().item()) == 0 and torch.equal(d_out_len, d_out_cap) and torch.equal(d_out, d_exp))
    out = {"metric": "batch_inflate_output_GBps", "value": nout / (ms * 1e-3) / 1e9, "unit": "GB/s (output bytes)", "ms_per_step": ms,
           "steps": steps, "streams": ns, "distinct_streams": nd, "output_bytes": nout, "compressed_bytes": ncomp,
           "producers": producers + ["zlib.dat", "weird.dat"], "bit_exact_vs_zlib_and_reference_inflater": ok, "gpu_launches": int(launches),
           "reference_streams_not_equal_to_their_input": int(ref_streams_differ), "reference_streams_zlib_disagrees_on": ref_streams_zlib_rejects,
           "reference_streams_the_reference_cannot_read": ref_streams_dropped,
           "config": {"workload": "100 000 independent raw DEFLATE streams, uncompressed size log-uniform in [1 KiB, 64 KiB], content "
                                  "T / I / R / long runs, 10 producers + the reference's fixtures (BASELINE configs[3]); 2 002 distinct "
                                  "streams replicated x50 in shuffled order"},
           "roofline": roofline(kern, steps, ncomp + nout, ms)}
    if ref is not None:
        secs, outb, bad = ref.inflate_mt(streams, [len(e) for e in expect], host_cores(), repeat=1)
        rep = int(max(1, min(64, 6.0 / max(secs, 1e-3))))
        secs, outb, bad = ref.inflate_mt(streams, [len(e) for e in expect], host_cores(), repeat=rep)
        out["cpu_baseline"] = {"value": outb / secs / 1e9, "unit": "GB/s (output bytes)", "cores": host_cores(), "kind": "reference",
                               "sample": f"the {nd} distinct streams decoded {rep}x by the reference inflater, {host_cores()} threads",
                               "failed_streams": bad}
    return out


# ---- foreign streams ------------------------------------------------------------------------------------
def pigz_style_stream(host, level, piece=4 << 20, threads=None):
    """ONE raw DEFLATE stream of `host` made the way pigz does it: pieces compressed in parallel, each primed with the
    32 KiB before it as a preset dictionary (so matches cross piece borders like in any single zlib stream) and closed
    with a sync flush; the last piece ends the stream."""
    npieces = (len(host) + piece - 1) // piece
    mv = memoryview(host)

    def one(i):
        lo = i * piece
        zd = bytes(mv[max(0, lo - 32768):lo])
        co = (zlib.compressobj(level, zlib.DEFLATED, -15, 8, zlib.Z_DEFAULT_STRATEGY, zd) if zd else
              zlib.compressobj(level, zlib.DEFLATED, -15, 8, zlib.Z_DEFAULT_STRATEGY))
        body = co.compress(mv[lo:lo + piece])
        return body + (co.flush() if i == npieces - 1 else co.flush(zlib.Z_SYNC_FLUSH))
    with ThreadPoolExecutor(max_workers=threads or host_cores()) as ex:
        return b"".join(ex.map(one, range(npieces)))


def section_foreign(B, S):
    args, torch, d, ctx, dev, st = B.args, B.torch, B.d, B.ctx, B.dev, B.st
    import numpy as np
    src, n = S["src"], S["n"]
    host = src.cpu().numpy()
    t0 = time.perf_counter()
    stream = pigz_style_stream(host, 6)
    make_s = time.perf_counter() - t0
    comp = torch.from_numpy(np.frombuffer(stream, dtype=np.uint8).copy()).to(dev)
    back = S["back"]
    back.zero_()
    steps = max(1, min(args.steps, 3))
    ms, kern, (w, full), launches = B.timed(lambda: ctx.inflate_dev(comp.data_ptr(), comp.numel(), back.data_ptr(), n, stream=st), steps, 1)
    ok = bool(full == n and torch.equal(back, src))
    out = {"metric": "foreign_stream_inflate_output_GBps",
           "zlib6_single_stream": {"value": n / (ms * 1e-3) / 1e9, "unit": "GB/s (output bytes)", "ms_per_step": ms, "steps": steps,
                                   "output_bytes": int(n), "compressed_bytes": int(comp.numel()), "bit_exact": ok,
                                   "gpu_launches": int(launches),
                                   "config": {"workload": "the whole corpus as ONE raw zlib level-6 stream (pigz-style: 4 MiB pieces primed "
                                                          "with the previous 32 KiB, sync-flushed; blocks start at arbitrary bit offsets and "
                                                          "reference earlier blocks)", "produced_in_s": round(make_s, 2)},
                                   "roofline": roofline(kern, steps, comp.numel() + n, ms)}}
    # the reference inflater beside it: it is single-threaded per stream, so all cores get an independent stream each
    if not args.no_cpu_baseline:
        try:
            pieces = [host[i * (3 << 20):(i + 1) * (3 << 20)].tobytes() for i in range(host_cores())]
            out["zlib6_single_stream"]["cpu_baseline"] = reference_inflate_baseline(pieces, budget_s=6.0)
        except Exception as e:  # noqa: BLE001
            out["zlib6_single_stream"]["cpu_baseline"] = {"value": None, "kind": "reference", "sample": f"unavailable: {e}"}
    # a stream the REFERENCE compressor wrote at level 3 (blocks of 32 KB joined at bit granularity)
    th, ref3 = S.get("ref3_thread"), S.get("ref3")
    if th is not None:
        th.join(timeout=120)
    if ref3 and "stream" in ref3:
        rs = ref3["stream"]
        data = ref3["data"]
        ref_out = RefLib().inflate(rs, len(data) + 1024)
        c = torch.from_numpy(np.frombuffer(rs, dtype=np.uint8).copy()).to(dev)
        o = torch.zeros(len(data) + 1024, dtype=torch.uint8, device=dev)
        ms3, _, (w3, full3), _ = B.timed(lambda: ctx.inflate_dev(c.data_ptr(), c.numel(), o.data_ptr(), o.numel(), stream=st), 3, 1, profile=False)
        got = bytes(o[:w3].cpu().numpy())
        out["reference_level3_stream"] = {"value": len(data) / (ms3 * 1e-3) / 1e9, "unit": "GB/s (output bytes)", "ms_per_step": ms3,
                                          "output_bytes": len(data), "compressed_bytes": len(rs),
                                          "equals_reference_inflater_output": bool(ref_out is not None and got == ref_out),
                                          "equals_input": bool(got == data),
                                          "config": {"workload": "256 KiB of the corpus (T, I, R, T chunks) compressed by the reference's "
                                                                 "deflate::compress(.., 3) in this run", "reference_compress_s": round(ref3.get("seconds", 0), 2)}}
    elif ref3 and "error" in ref3:
        out["reference_level3_stream"] = {"error": ref3["error"]}
    return out


# ---- e2e through the C++ drop-in headers ------------------------------------------------------------------
def section_dropin(B, S):
    args, torch, d = B.args, B.torch, B.d
    exe = "/tmp/b200_bench_dropin"
    r = subprocess.run(["g++", "-std=c++17", "-O2", "-o", exe, os.path.join(ROOT, "tests", "cpp", "bench_dropin.cpp"), "-ldl"],
                       capture_output=True, text=True)
    if r.returncode:
        return {"error": "g++ failed: " + r.stderr[-400:]}
    n = min(S["n"], 1 << 30)
    path = "/dev/shm/b200_dropin_input.bin" if os.path.isdir("/dev/shm") else "/tmp/b200_dropin_input.bin"
    S["src"][:n].cpu().numpy().tofile(path)
    env = dict(os.environ, B200_DEFLATE_LIB=d.lib_path())
    try:
        r = subprocess.run([exe, path, str(args.level), "3"], capture_output=True, text=True, env=env, timeout=600)
    finally:
        os.remove(path)
    try:
        j = json.loads(r.stdout.strip().splitlines()[-1])
    except Exception:  # noqa: BLE001
        return {"error": "harness output: " + (r.stdout + r.stderr)[-400:]}
    if "error" in j:
        return j
    return {"metric": "e2e through include/deflate.hpp + include/inflate.hpp on pageable memory",
            "compress": {"value": j["n"] / j["compress_s"] / 1e9, "best": j["n"] / j["compress_best_s"] / 1e9, "unit": UNIT,
                         "api": "deflate::compress(char*, size_t, int) -> std::vector<uint8_t>",
                         "h2d_bytes_per_step": j["n"], "d2h_bytes_per_step": j["comp"]},
            "decompress": {"value": j["n"] / j["inflate_s"] / 1e9, "best": j["n"] / j["inflate_best_s"] / 1e9, "unit": "GB/s (output bytes)",
                           "api": "inflate::decompress(void*, size_t, void*, size_t)",
                           "h2d_bytes_per_step": j["comp"], "d2h_bytes_per_step": j["n"]},
            "round_trip_bit_exact": j["round_trip"], "bytes": j["n"], "reps": j["reps"], "level": j["level"],
            "harness": "tests/cpp/bench_dropin.cpp (g++ -O2, separate process, wall clock per call, first call not counted)"}


# ---- config 5 on one GPU ---------------------------------------------------------------------------------
def section_config5(B, S):
    args, torch, d, ctx, dev, st = B.args, B.torch, B.d, B.ctx, B.dev, B.st
    for k in ("dst", "back"):
        S[k] = None
    torch.cuda.empty_cache()
    nchunks = args.total_gib * 16384
    n = nchunks * CHUNK
    free, _ = torch.cuda.mem_get_info()
    if free < 3.3 * n:
        return {"skipped": f"needs ~{3.3 * n / 2**30:.0f} GiB of device memory, {free / 2**30:.0f} free"}
    src = torch.empty(n, dtype=torch.uint8, device=dev)
    ctx.corpus_generate_dev(src.data_ptr(), SEED, 0, nchunks, stream=st)
    dst = torch.empty(d.deflate_bound(n) + 4096, dtype=torch.uint8, device=dev)
    back = torch.empty(n, dtype=torch.uint8, device=dev)
    out = {"workload": f"{args.total_gib} GiB of the same corpus on ONE GPU (BASELINE configs[4] at N = 1: the base of the strong-scaling curve)"}
    p2, _, _, _ = B.codec_point(src, n, 2, max(1, min(args.steps, 3)), 1, dst=dst, back=back, inflate_steps=2)
    out["level2"] = p2
    p3, _, _, _ = B.codec_point(src, n, 3, 1, 0, dst=dst, back=back, inflate_steps=1)
    out["level3"] = p3
    return out


# ===================================================================================================
def run_multi(B, numa):
    args, torch, d, ctx, dev, st, dist = B.args, B.torch, B.d, B.ctx, B.dev, B.st, B.dist
    rank, world = B.rank, B.world
    import importlib
    shard = importlib.import_module("deflate_hpp_b200.shard")
    total_chunks = args.total_gib * 16384
    per = total_chunks // world
    total_chunks = per * world
    total_n = total_chunks * CHUNK
    n = per * CHUNK
    sd = shard.ShardedDeflate(ctx, dev, per, bound=d.deflate_bound, not_last_flag=d.F_NOT_LAST,
                              transport=os.environ.get("B200_GATHER", "auto"))
    src = torch.empty(n, dtype=torch.uint8, device=dev)
    for first, nch, local_first in sd.layout:
        ctx.corpus_generate_dev(src.data_ptr() + local_first * CHUNK, SEED, first, nch, stream=st)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, steps, warmup):
        r = None
        for _ in range(warmup):
            r = fn()
        barrier()
        ctx.profile(True)
        l0 = d.launch_count()
        B.e0.record()
        for _ in range(steps):
            r = fn()
        B.e1.record()
        barrier()
        launches = d.launch_count() - l0
        ctx.profile(False)
        return max_over_ranks(B.e0.elapsed_time(B.e1)) / steps, ctx.profile_read(), r, launches

    sampler = ClockSampler(torch.cuda.current_device()) if rank == 0 else None
    if sampler:
        sampler.start()
    ms2, kern2, joined_n, launches = timed(lambda: sd.compress(src, 2), args.steps, args.warmup)
    clocks = sampler.stop() if sampler else None
    local2 = sd.local_compressed_bytes()
    value = total_n / (ms2 * 1e-3) / 1e9

    # ---- the way back: sharded inflate of the joined stream (every rank pulls its byte range from rank 0) ----
    out_cap = int(n * 1.25) + (8 << 20)
    out = torch.empty(out_cap, dtype=torch.uint8, device=dev)
    lo, hi = sd.byte_range(joined_n)
    window = torch.empty(hi - lo + shard.WINDOW_SLACK + 64, dtype=torch.uint8, device=dev)
    isteps = max(1, min(args.steps, 5))
    dms, dkern, (out_n, nch, out_first), _ = timed(lambda: sd.inflate(joined_n, out, window), isteps, 1)
    expect = torch.empty(out_n, dtype=torch.uint8, device=dev)
    ctx.corpus_generate_dev(expect.data_ptr(), SEED, out_first // CHUNK, out_n // CHUNK, stream=st)
    torch.cuda.synchronize()
    okv = torch.tensor([1 if (out_n % CHUNK == 0 and out_first % CHUNK == 0 and torch.equal(out[:out_n], expect)) else 0, out_n],
                       dtype=torch.int64, device=dev)
    dist.all_reduce(okv)
    sharded_ok = bool(int(okv[0].item()) == world and int(okv[1].item()) == total_n)
    del expect
    dec = {"value": total_n / (dms * 1e-3) / 1e9, "unit": "GB/s (output bytes)", "ms_per_step": dms, "steps": isteps,
           "bit_exact_all_ranks": sharded_ok, "outputs": "left sharded (rank r holds the bytes of the chunks that start in its byte range)",
           "includes": "pull of every rank's byte range of the joined stream from rank 0 over NVLink + find_sync + both inflate passes",
           "roofline": roofline(dkern, isteps, (joined_n + total_n) / world, dms)}

    # ---- rank 0: the joined bytes are ONE valid stream of the whole corpus (single-GPU inflate, compared piecewise) ----
    gathered_ok = None
    if rank == 0:
        free, _ = torch.cuda.mem_get_info()
        if free > total_n + (6 << 30):
            whole = torch.empty(total_n, dtype=torch.uint8, device=dev)
            w, full = ctx.inflate_dev(sd.joined_buf.data_ptr(), joined_n, whole.data_ptr(), total_n, stream=st)
            gathered_ok = bool(full == total_n)
            piece = 16384
            exp = torch.empty(piece * CHUNK, dtype=torch.uint8, device=dev)
            for c0 in range(0, total_chunks, piece):
                k = min(piece, total_chunks - c0)
                ctx.corpus_generate_dev(exp.data_ptr(), SEED, c0, k, stream=st)
                gathered_ok = gathered_ok and bool(torch.equal(whole[c0 * CHUNK:(c0 + k) * CHUNK], exp[:k * CHUNK]))
            del whole, exp
    barrier()

    # ---- level 3 (better), same call ----
    s3 = max(1, min(args.steps, 2))
    ms3, kern3, joined3, _ = timed(lambda: sd.compress(src, 3), s3, 1)
    local3 = sd.local_compressed_bytes()
    dms3, _, (out_n3, _, out_first3), _ = timed(lambda: sd.inflate(joined3, out, window), 1, 0)
    expect = torch.empty(out_n3, dtype=torch.uint8, device=dev)
    ctx.corpus_generate_dev(expect.data_ptr(), SEED, out_first3 // CHUNK, out_n3 // CHUNK, stream=st)
    torch.cuda.synchronize()
    ok3 = torch.tensor([1 if torch.equal(out[:out_n3], expect) else 0, out_n3], dtype=torch.int64, device=dev)
    dist.all_reduce(ok3)
    del expect
    better = {"metric": "compress_input_GBps_better_level", "value": total_n / (ms3 * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms3, "steps": s3,
              "ratio": joined3 / total_n, "round_trip_bit_exact_all_ranks": bool(int(ok3[0].item()) == world and int(ok3[1].item()) == total_n),
              "decompress": {"value": total_n / (dms3 * 1e-3) / 1e9, "unit": "GB/s (output bytes)", "ms_per_step": dms3},
              "roofline": roofline(kern3, s3, n + local3, ms3, 3)}

    # ---- e2e: every rank through the host-buffer C-ABI call on its own shard, pinned memory on the GPU's NUMA node ----
    e2e = None
    if not args.no_e2e:
        L = d.lib()
        en = min(n, 1 << 30)
        h_in = torch.empty(en, dtype=torch.uint8).pin_memory()
        h_in.copy_(src[:en])
        ecap = d.deflate_bound(en)
        h_out = torch.empty(ecap, dtype=torch.uint8).pin_memory()
        out_n_ = ctypes.c_size_t()
        times = []
        for i in range(4):
            dist.barrier()
            t0 = time.perf_counter()
            rc = L.b200_deflate_compress_into(h_in.data_ptr(), en, 2, h_out.data_ptr(), ecap, ctypes.byref(out_n_))
            t1 = time.perf_counter()
            if rc:
                raise d.B200Error(rc, "b200_deflate_compress_into")
            if i:
                times.append(t1 - t0)
        et = max_over_ranks(sum(times) / len(times))
        e2e = {"value": en * world / et / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(en), "d2h_bytes_per_step": int(out_n_.value),
               "ms_per_step": et * 1e3, "steps": 3, "numa": numa,
               "api": "b200_deflate_compress_into(host in, host out) on every rank at once, 1 GiB per rank, pinned host buffers "
                      "(no gather: the joined stream of the device-resident path stays in HBM)"}

    if rank != 0:
        return None
    return {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms2, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"{args.total_gib} GiB of the synthetic mixed-entropy corpus in total (BASELINE configs[4]), sharded "
                               f"block-cyclically over {world} GPUs in rounds of {sd.plan} chunks per rank, fast level, device-resident; "
                               f"compressed bytes gathered to rank 0 inside the timed region (ShardedDeflate.compress)",
                   "bytes_total": int(total_n), "bytes_per_gpu": int(n), "rounds": sd.plan, "seed": SEED,
                   "l2_hygiene": "inputs (>= 2 GiB per GPU) larger than the 126 MB L2"},
        "ratio": {"b200": joined_n / total_n},
        "gathered_stream_bit_exact": gathered_ok, "gather_transport": sd.transport,
        "decompress": dec, "better": better,
        "roofline": roofline(kern2, args.steps, n + local2, ms2, 2),
        "cpu_baseline": None, "e2e": e2e, "clocks": clocks, "gpu_launches": int(launches),
    }


if __name__ == "__main__":
    sys.exit(main())
