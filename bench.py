#!/usr/bin/env python
"""bench.py -- WAH compress/decompress throughput on B200 (the metric of BASELINE.json).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload NAME] [--density D] [--mode M]

One "step" = one pass of the hot path over one batch of synthetic input: compress it, then decompress the result.

N = 1   workload `clustered_16gbit` = BASELINE.json configs[2], the largest single-GPU configuration: one 16 Gbit
        run-clustered bitvector (2 GiB, far larger than the 126 MB L2).  `value` is quoted at density 0.01 in the
        reference-exact BLOCK1024 mode; `sweep` holds both kernels at every density of the config's sweep
        (1e-4 ... 0.5) in both encoder modes, and `roofline` is the WORST kernel of that sweep.  Extra keys:
        `sparse_1gbit` (configs[1]), `dense_1gbit`, and `bitmap_index` (configs[3] on one GPU: the N = 1 point of the
        curve the N > 1 runs continue).
N > 1   one process per GPU.  `value` = configs[3], the bitmap index of 1024 columns x 64 Mbit, the columns split
        across the ranks (strong scaling; no collective touches the data, the all-gather of the column lengths is
        inside the timed region).  `range_128gbit` = configs[4]: one 128 Gbit vector range-sharded over the ranks:
        local compress, NCCL all-gather of the shard records, seam plan, NCCL all-gather-v of the segments, local
        decompress -- all timed.  Both are checked before they are timed (`parity_checked`): every rank's shard
        against the CPU oracle on a prefix, the column batch against a column-by-column compress, the gathered
        stream bit for bit against a single-GPU compress of the whole vector, in both modes.

`value` counts uncompressed bytes through the path per second over all ranks (4n in on compress + 4n out on
decompress), inputs resident in HBM.  `e2e` is the same through the reference-facing host entry points
(wah_compress_host / wah_decompress_host = the reference's compress() / decompress(): a pageable, malloc()ed input
like the reference's callers pass, source.cpp:75,97-100; malloc()ed results), host <-> device copies inside the
timed region.  `--impl reference` times the CPU oracle port (the reference has no CPU implementation) with all host
threads on the same workload at full size.  One JSON line on stdout, from rank 0.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "WAH compress/decompress GB/s (uncompressed bytes)"
UNIT = "GB/s"
SWEEP = (0.0001, 0.001, 0.01, 0.1, 0.25, 0.5)      # BASELINE.json configs[2]: "density sweep 0.0001-0.5"
MEAN_RUN = 1000.0                                  # "Markov runs, mean length 1000 bits"
N_COLS, COL_WORDS = 1024, 1 << 21                  # configs[3]: 1024 columns x 64 Mbit
RANGE_WORDS = 1 << 32                              # configs[4]: one 128 Gbit vector

WORKLOADS = {
    # name: (words, generator, density, description)
    "clustered_16gbit": (1 << 29, "clustered", 0.01, "16 Gbit run-clustered bitvector (Markov runs, mean 1000 bits), BASELINE configs[2]"),
    "sparse_1gbit": (1 << 25, "uniform", 0.001, "1 Gbit sparse bitvector, density 0.001 (BASELINE configs[1])"),
    "uniform_32mbit": (1 << 20, "uniform", 0.5, "32 Mbit uniform random, density 0.5 (BASELINE configs[0])"),
    "dense_1gbit": (1 << 25, "uniform", 0.5, "1 Gbit uniform random, density 0.5"),
}
BITMAP_DESC = "bitmap index of 1024 columns x 64 Mbit (run clustered, density 0.01), columns sharded across the GPUs (BASELINE configs[3])"
RANGE_DESC = "one 128 Gbit (16 GB) run-clustered bitvector, density 0.01, range-sharded over the GPUs with boundary fill merge and NCCL all-gather of the compressed segments (BASELINE configs[4])"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="clustered_16gbit", choices=list(WORKLOADS), help="N = 1 only")
    ap.add_argument("--density", type=float, default=None)
    ap.add_argument("--mode", default="block1024", choices=["block1024", "canonical"])
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-range", action="store_true", help="N > 1: skip the range-sharded 128 Gbit vector")
    ap.add_argument("--cols", type=int, default=N_COLS, help="columns of the bitmap index (development only)")
    ap.add_argument("--range-log2", type=int, default=32, help="log2 words of the range-sharded vector (development only)")
    return ap.parse_args()


def host_threads():
    # all host cores, whatever OMP_NUM_THREADS says (torchrun sets it to 1 for every rank)
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def oracle():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as orc

    return orc


# ----------------------------------------------------------------------------- CPU arm (--impl reference)


def run_reference(args):
    """The reference has no CPU implementation: the oracle port (oracle/wah_oracle.c, OpenMP) on all host cores, on
    the very workload the GPU arm reports as `value`, at full size."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np

    orc = oracle()
    threads = host_threads()
    mode = 0 if args.mode == "block1024" else 1
    if args.gpus == 1:
        n_words, gen, density, desc = WORKLOADS[args.workload]
        if args.density is not None:
            density = args.density
        data = np.empty(n_words, dtype=np.uint32)
        if gen == "clustered":
            orc.gen_clustered(n_words, density, MEAN_RUN, 1337, out=data)
        else:
            orc.gen_uniform(n_words, density, 1337, out=data)

        def step():
            cw = orc.compress(data, mode, threads=threads)
            return orc.decompress(cw, threads=threads)

        back = step()
        assert back[: data.size].tobytes() == data.tobytes()
        del back
        nbytes = 4.0 * n_words
        config = {"workload": args.workload, "description": desc, "words": n_words, "density": density, "mode": args.mode}
        sample = f"the whole {n_words}-word vector per step (no sampling)"
    else:
        n_cols = args.cols
        data = np.empty(n_cols * COL_WORDS, dtype=np.uint32)
        orc.gen_clustered(data.size, 0.01, MEAN_RUN, 1337, out=data)
        cols = data.reshape(n_cols, COL_WORDS)

        def step():
            return orc.roundtrip_columns(cols, mode, threads, verify=False)

        _, bad = orc.roundtrip_columns(cols[: min(n_cols, 2 * threads)], mode, threads, verify=True)
        assert bad == 0
        nbytes = 4.0 * data.size
        config = {"workload": "bitmap_index", "description": BITMAP_DESC, "columns": n_cols, "words_per_column": COL_WORDS,
                  "density": 0.01, "mode": args.mode}
        sample = f"all {n_cols} columns per step (no sampling), columns dealt to the threads"
    for _ in range(max(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = 2 * nbytes * args.steps / dt / 1e9
    config["note"] = "the reference has no CPU implementation; this is the CPU oracle port (oracle/wah_oracle.c, OpenMP) on the host cores"
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak" if args.gpus == 1 else "strong", "vs_baseline": None, "dtype": "u32",
        "data": "synthetic", "config": config,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


# ----------------------------------------------------------------------------- helpers of the GPU arm


class ClockSampler:
    """Polls NVML from a thread DURING the timed region: SM clock and clock-event (throttle) reasons."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        import threading

        self.samples, self.bits, self.max_mhz, self.h = [], 0, None, None
        self._stop = threading.Event()
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.h = None
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        if self.h is None:
            return
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.bits |= nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            except Exception:
                break
            time.sleep(0.001)

    def stop(self):
        self._stop.set()
        self.t.join(timeout=2)
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": len(self.samples)}
        if self.samples:
            s = sorted(self.samples)
            out["sm_mhz"] = s[len(s) // 2]
            out["reasons"] = sorted(name for bit, name in self.REASONS.items() if self.bits & bit)
        return out


def load_peak():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except (OSError, KeyError, ValueError):
        return 6650.0, "fallback 6650 GB/s (B200_PROFILING.md)"


def ncu_traffic(workload, density, mode, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of this
    workload (profiles/ncu_traffic.json, written by profiles/ncu_traffic.py from the .ncu-rep); None if not captured."""
    try:
        table = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        e = table[workload][f"{density:g}"][mode][kernel]
        return float(e["dram_read_bytes"]) + float(e["dram_write_bytes"])
    except (OSError, KeyError, ValueError):
        return None


def gen_device(wah, gen, n_words, density, seed, dev):
    if gen == "uniform":
        return wah.gen_uniform_device(n_words, density, seed, dev)
    return wah.gen_clustered_device(n_words, density, MEAN_RUN, seed, dev)


def popcount_words(t):
    """set bits of an int32 tensor, on the device, in chunks"""
    import torch

    total = 0
    for i in range(0, t.numel(), 1 << 26):
        x = t[i:i + (1 << 26)].to(torch.int64) & 0xFFFFFFFF
        x = x - ((x >> 1) & 0x55555555)
        x = (x & 0x33333333) + ((x >> 2) & 0x33333333)
        x = (x + (x >> 4)) & 0x0F0F0F0F
        total += int(((x * 0x01010101) >> 24 & 0xFF).sum().item())
    return total


def popcount_stream(wah, d_stream, c_words):
    import torch

    bits = torch.zeros(1, dtype=torch.int64, device=d_stream.device)
    wah.popcount_device(d_stream, c_words, bits)
    return int(bits.item())


def check_prefix_against_oracle(orc, np, x, d_out, c, mode, what):
    """the oracle on the first 2114 * 992 words (whole 1024-group blocks) of a vector: BLOCK1024 streams start with
    exactly those words, CANONICAL ones with all but the prefix's last word (a run may go on behind the prefix)"""
    k = min(x.numel(), 2114 * 992)
    want = orc.compress(x[:k].cpu().numpy().view(np.uint32), mode)
    keep = want.size if (mode == 0 or k == x.numel()) else want.size - 1
    got = d_out[:keep].cpu().numpy().view(np.uint32)
    assert keep <= c and np.array_equal(got, want[:keep]), f"{what}: stream differs from the oracle's on the first {k} words"


class Vector:
    """One bitvector resident in HBM with everything a compress / decompress step needs."""

    def __init__(self, wah, x, mode, dev):
        import torch

        self.wah, self.x, self.mode, self.n = wah, x, mode, x.numel()
        self.cap = wah.max_compressed_words(self.n)
        self.out = torch.empty(self.cap, dtype=torch.int32, device=dev)
        self.cnt = torch.zeros(1, dtype=torch.int64, device=dev)
        self.dec = torch.empty(self.n + 32, dtype=torch.int32, device=dev)
        self.info = torch.zeros(3, dtype=torch.int64, device=dev)
        self.ws_c = wah.Workspace.for_compress(self.n, dev)
        self.compress()
        self.c = int(self.cnt.item())
        self.ws_d = wah.Workspace.for_decompress(self.c, self.n + 32, dev)

    def compress(self):
        self.wah.compress_device(self.x, self.n, self.out, self.cap, self.cnt, self.ws_c, self.mode)

    def decompress(self):
        # (the caller of decompress knows the compressed size, exactly like the reference's outputSize)
        self.wah.decompress_device(self.out, self.c, self.dec, self.n + 32, self.info, self.ws_d)

    def verify(self, orc, np, what):
        import torch

        self.compress()
        self.decompress()
        torch.cuda.synchronize()
        words, groups, status = self.info.tolist()
        assert status == 0, f"{what}: decode status {status:#x}"
        assert int(self.cnt.item()) == self.c and words in (self.n, self.n + 1), f"{what}: sizes"
        assert torch.equal(self.dec[: self.n], self.x), f"{what}: device round trip failed"
        assert popcount_stream(self.wah, self.out, self.c) == popcount_words(self.x), f"{what}: set bits differ"
        check_prefix_against_oracle(orc, np, self.x, self.out, self.c, self.mode, what)


def time_kernels(fn_c, fn_d, reps, stream, batch=1):
    """CUDA events on the launching stream around every launch (batch = 1) or around batches of consecutive launches of
    one kernel; returns per-launch milliseconds {mean, median, best} for compress and decompress"""
    import torch

    def one(fn):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for a, b in ev:
            a.record(stream)
            for _ in range(batch):
                fn()
            b.record(stream)
        torch.cuda.synchronize()
        ts = sorted(a.elapsed_time(b) / batch for a, b in ev)
        return {"ms": sum(ts) / len(ts), "ms_median": ts[len(ts) // 2], "ms_best": ts[0], "launches": reps * batch}

    return one(fn_c), one(fn_d)


def kernel_entry(tc, td, n_words, c_words, peak):
    alg = 4.0 * (n_words + c_words)
    out = {}
    for name, t in (("compress", tc), ("decompress", td)):
        gbs = alg / (t["ms"] * 1e-3) / 1e9
        out[name] = {**t, "achieved": gbs, "frac": gbs / peak, "uncompressed_gbs": 4.0 * n_words / (t["ms"] * 1e-3) / 1e9}
    return out


# ----------------------------------------------------------------------------- N = 1


def run_single(args):
    import numpy as np
    import torch

    import gpu_wah_b200 as wah

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream()
    orc = oracle()
    peak, peak_src = load_peak()
    n_words, gen, density, desc = WORKLOADS[args.workload]
    if args.density is not None:
        density = args.density
    mode = wah.WAH_BLOCK1024 if args.mode == "block1024" else wah.WAH_CANONICAL
    big = 4 * n_words >= (1 << 30)   # a step touches far more than the 126 MB L2
    nbuf = 1 if big else 4

    # ---- headline: K steps of compress + decompress between two events
    vecs = [Vector(wah, gen_device(wah, gen, n_words, density, 1337 + b, dev), mode, dev) for b in range(nbuf)]
    vecs[0].verify(orc, np, f"{args.workload} d={density}")

    def step(i):
        v = vecs[i % nbuf]
        v.compress()
        v.decompress()

    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(stream)
    for i in range(args.steps):
        step(i)
    t1.record(stream)
    torch.cuda.synchronize()
    ms_per_step = t0.elapsed_time(t1) / args.steps
    # per-kernel launch durations of the same launches: an event pair around every launch (a launch lasts >= 0.3 ms at
    # this size; the small vectors, whose launches last 40 us, are timed in batches of 10 as well)
    it = [0]

    def fc():
        vecs[it[0] % nbuf].compress()

    def fd():
        vecs[it[0] % nbuf].decompress()
        it[0] += 1

    reps = max(args.steps, 10)
    tc, td = time_kernels(fc, fd, reps, stream)
    clocks = sampler.stop()
    c_avg = sum(v.c for v in vecs) / nbuf
    head = kernel_entry(tc, td, n_words, c_avg, peak)
    if not big:
        bc, bd = time_kernels(fc, fd, max(reps // 10, 3), stream, batch=10)
        head["batched_launches"] = kernel_entry(bc, bd, n_words, c_avg, peak)
    head_x = vecs[0].x
    h_in = None
    if not args.no_e2e or not args.no_cpu_baseline:
        h_in = head_x.cpu().numpy().view(np.uint32)   # pageable host memory, like the malloc()ed input of source.cpp:75
    del vecs

    # ---- the density sweep of configs[2], both encoder modes; the roofline is the worst kernel of it
    sweep = []
    worst = None
    if args.workload == "clustered_16gbit" and not args.no_sweep:
        for d in SWEEP:
            x = gen_device(wah, "clustered", n_words, d, 1337, dev)
            for m, mname in ((wah.WAH_BLOCK1024, "block1024"), (wah.WAH_CANONICAL, "canonical")):
                v = Vector(wah, x, m, dev)
                v.verify(orc, np, f"sweep d={d} {mname}")
                for _ in range(2):
                    v.compress()
                    v.decompress()
                stc, std = time_kernels(v.compress, v.decompress, 8, stream)
                e = {"density": d, "mode": mname, "compressed_words": v.c, "ratio": v.c / n_words,
                     **kernel_entry(stc, std, n_words, v.c, peak)}
                sweep.append(e)
                for kname, key in (("wah_compress_kernel", "compress"), ("wah_decode_kernel", "decompress")):
                    if worst is None or e[key]["frac"] < worst["frac"]:
                        worst = {"kernel": kname, "density": d, "mode": mname, "frac": e[key]["frac"], "achieved": e[key]["achieved"],
                                 "launch_ms": e[key]["ms"], "algorithmic_bytes_per_launch": 4.0 * (n_words + v.c)}
                del v
            del x
    if worst is None:
        key = "compress" if head["compress"]["frac"] < head["decompress"]["frac"] else "decompress"
        worst = {"kernel": "wah_compress_kernel" if key == "compress" else "wah_decode_kernel", "density": density, "mode": args.mode,
                 "frac": head[key]["frac"], "achieved": head[key]["achieved"], "launch_ms": head[key]["ms"],
                 "algorithmic_bytes_per_launch": 4.0 * (n_words + c_avg)}
    roofline = {
        "bound": "hbm", "kernel": worst["kernel"], "achieved": worst["achieved"], "peak": peak, "unit": "GB/s", "frac": worst["frac"],
        "traffic": ncu_traffic(args.workload, worst["density"], worst["mode"], worst["kernel"]),
        "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full of this kernel on this workload (profiles/ncu_traffic.json); null = no committed capture",
        "peak_source": peak_src, "frac_of_nominal_8tbs": worst["achieved"] / 8000.0,
        "worst_of": f"both kernels x {len(SWEEP)} densities x 2 modes of the {args.workload} sweep" if sweep else "the two kernels of the headline workload",
        "at": {"density": worst["density"], "mode": worst["mode"]},
        "algorithmic_bytes_per_launch": worst["algorithmic_bytes_per_launch"], "launch_ms": worst["launch_ms"],
        "launch_ms_note": "CUDA events on the launching stream around every launch, mean over the launches",
        "headline": head,
    }

    # ---- other configurations as extra keys
    extras = {}
    if not args.no_extras:
        for name in ("sparse_1gbit", "dense_1gbit"):
            if name == args.workload:
                continue
            nw, g, d, _ = WORKLOADS[name]
            vs = [Vector(wah, gen_device(wah, g, nw, d, 4242 + b, dev), wah.WAH_BLOCK1024, dev) for b in range(4)]
            vs[0].verify(orc, np, name)
            k = [0]

            def xc():
                vs[k[0] % 4].compress()

            def xd():
                vs[k[0] % 4].decompress()
                k[0] += 1

            for _ in range(8):
                xc()
                xd()
            bc, bd = time_kernels(xc, xd, 10, stream, batch=10)
            ca = sum(v.c for v in vs) / 4
            extras[name] = {"words": nw, "density": d, "mode": "block1024", "compressed_words": ca,
                            "l2": "4 distinct vectors rotated (a step touches 272 MiB > 126 MB L2); launch durations from batches of 10 consecutive launches",
                            **kernel_entry(bc, bd, nw, ca, peak)}
            del vs
        bm = bitmap_block(wah, orc, np, torch, dev, stream, args, 0, 1, peak, None)
        for k in ("_x", "_mode", "_total_ms"):
            bm.pop(k)
        extras["bitmap_index"] = bm

    # ---- end to end through the reference-facing host entry points
    e2e = None
    if not args.no_e2e:
        e2e = e2e_single(wah, np, torch, h_in, n_words, mode, args.e2e_steps, head_x)

    cpu_baseline = None
    if not args.no_cpu_baseline:
        threads = host_threads()
        r = cpu_round_trips(orc, h_in, 0 if mode == wah.WAH_BLOCK1024 else 1, 10.0, threads)
        cpu_baseline = {"value": r["value"], "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": f"the whole {n_words}-word vector, {r['reps']} round trips",
                        "compress_gbs": r["compress_gbs"], "decompress_gbs": r["decompress_gbs"]}

    ref_ctx = None
    if not args.no_cpu_baseline and h_in is not None and n_words >= 33_554_400:
        ref_ctx = reference_kernels_context(np, h_in, density)

    nbytes = 4.0 * n_words
    emit({
        "metric": METRIC, "value": 2 * nbytes / (ms_per_step * 1e-3) / 1e9, "unit": UNIT, "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": args.workload, "description": desc, "words": n_words, "density": density, "mode": args.mode,
                   "compressed_words": c_avg, "ratio": c_avg / n_words, "step": "compress the vector, then decompress it",
                   "l2": "the vector is 2 GiB: every launch streams 16x the 126 MB L2 from / to HBM" if big else
                         f"{nbuf} distinct vectors rotated so that no step finds its input in the 126 MB L2"},
        "compress_gbs": head["compress"]["uncompressed_gbs"], "decompress_gbs": head["decompress"]["uncompressed_gbs"],
        "clocks": clocks, "e2e": e2e, "gpu_launches": 2 * args.steps,
        "gpu_launches_note": "per step: wah_compress_kernel, wah_decode_kernel (scan + expand fused); no memset, no other kernel",
        "roofline": roofline, "sweep": sweep, **extras, "cpu_baseline": cpu_baseline, "reference_kernels": ref_ctx, "parity_checked": True,
    })


def reference_kernels_context(np, h_in, density):
    """Context, not a target (north_star: "the reference's original CUDA kernels on the same B200"): the reference's own
    compress() / decompress(), untouched, built for sm_100a with the three-macro shuffle shim (oracle/_ref/
    libgpuwah_ref.so, made by oracle/Makefile from the sources where they lie), on the first 33 554 400 words of the
    workload vector (the reference is defined for n % 992 == 0; its host path is pageable and serial), in a process of
    its own (tests/ref_runner.py), medians of its own three timers over 10 repetitions like its benchmark loop
    (source.cpp:70,83-126).  TEST INFRASTRUCTURE: nothing of it is on the measured path."""
    import subprocess
    import tempfile

    lib = os.path.join(ROOT, "oracle", "_ref", "libgpuwah_ref.so")
    if not os.path.exists(lib):
        return {"unavailable": "oracle/_ref/libgpuwah_ref.so has not been built (python -c 'import __graft_entry__ as g; g.build()' where /root/reference is mounted)"}
    n = 33_554_400
    try:
        with tempfile.TemporaryDirectory() as d:
            fin, fout = os.path.join(d, "in.npy"), os.path.join(d, "out.npy")
            np.save(fin, h_in[:n])
            subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ref_runner.py"), lib, "time", fin, fout], check=True, timeout=300,
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            t = np.load(fout)
    except Exception as e:  # the reference is what it is: context only
        return {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    nb = 4.0 * n
    return {"words": n, "density": density, "sample": "the first 33 554 400 words of the workload vector (n % 992 == 0)",
            "compress_ms": {"h2d": float(t[0][0]), "compute": float(t[0][1]), "d2h": float(t[0][2])},
            "decompress_ms": {"h2d": float(t[1][0]), "compute": float(t[1][1]), "d2h": float(t[1][2])},
            "compress_gbs_compute_timer": nb / float(t[0][1]) / 1e6, "decompress_gbs_compute_timer": nb / float(t[1][1]) / 1e6,
            "e2e_gbs": 2 * nb / float(t[0].sum() + t[1].sum()) / 1e6, "unit": UNIT,
            "note": "the reference's kernels recompiled with a shuffle shim, its own timers; reported for context"}


def cpu_round_trips(orc, data, mode, budget_s, threads):
    """oracle (OpenMP) compress + decompress round trips on `data`; GB/s of uncompressed bytes"""
    cw = orc.compress(data, mode, threads=threads)   # warm-up + output for the decoder
    reps, t_total, t_c, t_d = 0, 0.0, 0.0, 0.0
    while reps < 3 or (t_total < budget_s and reps < 200):
        a = time.perf_counter()
        cw = orc.compress(data, mode, threads=threads)
        b = time.perf_counter()
        back = orc.decompress(cw, threads=threads)
        c = time.perf_counter()
        t_c, t_d, t_total, reps = t_c + b - a, t_d + c - b, t_total + c - a, reps + 1
    assert back[: data.size].tobytes() == data.tobytes()
    nbytes = data.size * 4
    return {"value": 2 * nbytes * reps / t_total / 1e9, "compress_gbs": nbytes * reps / t_c / 1e9,
            "decompress_gbs": nbytes * reps / t_d / 1e9, "reps": reps}


def e2e_single(wah, np, torch, h_in, n_words, mode, steps, d_x):
    """wah_compress_host + wah_decompress_host on a PAGEABLE input (what the reference's callers pass); malloc()ed
    results freed by the caller.  Secondary figures: the same from a page-locked input, and with caller-provided
    page-locked result buffers."""
    lib = wah.lib
    outp, outn = ctypes.c_void_p(), ctypes.c_uint64()
    decp, decn = ctypes.c_void_p(), ctypes.c_uint64()
    cap = wah.max_compressed_words(n_words)
    nbytes = 4.0 * n_words

    moved = [0, 0]   # bytes the last round trip moved over PCIe, each way (the library's own count)
    t_h2d, t_d2h = ctypes.c_uint64(), ctypes.c_uint64()

    def round_trip(src_ptr, check=False):
        rc = lib.wah_compress_host(src_ptr, n_words, mode, ctypes.byref(outp), ctypes.byref(outn), None, None, None)
        assert rc == 0, lib.wah_last_error_string()
        lib.wah_host_last_transfer_bytes(ctypes.byref(t_h2d), ctypes.byref(t_d2h))
        moved[0], moved[1] = t_h2d.value, t_d2h.value
        rc = lib.wah_decompress_host(outp.value, outn.value, ctypes.byref(decp), ctypes.byref(decn), None, None, None)
        assert rc == 0, lib.wah_last_error_string()
        lib.wah_host_last_transfer_bytes(ctypes.byref(t_h2d), ctypes.byref(t_d2h))
        moved[0] += t_h2d.value
        moved[1] += t_d2h.value
        c, n = outn.value, decn.value
        if check:
            back = np.frombuffer((ctypes.c_uint32 * n_words).from_address(decp.value), dtype=np.uint32)
            assert np.array_equal(back, h_in), "host round trip failed"
        lib.wah_free(outp)
        lib.wah_free(decp)
        return c, n

    def timed(fn):
        fn()
        t0 = time.perf_counter()
        for _ in range(steps):
            c_, n_ = fn()
        torch.cuda.synchronize()
        return time.perf_counter() - t0, c_, n_

    round_trip(h_in.ctypes.data, check=True)
    dt, c_e2e, n_e2e = timed(lambda: round_trip(h_in.ctypes.data))
    moved_e2e = tuple(moved)
    # the reference's three timers (ms): H2D / compute / D2H, one extra untimed step
    fl = [ctypes.c_float() for _ in range(6)]
    lib.wah_compress_host(h_in.ctypes.data, n_words, mode, ctypes.byref(outp), ctypes.byref(outn),
                          ctypes.byref(fl[0]), ctypes.byref(fl[1]), ctypes.byref(fl[2]))
    lib.wah_decompress_host(outp.value, outn.value, ctypes.byref(decp), ctypes.byref(decn),
                            ctypes.byref(fl[3]), ctypes.byref(fl[4]), ctypes.byref(fl[5]))
    lib.wah_free(outp)
    lib.wah_free(decp)
    segments = dict(zip(("compress_h2d", "compress_compute", "compress_d2h", "decompress_h2d", "decompress_compute", "decompress_d2h"),
                        (f.value for f in fl)))
    # secondary: page-locked input; page-locked caller-provided result buffers
    h_pin = torch.empty(n_words, dtype=torch.int32).pin_memory()
    h_pin.copy_(d_x)
    torch.cuda.synchronize()
    dt_pin, _, _ = timed(lambda: round_trip(h_pin.data_ptr()))
    h_comp = torch.empty(cap, dtype=torch.int32).pin_memory()
    h_dec = torch.empty(n_words + 32, dtype=torch.int32).pin_memory()

    def into_step():
        rc = lib.wah_compress_host_into(h_pin.data_ptr(), n_words, mode, h_comp.data_ptr(), cap, ctypes.byref(outn))
        assert rc == 0, lib.wah_last_error_string()
        rc = lib.wah_decompress_host_into(h_comp.data_ptr(), outn.value, h_dec.data_ptr(), n_words + 32, ctypes.byref(decn))
        assert rc == 0, lib.wah_last_error_string()
        return outn.value, decn.value

    dt_into, _, _ = timed(into_step)
    assert torch.equal(h_dec[:n_words], h_pin), "host round trip failed"
    return {
        "value": 2 * nbytes * steps / dt / 1e9, "unit": UNIT,
        "h2d_bytes_per_step": int(moved_e2e[0]), "d2h_bytes_per_step": int(moved_e2e[1] + 48),
        "bytes_note": f"counted by the library: 4 KiB blocks that are all zero do not cross PCIe (the vector itself is {int(nbytes)} bytes each way; "
                      "WAH_B200_SPARSE_COPY=0 moves every byte)",
        "steps": steps, "ms_per_step": dt / steps * 1e3, "segments_ms": segments,
        "api": "wah_compress_host + wah_decompress_host (= the reference's compress()/decompress()): PAGEABLE malloc()ed input "
               "(source.cpp:75,97-100), malloc()ed results freed by the caller",
        "pinned_input": {"value": 2 * nbytes * steps / dt_pin / 1e9, "unit": UNIT, "ms_per_step": dt_pin / steps * 1e3},
        "caller_buffers": {"value": 2 * nbytes * steps / dt_into / 1e9, "unit": UNIT, "ms_per_step": dt_into / steps * 1e3,
                           "api": "wah_compress_host_into + wah_decompress_host_into, page-locked input and result buffers"},
    }


# ----------------------------------------------------------------------------- bitmap index (configs[3]), any N


def bitmap_block(wah, orc, np, torch, dev, stream, args, rank, world, peak, dist):
    """This rank's columns of the 1024 x 64 Mbit bitmap index: one batched compress launch and one batched decode launch
    per step.  With more than one rank the column lengths are all-gathered inside the step (the only exchange; it
    overlaps the decode).  Returns the block of the JSON line (timings are this rank's; the caller takes the max)."""
    from gpu_wah_b200 import mgpu

    n_cols_all, wpc = args.cols, COL_WORDS
    c0, c1 = mgpu.column_range(n_cols_all, rank, world)
    n_cols = c1 - c0
    mode = wah.WAH_BLOCK1024 if args.mode == "block1024" else wah.WAH_CANONICAL
    # every rank's block is a stretch of one long run-clustered vector (seeded per rank), cut into columns
    x = gen_device(wah, "clustered", n_cols * wpc, 0.01, 7000 + rank, dev)
    cap = n_cols * wpc // 8 + 65536     # these columns compress 400 : 1; the kernel enforces the capacity anyway
    out = torch.empty(cap, dtype=torch.int32, device=dev)
    offs = torch.zeros(n_cols + 1, dtype=torch.int64, device=dev)
    ws_c = wah.Workspace.for_compress_batch(n_cols, wpc, dev)
    stride = wpc + 4
    back = torch.empty(n_cols * stride, dtype=torch.int32, device=dev)
    info = torch.zeros(3, dtype=torch.int64, device=dev)

    def compress():
        wah.compress_batch_device(x, n_cols, wpc, wpc, out, cap, offs, ws_c, mode)

    compress()
    h_offs = offs.cpu().numpy()
    c_total = int(h_offs[-1])
    assert c_total <= cap
    ws_d = wah.Workspace.for_decompress_batch(n_cols, c_total, wpc, dev)

    def decompress():
        wah.decompress_batch_device(out, c_total, n_cols, wpc, back, stride, wpc + 1, info, ws_d)

    # ---- parity before timing: columns against the oracle, all columns against single-stream compress on the
    #      device, set bits, round trip
    decompress()
    torch.cuda.synchronize()
    assert info.tolist() == [wah.decoded_words(wah.num_groups(wpc)), wah.num_groups(wpc) * n_cols, 0], f"batch decode info {info.tolist()}"
    assert torch.equal(back.view(n_cols, stride)[:, :wpc], x.view(n_cols, wpc)), "bitmap index: device round trip failed"
    assert popcount_stream(wah, out, c_total) == popcount_words(x), "bitmap index: set bits differ"
    for j in sorted({0, n_cols // 2, n_cols - 1}):
        want = orc.compress(x[j * wpc:(j + 1) * wpc].cpu().numpy().view(np.uint32), 0 if mode == wah.WAH_BLOCK1024 else 1)
        got = out[int(h_offs[j]):int(h_offs[j + 1])].cpu().numpy().view(np.uint32)
        assert np.array_equal(got, want), f"bitmap index: column {c0 + j} differs from the oracle's"
    one = Vector(wah, x[:wpc], mode, dev)
    for j in range(0, n_cols, max(1, n_cols // 16)):
        one.x = x[j * wpc:(j + 1) * wpc]
        one.compress()
        cj = int(one.cnt.item())
        assert cj == int(h_offs[j + 1] - h_offs[j]) and torch.equal(one.out[:cj], out[int(h_offs[j]):int(h_offs[j + 1])]), \
            f"bitmap index: column {c0 + j} differs from a single-stream compress"
    del one

    lens = torch.empty(n_cols_all, dtype=torch.int64, device=dev) if world > 1 else None
    even = n_cols_all % world == 0

    def step():
        compress()
        work = None
        if world > 1:
            mine = offs[1:] - offs[:-1]
            if even:
                work = dist.all_gather_into_tensor(lens, mine, async_op=True)   # NCCL, overlaps the decode below
            else:
                pieces = [lens[slice(*mgpu.column_range(n_cols_all, r, world))] for r in range(world)]
                work = dist.all_gather(pieces, mine, async_op=True)
        decompress()
        if work is not None:
            work.wait()

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(stream)
    for _ in range(args.steps):
        step()
    t1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    total_ms = t0.elapsed_time(t1)
    tc, td = time_kernels(compress, decompress, max(args.steps, 10), stream)
    if world > 1:
        assert lens.cpu().tolist()[c0:c1] == np.diff(h_offs).tolist(), "all-gathered column lengths differ"
    nbytes = 4.0 * n_cols * wpc
    block = {"columns": n_cols_all, "columns_this_rank": n_cols, "words_per_column": wpc, "density": 0.01, "mode": args.mode,
             "compressed_words_this_rank": c_total, "ms_per_step": total_ms / args.steps,
             "value": 2 * nbytes / (total_ms / args.steps * 1e-3) / 1e9, "unit": UNIT,
             "launches_per_step": "1 wah_compress_kernel (all columns) + 1 wah_decode_kernel (all columns)" + (" + 1 NCCL all-gather of the column lengths" if world > 1 else ""),
             **kernel_entry(tc, td, n_cols * wpc, c_total, peak), "parity_checked": True}
    block["_total_ms"] = total_ms
    block["_x"], block["_mode"] = x, mode
    return block


# ----------------------------------------------------------------------------- N > 1


def run_multi(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import gpu_wah_b200 as wah
    from gpu_wah_b200 import mgpu

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}"
    # the ranks of one node share its host cores: split them between the ranks' copy threads (host path only)
    if "WAH_B200_COPY_THREADS" not in os.environ:
        os.environ["WAH_B200_COPY_THREADS"] = str(max(2, host_threads() // world))
    assert torch.cuda.is_available(), "bench.py needs CUDA devices (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.current_stream()
    orc = oracle()
    peak, peak_src = load_peak()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    # ---- configs[3]: the bitmap index, columns split across the ranks (value)
    bm = bitmap_block(wah, orc, np, torch, dev, stream, args, rank, world, peak, dist)
    x_cols, mode = bm.pop("_x"), bm.pop("_mode")
    t = torch.tensor([bm.pop("_total_ms"), bm["compress"]["ms"], bm["decompress"]["ms"]], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, tc_ms, td_ms = t.tolist()
    csum = torch.tensor([bm["compressed_words_this_rank"]], dtype=torch.int64, device=dev)
    dist.all_reduce(csum)
    clocks = sampler.stop() if sampler else None
    ms_per_step = total_ms / args.steps
    all_bytes = 4.0 * args.cols * COL_WORDS
    value = 2 * all_bytes / (ms_per_step * 1e-3) / 1e9

    # ---- e2e: the same columns through the reference-facing host entry points, one compress() + decompress() per column
    e2e = None
    if not args.no_e2e:
        e2e = e2e_columns(wah, np, torch, dist, x_cols, mode, args, rank, world, dev)
    del x_cols

    # ---- configs[4]: one 128 Gbit vector, range-sharded
    rng_block = None
    if not args.no_range:
        rng_block = range_block(wah, mgpu, orc, np, torch, dist, dev, stream, args, rank, world, peak)

    if rank == 0:
        per_rank_alg = 4.0 * (args.cols * COL_WORDS + int(csum.item())) / world
        worst_key = "compress" if tc_ms >= td_ms else "decompress"
        worst_ms = max(tc_ms, td_ms)
        roofline = {
            "bound": "hbm", "kernel": "wah_compress_kernel" if worst_key == "compress" else "wah_decode_kernel",
            "achieved": per_rank_alg / (worst_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
            "frac": per_rank_alg / (worst_ms * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
            "note": "per GPU: this rank's columns, algorithmic bytes 4 (n + c) / slowest rank's launch duration; the slower of the two batched kernels",
            "launch_ms": worst_ms, "algorithmic_bytes_per_launch": per_rank_alg,
            "compress": {"ms": tc_ms, "frac": per_rank_alg / (tc_ms * 1e-3) / 1e9 / peak},
            "decompress": {"ms": td_ms, "frac": per_rank_alg / (td_ms * 1e-3) / 1e9 / peak},
        }
        emit({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic",
            "config": {"workload": "bitmap_index", "description": BITMAP_DESC, "columns": args.cols, "words_per_column": COL_WORDS,
                       "density": 0.01, "mode": args.mode, "columns_per_rank": args.cols // world,
                       "compressed_words": int(csum.item()), "step": "every rank: ONE batched compress launch over its columns, "
                       "all-gather of the column lengths (NCCL), ONE batched decode launch",
                       "l2": f"every rank streams {8 * 1024 // world} MiB per launch: far larger than the 126 MB L2",
                       "single_gpu_point": "the N = 1 line carries the same workload on one GPU as `bitmap_index`"},
            "compress_gbs": all_bytes / (tc_ms * 1e-3) / 1e9, "decompress_gbs": all_bytes / (td_ms * 1e-3) / 1e9,
            "clocks": clocks, "e2e": e2e, "gpu_launches": 2 * args.steps * world,
            "gpu_launches_note": "per step and rank: wah_compress_kernel (batch), wah_decode_kernel (batch); the all-gather is NCCL's kernel",
            "roofline": roofline, "bitmap_index_rank0": bm, "range_128gbit": rng_block, "cpu_baseline": None, "parity_checked": True,
        })
    dist.barrier()
    dist.destroy_process_group()


def e2e_columns(wah, np, torch, dist, x_cols, mode, args, rank, world, dev):
    """every column of this rank through wah_compress_host + wah_decompress_host (pageable input, malloc()ed results):
    what a caller of the reference's compress() / decompress() does with a bitmap index -- one call pair per column"""
    lib = wah.lib
    wpc = COL_WORDS
    n_cols = x_cols.numel() // wpc
    h = x_cols.cpu().numpy().view(np.uint32).reshape(n_cols, wpc)
    outp, outn = ctypes.c_void_p(), ctypes.c_uint64()
    decp, decn = ctypes.c_void_p(), ctypes.c_uint64()
    tot = [0, 0]
    moved = [0, 0]   # bytes moved over PCIe each way (the library's own count: all-zero 4 KiB blocks stay where they are)
    t_h2d, t_d2h = ctypes.c_uint64(), ctypes.c_uint64()

    def one_pass(check=False):
        tot[0] = tot[1] = moved[0] = moved[1] = 0
        for j in range(n_cols):
            rc = lib.wah_compress_host(h[j].ctypes.data, wpc, mode, ctypes.byref(outp), ctypes.byref(outn), None, None, None)
            assert rc == 0, lib.wah_last_error_string()
            lib.wah_host_last_transfer_bytes(ctypes.byref(t_h2d), ctypes.byref(t_d2h))
            moved[0] += t_h2d.value
            moved[1] += t_d2h.value
            rc = lib.wah_decompress_host(outp.value, outn.value, ctypes.byref(decp), ctypes.byref(decn), None, None, None)
            assert rc == 0, lib.wah_last_error_string()
            lib.wah_host_last_transfer_bytes(ctypes.byref(t_h2d), ctypes.byref(t_d2h))
            moved[0] += t_h2d.value
            moved[1] += t_d2h.value
            if check and j % 37 == 0:
                back = np.frombuffer((ctypes.c_uint32 * wpc).from_address(decp.value), dtype=np.uint32)
                assert np.array_equal(back, h[j]), "host round trip failed"
            tot[0] += outn.value
            tot[1] += decn.value
            lib.wah_free(outp)
            lib.wah_free(decp)

    one_pass(check=True)
    steps = max(1, min(args.e2e_steps, 3))
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_pass()
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    dt = dt.item()
    nbytes = 4.0 * args.cols * wpc
    return {"value": 2 * nbytes * steps / dt / 1e9, "unit": UNIT,
            "h2d_bytes_per_step": int(moved[0]), "d2h_bytes_per_step": int(moved[1] + 48 * n_cols),
            "steps": steps, "ms_per_step": dt / steps * 1e3, "bytes_are": "this rank's (every rank moves its own columns)",
            "api": "per column: wah_compress_host + wah_decompress_host (= the reference's compress()/decompress()), pageable input, malloc()ed results"}


def range_block(wah, mgpu, orc, np, torch, dist, dev, stream, args, rank, world, peak):
    """configs[4]: one 128 Gbit vector split into ranges of whole 992-word blocks.  Step: local compress, shard record,
    all-gather of the records (NCCL), seam plan (host), all-gather-v of the segments into the global stream on every
    rank (NCCL send / recv), seam words patched, local decompress of the own range."""
    n_all = 1 << args.range_log2
    lo, hi = mgpu.word_range(n_all, rank, world)
    n = hi - lo
    out_block = {"words": n_all, "words_this_rank": n, "density": 0.01, "description": RANGE_DESC, "modes": {}}
    x = gen_device(wah, "clustered", n, 0.01, 9000 + rank, dev)
    backend = mgpu.CudaBackend(dev)
    dec = torch.empty(n + 32, dtype=torch.int32, device=dev)
    info = torch.zeros(3, dtype=torch.int64, device=dev)
    # the whole vector on every rank, for the single-GPU reference stream of the parity check (untimed)
    sizes = [mgpu.word_range(n_all, r, world) for r in range(world)]
    full = torch.empty(n_all, dtype=torch.int32, device=dev)
    dist.all_gather([full[a:b] for a, b in sizes], x)
    for mode, mname in ((wah.WAH_BLOCK1024, "block1024"), (wah.WAH_CANONICAL, "canonical")):
        # ---- parity before timing
        ss = mgpu.compress_range_sharded(x, mode, backend=backend)
        stream_all = mgpu.gather_stream(ss)
        whole = Vector(wah, full, mode, dev)
        assert whole.c == ss.total_words and torch.equal(stream_all, whole.out[: whole.c]), \
            f"range-sharded {mname}: the gathered stream differs from a single-GPU compress of the whole vector"
        assert popcount_stream(wah, stream_all, ss.total_words) == popcount_words(full), f"range-sharded {mname}: set bits differ"
        seams = sum(1 for w in ss.plan["seam_words"] if w)
        del whole
        check_prefix_against_oracle(orc, np, x, ss.segment, ss.segment.numel(), mode, f"range shard of rank {rank} ({mname})")
        c_local = ss.segment.numel()
        ws_d = wah.Workspace.for_decompress(c_local, n + 32, dev)
        wah.decompress_device(ss.segment, c_local, dec, n + 32, info, ws_d)
        torch.cuda.synchronize()
        assert info.tolist()[2] == 0 and torch.equal(dec[:n], x), f"range-sharded {mname}: local round trip failed"
        total_words = ss.total_words
        del ss, stream_all

        def step():
            s = mgpu.compress_range_sharded(x, mode, backend=backend)
            g = mgpu.gather_stream(s)
            wah.decompress_device(s.segment, s.segment.numel(), dec, n + 32, info, ws_d)
            return g

        for _ in range(2):
            step()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = max(3, min(args.steps, 10))
        t0.record(stream)
        for _ in range(steps):
            step()
        t1.record(stream)
        torch.cuda.synchronize()
        dist.barrier()
        t = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item() / steps
        out_block["modes"][mname] = {
            "ms_per_step": ms, "value": 2 * 4.0 * n_all / (ms * 1e-3) / 1e9, "unit": UNIT, "steps": steps,
            "compressed_words_total": total_words, "seams_merged": seams,
            "nccl_bytes_per_step": {"records_all_gather": 56 * world, "segments_all_gather_v_received_per_rank": 4 * (total_words - c_local)},
            "step": "local compress, shard record, all-gather of 56-byte records, seam plan, all-gather-v of the segments "
                    "(every rank ends up with the whole stream), seam patch, local decompress",
            "parity_checked": True,
        }
    del full
    return out_block


_REAL_STDOUT = None


def emit(line):
    """the ONE line of the bench contract, on the process's original stdout"""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # Libraries talk on fd 1 too (NCCL prints its version there under torchrun): everything but the result line goes
    # to stderr, so that stdout carries exactly one JSON line.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    elif args.gpus > 1 or int(os.environ.get("WORLD_SIZE", "1")) > 1:
        run_multi(args)
    else:
        run_single(args)


if __name__ == "__main__":
    main()
