#!/usr/bin/env python
"""bench.py -- WAH compress/decompress throughput on B200 (the metric of BASELINE.json).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload NAME]

One "step" = one pass of the hot path over one batch of synthetic input: compress the
bitvector, then decompress the result.  Default workload at every N is BASELINE.json
configs[1], "1 Gbit sparse bitvector (density 0.001)": each rank owns one such vector
(independent objects, no data-path collective: weak scaling).  `value` = uncompressed bytes
through the path per second over all ranks (4n in on compress + 4n out on decompress),
inputs resident in HBM; `e2e` = the same through the host-buffer C ABI
(wah_compress_host / wah_decompress_host: the reference's compress()/decompress()), host
<-> device copies inside the timed region.

--impl reference times the CPU oracle port (the reference has no CPU implementation) with all
host threads on the same workload.  One JSON line on stdout, from rank 0.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "WAH compress/decompress GB/s (uncompressed bytes)"
UNIT = "GB/s"

WORKLOADS = {
    # name: (words per rank, generator, density, description)
    "sparse_1gbit": (1 << 25, "uniform", 0.001, "1 Gbit sparse bitvector, density 0.001 (BASELINE configs[1])"),
    "uniform_32mbit": (1 << 20, "uniform", 0.5, "32 Mbit uniform random, density 0.5 (BASELINE configs[0])"),
    "clustered_16gbit": (1 << 29, "clustered", 0.01, "16 Gbit run-clustered (Markov, mean run 1000 bits), density 0.01 (configs[2])"),
    "dense_1gbit": (1 << 25, "uniform", 0.5, "1 Gbit uniform random, density 0.5"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="sparse_1gbit", choices=list(WORKLOADS))
    ap.add_argument("--density", type=float, default=None)
    ap.add_argument("--mode", default="block1024", choices=["block1024", "canonical"])
    ap.add_argument("--buffers", type=int, default=4, help="distinct input buffers rotated between steps")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ----------------------------------------------------------------------------- CPU arm


def host_input(n_words, gen, density, seed):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import datagen

    if gen == "uniform":
        return datagen.uniform(n_words, density, seed)
    return datagen.clustered(n_words, density, 1000.0, seed)


def cpu_arm(data, mode, budget_s, threads):
    """oracle (OpenMP) compress + decompress round trips on `data`; returns GB/s of uncompressed bytes"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as orc

    cw = orc.compress(data, mode, threads=threads)   # warm-up + output for the decoder
    reps, t_total = 0, 0.0
    t_c = t_d = 0.0
    while reps < 3 or (t_total < budget_s and reps < 200):
        t0 = time.perf_counter()
        cw = orc.compress(data, mode, threads=threads)
        t1 = time.perf_counter()
        back = orc.decompress(cw, threads=threads)
        t2 = time.perf_counter()
        t_c += t1 - t0
        t_d += t2 - t1
        t_total += t2 - t0
        reps += 1
    assert back[: data.size].tobytes() == data.tobytes()
    nbytes = data.size * 4
    return {
        "value": 2 * nbytes * reps / t_total / 1e9,
        "compress_gbs": nbytes * reps / t_c / 1e9,
        "decompress_gbs": nbytes * reps / t_d / 1e9,
        "reps": reps,
        "ms_per_step": t_total / reps * 1e3,
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as orc

    n_words, gen, density, desc = WORKLOADS[args.workload]
    if args.density is not None:
        density = args.density
    mode = 0 if args.mode == "block1024" else 1
    # all host cores, whatever OMP_NUM_THREADS says (torchrun sets it to 1 for every rank)
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    # bounded sample: the first 2^23 words (256 Mbit) of the workload's vector per step
    sample_words = min(n_words, 1 << 23)
    data = host_input(sample_words, gen, density, 1337)
    for _ in range(max(args.warmup, 1)):
        cw = orc.compress(data, mode, threads=threads)
        orc.decompress(cw, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cw = orc.compress(data, mode, threads=threads)
        back = orc.decompress(cw, threads=threads)
    dt = time.perf_counter() - t0
    assert back[: data.size].tobytes() == data.tobytes()
    value = 2 * sample_words * 4 * args.steps / dt / 1e9
    sample = f"first {sample_words} words ({sample_words * 32 // (1 << 20)} Mbit) of the workload vector per step"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": args.workload, "description": desc, "words_per_rank": n_words, "density": density,
                   "mode": args.mode, "note": "the reference has no CPU implementation; this is the CPU oracle "
                   "port (oracle/wah_oracle.c, OpenMP) on the host cores"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ----------------------------------------------------------------------------- GPU arm


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


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import gpu_wah_b200 as wah

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # the ranks of one node share its host cores: split them between the ranks' copy threads (host path only)
    if world > 1 and "WAH_B200_COPY_THREADS" not in os.environ:
        os.environ["WAH_B200_COPY_THREADS"] = str(max(2, (os.cpu_count() or 16) // world))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n_words, gen, density, desc = WORKLOADS[args.workload]
    if args.density is not None:
        density = args.density
    mode = wah.WAH_BLOCK1024 if args.mode == "block1024" else wah.WAH_CANONICAL
    nbuf = max(1, args.buffers)
    if n_words * 4 * nbuf > 32 << 30:
        nbuf = max(1, (32 << 30) // (n_words * 4))

    # ---- synthetic inputs, resident in HBM; distinct buffers rotated so no step finds its input in L2
    inputs = []
    for b in range(nbuf):
        seed = 1337 + 1000 * rank + b
        if gen == "uniform":
            inputs.append(wah.gen_uniform_device(n_words, density, seed, dev))
        else:
            inputs.append(wah.gen_clustered_device(n_words, density, 1000.0, seed, dev))
    cap = wah.max_compressed_words(n_words)
    d_comp = [torch.empty(cap, dtype=torch.int32, device=dev) for _ in range(min(nbuf, 2))]
    d_cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    d_dec = torch.empty(n_words + 32, dtype=torch.int32, device=dev)
    d_info = torch.zeros(2, dtype=torch.int64, device=dev)
    ws_c = wah.Workspace.for_compress(n_words, dev)

    # compressed sizes (needed by the caller of decompress, exactly like the reference's outputSize)
    c_words = []
    for b in range(nbuf):
        wah.compress_device(inputs[b], n_words, d_comp[0], cap, d_cnt, ws_c, mode)
        c_words.append(int(d_cnt.item()))
    ws_d = wah.Workspace.for_decompress(max(c_words), n_words + 32, dev)

    # correctness of exactly what is timed: round trip on the device + oracle on a slice
    wah.compress_device(inputs[0], n_words, d_comp[0], cap, d_cnt, ws_c, mode)
    wah.decompress_device(d_comp[0], c_words[0], d_dec, n_words + 32, d_info, ws_d)
    torch.cuda.synchronize()
    assert int(d_info[0].item()) in (n_words, n_words + 1)
    assert torch.equal(d_dec[:n_words], inputs[0]), "device round trip failed"

    stream = torch.cuda.current_stream()

    def step(i, events=None):
        b = i % nbuf
        out = d_comp[i % len(d_comp)]
        if events:
            events[0].record(stream)
        wah.compress_device(inputs[b], n_words, out, cap, d_cnt, ws_c, mode)
        if events:
            events[1].record(stream)
        wah.decompress_device(out, c_words[b], d_dec, n_words + 32, d_info, ws_d)
        if events:
            events[2].record(stream)

    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    torch.cuda.synchronize()
    # ---- the timed region: exactly `steps` steps between two events, nothing else on the stream
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record(stream)
    for i in range(args.steps):
        step(i)
    t_end.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    total_ms = t_start.elapsed_time(t_end)
    # ---- per-kernel launch durations for the roofline: batches of BATCH consecutive launches of ONE kernel between two
    #      events (an event pair around every single 40 us launch costs several us of its own and keeps the next kernel
    #      from being scheduled behind the running one, which the timed region above does not suffer from).  Inputs
    #      rotate as in the timed region; the decoder reads a stream per input buffer.
    BATCH = 10
    nbatch = max(1, min(args.steps, 200) // BATCH)
    d_streams = []
    for b in range(nbuf):
        wah.compress_device(inputs[b], n_words, d_comp[0], cap, d_cnt, ws_c, mode)
        d_streams.append(d_comp[0][: c_words[b] + 8].clone())
    torch.cuda.synchronize()
    evc = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(nbatch)]
    evd = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(nbatch)]
    for j in range(nbatch):
        evc[j][0].record(stream)
        for k in range(BATCH):
            i = j * BATCH + k
            wah.compress_device(inputs[i % nbuf], n_words, d_comp[i % len(d_comp)], cap, d_cnt, ws_c, mode)
        evc[j][1].record(stream)
    for j in range(nbatch):
        evd[j][0].record(stream)
        for k in range(BATCH):
            b = (j * BATCH + k) % nbuf
            wah.decompress_device(d_streams[b], c_words[b], d_dec, n_words + 32, d_info, ws_d)
        evd[j][1].record(stream)
    torch.cuda.synchronize()
    bc = sorted(e[0].elapsed_time(e[1]) / BATCH for e in evc)   # per-launch time of every batch
    bd = sorted(e[0].elapsed_time(e[1]) / BATCH for e in evd)
    tc_ms, td_ms = sum(bc) / nbatch, sum(bd) / nbatch
    batch_stats = {"compress": {"batches": nbatch, "ms_median": bc[nbatch // 2], "ms_best": bc[0]},
                   "decompress": {"batches": nbatch, "ms_median": bd[nbatch // 2], "ms_best": bd[0]}}
    # the same two numbers taken inside the alternating sequence, an event between the two halves of every step
    ksteps = min(args.steps, 200)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(ksteps)]
    for i in range(ksteps):
        step(i, ev[i])
    torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    tc_alt_ms = sum(e[0].elapsed_time(e[1]) for e in ev) / ksteps
    td_alt_ms = sum(e[1].elapsed_time(e[2]) for e in ev) / ksteps
    del d_streams
    if world > 1:
        t = torch.tensor([total_ms, tc_ms, td_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, tc_ms, td_ms = t.tolist()
    ms_per_step = total_ms / args.steps
    c_avg = sum(c_words[i % nbuf] for i in range(args.steps)) / args.steps
    nbytes = 4.0 * n_words

    # ---- end to end through the host-buffer C ABI (pinned host input, H2D + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        h_in = torch.empty(n_words, dtype=torch.int32).pin_memory()
        h_in.copy_(inputs[0])
        torch.cuda.synchronize()
        lib = wah.lib
        outp, outn = ctypes.c_void_p(), ctypes.c_uint64()
        decp, decn = ctypes.c_void_p(), ctypes.c_uint64()

        def e2e_step():
            rc = lib.wah_compress_host(h_in.data_ptr(), n_words, mode, ctypes.byref(outp), ctypes.byref(outn), None, None, None)
            assert rc == 0, lib.wah_last_error_string()
            rc = lib.wah_decompress_host(outp.value, outn.value, ctypes.byref(decp), ctypes.byref(decn), None, None, None)
            assert rc == 0, lib.wah_last_error_string()
            c, n = outn.value, decn.value
            lib.wah_free(outp)
            lib.wah_free(decp)
            return c, n

        # the same work with caller-provided page-locked result buffers (wah_*_host_into): pure DMA both ways
        h_comp = torch.empty(cap, dtype=torch.int32).pin_memory()
        h_dec = torch.empty(n_words + 32, dtype=torch.int32).pin_memory()

        def into_step():
            rc = lib.wah_compress_host_into(h_in.data_ptr(), n_words, mode, h_comp.data_ptr(), cap, ctypes.byref(outn))
            assert rc == 0, lib.wah_last_error_string()
            rc = lib.wah_decompress_host_into(h_comp.data_ptr(), outn.value, h_dec.data_ptr(), n_words + 32, ctypes.byref(decn))
            assert rc == 0, lib.wah_last_error_string()
            return outn.value, decn.value

        def timed(fn):
            fn()
            fn()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                c_, n_ = fn()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = t.item()
            return dt, c_, n_

        dt, c_e2e, n_e2e = timed(e2e_step)
        # the reference's three timers (ms): H2D / compute / D2H, one extra untimed step
        fl = [ctypes.c_float() for _ in range(6)]
        lib.wah_compress_host(h_in.data_ptr(), n_words, mode, ctypes.byref(outp), ctypes.byref(outn),
                              ctypes.byref(fl[0]), ctypes.byref(fl[1]), ctypes.byref(fl[2]))
        lib.wah_decompress_host(outp.value, outn.value, ctypes.byref(decp), ctypes.byref(decn),
                                ctypes.byref(fl[3]), ctypes.byref(fl[4]), ctypes.byref(fl[5]))
        lib.wah_free(outp)
        lib.wah_free(decp)
        segments = {"compress_h2d": fl[0].value, "compress_compute": fl[1].value, "compress_d2h": fl[2].value,
                    "decompress_h2d": fl[3].value, "decompress_compute": fl[4].value, "decompress_d2h": fl[5].value}
        dt_into, _, _ = timed(into_step)
        assert torch.equal(h_dec[:n_words], h_in), "host round trip failed"
        e2e = {
            "value": world * 2 * nbytes * args.e2e_steps / dt / 1e9, "unit": UNIT,
            "h2d_bytes_per_step": int(nbytes + 4 * c_e2e), "d2h_bytes_per_step": int(4 * c_e2e + 4 * n_e2e + 24),
            "steps": args.e2e_steps, "ms_per_step": dt / args.e2e_steps * 1e3, "segments_ms": segments,
            "api": "wah_compress_host + wah_decompress_host (= the reference's compress()/decompress(): pinned input, "
                   "malloc()ed results freed by the caller)",
            "caller_buffers": {"value": world * 2 * nbytes * args.e2e_steps / dt_into / 1e9, "unit": UNIT,
                               "ms_per_step": dt_into / args.e2e_steps * 1e3,
                               "api": "wah_compress_host_into + wah_decompress_host_into, page-locked result buffers"},
        }

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    alg_bytes = 4.0 * (n_words + c_avg)
    # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture of this very
    # command (profiles/r1_ncu_full_summary.md); known for the default workload only
    ncu_traffic = {"wah_compress_kernel": 134.29e6 + 6.62e6, "wah_decode_kernel": 10.84e6 + 75.51e6}
    default_wl = args.workload == "sparse_1gbit" and args.density is None and args.mode == "block1024"
    comp_dom = tc_ms >= td_ms
    dom_ms = tc_ms if comp_dom else td_ms
    roofline = {
        "bound": "hbm", "kernel": "wah_compress_kernel" if comp_dom else "wah_decode_kernel",
        "achieved": alg_bytes / (dom_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
        "frac": alg_bytes / (dom_ms * 1e-3) / 1e9 / peak,
        "frac_of_nominal_8tbs": alg_bytes / (dom_ms * 1e-3) / 1e9 / 8000.0,
        "traffic": ncu_traffic["wah_compress_kernel" if comp_dom else "wah_decode_kernel"] if default_wl else None,
        "traffic_note": "DRAM bytes read + written inside the launch (ncu); output still in the 126 MB L2 at kernel end is not in it",
        "peak_source": peak_src,
        "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": dom_ms,
        "launch_ms_note": f"CUDA events around batches of {BATCH} consecutive launches of the kernel, / {BATCH} (launch gaps included)",
        "interleaved_ms": {"compress": tc_alt_ms, "decompress": td_alt_ms,
                           "note": "the same kernels with an event record before and after every single launch, compress and decompress alternating"},
        "compress": {"ms": tc_ms, "achieved": alg_bytes / (tc_ms * 1e-3) / 1e9, "frac": alg_bytes / (tc_ms * 1e-3) / 1e9 / peak,
                     **batch_stats["compress"]},
        "decompress": {"ms": td_ms, "achieved": alg_bytes / (td_ms * 1e-3) / 1e9, "frac": alg_bytes / (td_ms * 1e-3) / 1e9 / peak,
                       **batch_stats["decompress"]},
    }

    cpu_baseline = None
    if not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as orc

        sample_words = min(n_words, 1 << 23)
        data = inputs[0][:sample_words].cpu().numpy().view(np.uint32)
        threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        r = cpu_arm(data, 0 if mode == wah.WAH_BLOCK1024 else 1, 10.0, threads)
        cpu_baseline = {"value": r["value"], "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": f"first {sample_words} words of rank 0's vector, {r['reps']} round trips",
                        "compress_gbs": r["compress_gbs"], "decompress_gbs": r["decompress_gbs"]}

    line = {
        "metric": METRIC, "value": world * 2 * nbytes / (ms_per_step * 1e-3) / 1e9, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": args.workload, "description": desc, "words_per_rank": n_words, "density": density,
                   "mode": args.mode, "compressed_words": c_avg, "ratio": c_avg / n_words,
                   "l2": f"{nbuf} distinct input buffers rotated; a step touches {(2 * nbytes + 8 * c_avg) / 2**20:.0f} MiB (> 126 MB L2)",
                   "step": "compress the vector, then decompress it"},
        "compress_gbs": world * nbytes / (tc_ms * 1e-3) / 1e9, "decompress_gbs": world * nbytes / (td_ms * 1e-3) / 1e9,
        "clocks": clocks, "e2e": e2e, "gpu_launches": 2 * args.steps,
        "gpu_launches_note": "per step: wah_compress_kernel, wah_decode_kernel (scan + expand fused); no memset, no other kernel",
        "roofline": roofline, "cpu_baseline": cpu_baseline,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


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
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
