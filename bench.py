#!/usr/bin/env python
"""bench.py -- the judged measurement of the bit-information hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): one ERA5-like float32 variable of shape
(8760, 721, 1440) (36.4 GB), get_bitinformation along ``longitude`` (the last axis).
One *step* = one pass of the hot path over that array, resident in HBM: workspace clear, the
K1 pair-counting kernel (TMA rows) and ONE finish kernel (leftover pairs, combine, the all-reduce
of the int64[32,2,2] counts over NVLink peer memory when N > 1, K2 mutual information) --
``xbi_bitinformation_dev[_allreduce]``.  ``value`` is GB/s of input for the whole job.  With N > 1
the named array is split over time, 8760/N steps per GPU ("strong": BASELINE configs[1] "sharded
over time to 8"; ``--scaling weak`` gives every rank its own (8760, 721, 1440) slab instead).
The data is generated in 73-step blocks seeded by their position in the global array, so every N
analyses the same array and the printed counts checksum must not depend on N.

The same line carries ``configs.c2 / c3 / c4``: BASELINE configs[2..4] at sizes one GPU holds
(float64 ocean field with a land mask along latitude; one float32 (720,137,256,512) variable along
all four dimensions with the significance filter; float16 / float32 bitinfo -> keepbits -> bitround
on 5e9-element shares), each with GB/s, fraction of the measured copy peak and an in-run bit-exact
check of a slab against oracle/fast_count (outside the timed regions), and ``configs.chunks``: the
per-chunk workflow of the reference's docs/chunking.ipynb (bitinformation, keepbits and bitround of
every block of a chunked variable on its own) on a 3.1 GB month of the same field in 56 blocks.

``--impl reference`` times the reference's CPU algorithm (the numpy restatement in
oracle/, the reference package itself cannot be imported in this image) on a bounded
sample of the same workload with the thread-level parallelism its dask graph has (one
task per byte plane and dask block), on every host core.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "get_bitinformation_input_throughput"
UNIT = "GB/s"
FULL_SHAPE = (8760, 721, 1440)
HBM_NOMINAL_GBS = 8000.0  # north star: "~8 TB/s peak"
HBM_FALLBACK_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"])
    ap.add_argument("--collective", default="auto", choices=["auto", "peer", "nccl"],
                    help="N > 1: all-reduce inside the finish kernel over NVLink peer memory, or NCCL (auto: the faster of a short trial)")
    ap.add_argument("--no-configs", action="store_true", help="skip the configs[2..4] sub-benchmarks")
    ap.add_argument("--config-steps", type=int, default=10)
    ap.add_argument("--time-steps", type=int, default=FULL_SHAPE[0], help="override the length of the time axis (debug)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--ref-budget-s", type=float, default=150.0, help="CPU seconds the whole --impl reference run may take")
    return ap.parse_args()


# --------------------------------------------------------------------------------------
# synthetic data: base + cumulative sum of noise along longitude
# --------------------------------------------------------------------------------------
BLOCK_ROWS = 73  # 8760 = 120 x 73; the shards of 2, 4 and 8 ranks are whole blocks


def synth_rows(lo, hi, device, seed=1234):
    """Time steps [lo, hi) of THE global workload array: block b (steps 73 b .. 73 b + 72) is drawn from a
    generator seeded with seed + b, whoever generates it -- 1 GPU or a shard of 8."""
    import torch

    x = torch.empty((hi - lo,) + FULL_SHAPE[1:], dtype=torch.float32, device=device)
    g = torch.Generator(device=device)
    tmp = None
    for b in range(lo // BLOCK_ROWS, -(-hi // BLOCK_ROWS)):
        b_lo, b_hi = b * BLOCK_ROWS, (b + 1) * BLOCK_ROWS
        g.manual_seed(seed + b)
        whole = lo <= b_lo and b_hi <= hi
        if whole:
            blk = x[b_lo - lo: b_hi - lo]
        else:
            if tmp is None:
                tmp = torch.empty((BLOCK_ROWS,) + FULL_SHAPE[1:], dtype=torch.float32, device=device)
            blk = tmp
        blk.normal_(0.0, 0.35, generator=g)
        torch.cumsum(blk, dim=-1, out=blk)
        blk += 287.0
        if not whole:
            s_lo, s_hi = max(lo, b_lo), min(hi, b_hi)
            x[s_lo - lo: s_hi - lo] = tmp[s_lo - b_lo: s_hi - b_lo]
    return x


def synth_host(shape, seed):
    import numpy as np

    rng = np.random.default_rng(seed)
    return (287.0 + np.cumsum(rng.standard_normal(shape, dtype=np.float32) * np.float32(0.35), axis=-1)).astype(np.float32)


# --------------------------------------------------------------------------------------
# clocks during the timed region
# --------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {
        0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
        0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost",
    }

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                getter = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
                mask = int(getter(self.h))
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.02)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


# --------------------------------------------------------------------------------------
# the reference's CPU algorithm (oracle port) -- only the checker/baseline legs use it
# --------------------------------------------------------------------------------------
def reference_step(x, threads):
    """One pass of the reference python path over ``x`` along the last axis with the parallelism its
    dask graph has: one task per byte plane (xbitinfo/_py_bitinfo.py:128-140) and per dask block of the
    array (``da.array`` cuts a large variable into ~128 MiB blocks, xbitinfo/xbitinfo.py:349; blocks along
    the first axis never split a pair along the last one, and the per-block tables add up)."""
    from concurrent.futures import ThreadPoolExecutor

    import numpy as np

    from oracle import bitinfo_oracle as bo

    X = bo.to_analysis_words(x)
    nbytes = X.dtype.itemsize
    nblocks = max(1, min(X.shape[0], -(-threads // nbytes))) if threads > 1 else 1
    edges = np.linspace(0, X.shape[0], nblocks + 1).astype(int)
    views = []
    for lo, hi in zip(edges[:-1], edges[1:]):
        a, b = bo.adjacent_views(X[lo:hi], x.ndim - 1)
        views.append(np.broadcast_arrays(a, b))

    def task(t):
        blk, i = divmod(t, nbytes)
        a, b = views[blk]
        s = (nbytes - 1 - i) * 8
        return bo._paircount_byteplane((a >> s).astype("u1"), (b >> s).astype("u1"))

    ntasks = nblocks * nbytes
    if threads > 1:
        with ThreadPoolExecutor(max_workers=threads) as ex:
            parts = list(ex.map(task, range(ntasks)))
    else:
        parts = [task(t) for t in range(ntasks)]
    planes = [sum(parts[blk * nbytes + i].astype(np.int64) for blk in range(nblocks)) for i in range(nbytes)]
    counts = np.concatenate(planes, axis=0).astype(np.int64)
    npairs = sum(a.size for a, _ in views)
    return bo.mutual_information_from_counts(counts, npairs)


def fast_cpu_figure(rows=64):
    """For scale only, NOT the reference's algorithm cost: the same counts from oracle/fast_count.c (bit planes by
    64 x 64 transposes + popcounts, every host core) on a (rows,721,1440) slab of the workload."""
    try:
        from oracle import fast_count as fc

        x = synth_host((rows,) + FULL_SHAPE[1:], 1234)
        fc.count_pairs(x[:2], 2)  # loads (or builds) the library
        t0 = time.perf_counter()
        fc.count_pairs(x, 2)
        dt = time.perf_counter() - t0
        return {"value": x.nbytes / dt / 1e9, "unit": UNIT, "cores": os.cpu_count(),
                "note": "oracle/fast_count.c on all host cores; a different (popcount) formulation, not the reference's"}
    except Exception as e:  # never let the extra figure break the line
        return {"error": f"{type(e).__name__}: {e}"}


def cpu_sample_rows(budget_s, per_step_steps, threads):
    """How many time steps of the workload one CPU step may cover so that
    ``per_step_steps`` steps fit in ``budget_s`` (calibrated on one time step)."""
    x = synth_host((1,) + FULL_SHAPE[1:], 99)
    t0 = time.perf_counter()
    reference_step(x, threads)
    t1 = time.perf_counter() - t0
    rows = int(budget_s / max(per_step_steps, 1) / max(t1, 1e-3))
    if threads > 1:
        rows *= max(1, threads // 2)  # the one-row calibration ran four tasks; a real step feeds every core
    return max(1, min(rows, 16 * max(1, threads // 4)))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    threads = cores  # byte planes x dask blocks: every host core gets a task
    steps, warm = args.steps, args.warmup
    rows = cpu_sample_rows(args.ref_budget_s, steps + warm, threads)
    x = synth_host((rows,) + FULL_SHAPE[1:], 1234)
    for _ in range(warm):
        reference_step(x, threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        reference_step(x, threads)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    gbs = x.nbytes / dt / 1e9
    sample = f"({rows},{FULL_SHAPE[1]},{FULL_SHAPE[2]}) float32 time-slab of the workload per step, dim=longitude"
    line = {
        "impl": "reference", "metric": METRIC, "value": gbs, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": gbs, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "host_cores_available": cores,
                         "note": "numpy restatement of xbitinfo/_py_bitinfo.py (oracle/); the reference package needs xarray/dask/numcodecs, absent here"},
        "e2e": {"value": gbs, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    shape = (args.time_steps,) + FULL_SHAPE[1:]
    strong = args.scaling == "strong"
    per_gpu = [-(-shape[0] // world) if strong else shape[0], shape[1], shape[2]]
    return {
        "workload": "configs[1]: ERA5-like float32 (8760,721,1440), get_bitinformation dim='longitude' "
                    "(masked_value=None, set_zero_insignificant=False: the reference python path's semantics)",
        "per_gpu_shape": per_gpu,
        "global_shape": [shape[0] if strong else shape[0] * world, shape[1], shape[2]],
        "sharding": ("time axis, one slab per GPU (the named array split N ways), one all-reduce of int64[32,2,2] counts"
                     if world > 1 else "none"),
        "l2": "input per GPU is far larger than the 126 MB L2 (no flush needed)",
    }


def _time_steps(fn, steps, sync_all, dist_max):
    """ms per call of fn over `steps` calls: CUDA events on the current stream, barrier + synchronize on both
    sides, max over ranks."""
    import torch

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    sync_all()
    return dist_max(e0.elapsed_time(e1)) / steps


def _checksum(counts_np):
    import hashlib

    return hashlib.sha256(counts_np.astype("<i8").tobytes()).hexdigest()[:16]


# --------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    # one staging-thread pool per rank, sized to its share of the host cores (read when the library starts its pool)
    os.environ.setdefault("XBI_HOST_THREADS", str(max(1, min(8, (os.cpu_count() or 8) // max(world, 1)))))

    from xbitinfo_b200 import _native as nat
    from xbitinfo_b200 import distributed as xdist
    from xbitinfo_b200.device import PairCounter

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_node = None
    if world > 1:
        # one process per GPU: keep each rank's host staging memory on its GPU's NUMA node
        numa_node = xdist.bind_to_gpu_numa_node(local_rank)
        dist.init_process_group("nccl", device_id=dev)

    def dist_max(v):
        if world == 1:
            return float(v)
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    T = args.time_steps
    if args.scaling == "strong":
        lo, hi = xdist.shard_bounds(T, world, rank)
        global_rows = T
    else:
        lo, hi = rank * T, (rank + 1) * T
        global_rows = T * world
    x = synth_rows(lo, hi, dev)
    nbytes_local = x.numel() * 4
    total_bytes = global_rows * FULL_SHAPE[1] * FULL_SHAPE[2] * 4
    pc = PairCounter(x.dtype, dev)
    sync_all = (lambda: (dist.barrier(), torch.cuda.synchronize())) if world > 1 else torch.cuda.synchronize

    # ---- the step: K1 + finish kernel (+ the all-reduce inside it) ----
    peers, peer_error = None, None
    if world > 1 and args.collective in ("auto", "peer"):
        try:
            peers = xdist.PeerExchange(dev)
        except Exception as e:  # IPC not permitted in this container, ...: NCCL takes over, and the line says so
            peer_error = f"{type(e).__name__}: {e}"

    def step_local():
        return pc.bitinformation(x, -1, masked_value=None)

    def step_peer():
        return pc.bitinformation(x, -1, masked_value=None, peers=peers)

    def step_nccl():
        pc.bitinformation(x, -1, masked_value=None, with_information=False)  # K1 -> per-rank table
        xdist.allreduce_counts(pc.counts)                                     # the path's one collective
        return pc.information(False)                                          # K2

    def trial(fn, steps=15):
        """ms per step (max over ranks) and the per-step time outside the counting kernel (events around that kernel on
        its own stream; max over ranks) -- the second is what the collectives differ in, the first drifts with clocks."""
        for _ in range(max(args.warmup, 3)):
            fn()
        sync_all()
        nat.profile_enable(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        sync_all()
        el = e0.elapsed_time(e1)
        kms, kn, _ = nat.profile_collect()
        nat.profile_enable(False)
        return {"ms_per_step": dist_max(el) / steps, "step_minus_kernel_us": dist_max(1e3 * (el - kms) / steps)}

    collective = {"kind": "none"}
    step = step_local
    if world > 1:
        trials = {"nccl": trial(step_nccl)}
        if peers is not None:
            trials["peer"] = trial(step_peer)
        use_peer = peers is not None and (
            args.collective == "peer" or trials["peer"]["step_minus_kernel_us"] <= trials["nccl"]["step_minus_kernel_us"])
        step = step_peer if use_peer else step_nccl
        collective = {
            "kind": ("push over NVLink peer memory inside the finish kernel (xbi_bitinformation_dev_allreduce)" if use_peer
                     else "torch.distributed.all_reduce (NCCL) on int64[32,2,2] + separate K2 launch"),
            "trial_15_steps": trials, "peer_unavailable": peer_error,
        }

    for _ in range(max(args.warmup, 3)):
        step()
    sync_all()
    launches0 = nat.kernel_launches()
    nat.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        sync_all()
        e0.record()
        for _ in range(args.steps):
            mi_dev = step()
        e1.record()
        sync_all()
    elapsed_local = e0.elapsed_time(e1)
    kern_ms, kern_n, kern_bytes = nat.profile_collect()
    nat.profile_enable(False)
    launches = nat.kernel_launches() - launches0
    elapsed_ms = dist_max(elapsed_local)
    ms_per_step = elapsed_ms / args.steps
    value = total_bytes / (ms_per_step * 1e-3) / 1e9
    mi = mi_dev.cpu().numpy()
    counts_np = pc.counts.cpu().numpy()

    # ---- in-run verification of what was timed (outside the timed region) ----
    npairs_global = global_rows * FULL_SHAPE[1] * (FULL_SHAPE[2] - 1)
    verify = {
        "pairs_expected": int(npairs_global),
        "cells_sum_to_pairs": bool((counts_np.reshape(32, 4).sum(axis=1) == npairs_global).all()),
        "counts_checksum": _checksum(counts_np),  # the same for every N (strong scaling analyses one array)
    }
    if world > 1:
        # the all-reduced table must equal the sum of the per-rank tables, summed here once more with NCCL
        ref = PairCounter(x.dtype, dev)
        ref.bitinformation(x, -1, masked_value=None, with_information=False)
        local_sum_ok = bool((ref.counts.cpu().numpy().reshape(32, 4).sum(axis=1) == (hi - lo) * FULL_SHAPE[1] * (FULL_SHAPE[2] - 1)).all())
        dist.all_reduce(ref.counts, op=dist.ReduceOp.SUM)
        same = torch.tensor([int(torch.equal(ref.counts, pc.counts)) & int(local_sum_ok)], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        verify["allreduce_equals_nccl_sum_on_every_rank"] = bool(same.item())
        if peers is not None:
            failed = torch.tensor([peers.failed_epoch()], device=dev)
            dist.all_reduce(failed, op=dist.ReduceOp.MAX)
            verify["peer_wait_timeouts"] = int(failed.item() != 0)
        del ref
    if not verify["cells_sum_to_pairs"] or verify.get("allreduce_equals_nccl_sum_on_every_rank") is False:
        raise SystemExit(f"bench.py: counts verification failed: {verify}")

    # ---- roofline of the dominant kernel (TMA row kernel), timed with events on its own stream ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", HBM_FALLBACK_GBS))
    achieved = (kern_bytes / kern_n) / (kern_ms / kern_n * 1e-3) / 1e9 if kern_n else None
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": (achieved / peak) if achieved else None,
        "peak_source": "MEASURED_PEAKS.json hbm_gbs (copy, of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (of fallback)",
        "frac_of_nominal_8TBs": (achieved / HBM_NOMINAL_GBS) if achieved else None,
        "kernel": "k_pairs_rows_tma<f32,unmasked>", "launches_timed": kern_n,
        "algorithmic_bytes_per_launch": (kern_bytes / kern_n) if kern_n else None,
        "kernel_ms_avg": (kern_ms / kern_n) if kern_n else None,
        "kernel_share_of_step": (kern_ms / elapsed_local) if kern_n else None,
        "step_minus_kernel_us": (1e3 * (elapsed_local - kern_ms) / args.steps) if kern_n else None,  # this rank
        "traffic": None,  # dram__bytes per launch from ncu --set full: see profiles/
    }
    traffic_file = os.path.join(ROOT, "profiles", "k1_rows_f32_dram_bytes.json")
    if os.path.exists(traffic_file):
        try:
            tr = json.load(open(traffic_file))
            roofline["traffic"] = tr.get("dram_bytes_per_input_byte", 0) * roofline["algorithmic_bytes_per_launch"]
            roofline["traffic_source"] = tr.get("source")
        except Exception:
            pass

    # ---- the same step replayed from a CUDA graph (launch gaps out of the picture; informational, 1 GPU) ----
    graph_ms = None
    if world == 1:
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                step()
            for _ in range(3):
                g.replay()
            graph_ms = _time_steps(g.replay, args.steps, sync_all, dist_max)
            del g
        except Exception as e:
            graph_ms = f"unavailable: {type(e).__name__}: {e}"

    # ---- K3 on the same array (second half of the headline metric) ----
    bitround_gbs = None
    if rank == 0:
        from xbitinfo_b200.device import bitround_

        y = x[: min(x.shape[0], 2048)].clone()
        for _ in range(3):
            bitround_(y, 7)
        bitround_gbs = y.numel() * 4 / (_time_steps(lambda: bitround_(y, 7), 10, torch.cuda.synchronize, float) * 1e-3) / 1e9
        del y

    # the same step with the public API's default keywords (masked_value=NaN: pairs holding a NaN are dropped;
    # the data has none, the probe in the kernel has to find that out; significance filter on)
    nan_masked_gbs = None
    if rank == 0:
        def masked_step():
            return pc.bitinformation(x, -1, masked_value=float("nan"), set_zero_insignificant=True, confidence=0.99)

        for _ in range(3):
            masked_step()
        nan_masked_gbs = nbytes_local / (_time_steps(masked_step, 10, torch.cuda.synchronize, float) * 1e-3) / 1e9

    # ---- end to end through the public API with host buffers (H2D inside the timed region) ----
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, x, rank, world, dev, sync_all, dist_max)
        e2e["numa_node_of_rank0"] = numa_node
        e2e["host_copy_threads_per_rank"] = int(os.environ["XBI_HOST_THREADS"])
    del x, pc
    torch.cuda.empty_cache()

    # ---- BASELINE configs[2..4] at sizes one GPU holds; N > 1: configs[4], the sharded pipeline ----
    configs = None
    if not args.no_configs:
        if world == 1:
            configs = run_configs(args, dev, peak)
        else:
            configs = {"c4": run_pipeline_sharded(args, dev, peak, rank, world, peers, sync_all, dist_max)}
    if peers is not None:
        peers.close()

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rows = cpu_sample_rows(20.0, 1, 1)
        xs = synth_host((rows,) + FULL_SHAPE[1:], 1234)
        t0 = time.perf_counter()
        reference_step(xs, 1)
        dt = time.perf_counter() - t0
        cpu_baseline = {
            "value": xs.nbytes / dt / 1e9, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"({rows},{FULL_SHAPE[1]},{FULL_SHAPE[2]}) float32 time-slab of the workload, one pass, dim=longitude",
            "host_cores_available": os.cpu_count(),
        }
        cpu_baseline["fast_cpu_popcount"] = fast_cpu_figure()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": workload_config(args, world),
            "pct_hbm_peak_nominal": 100.0 * value / (HBM_NOMINAL_GBS * world),
            "pct_hbm_peak_measured": 100.0 * value / (peak * world),
            "launches_per_step": launches / args.steps,
            "collective": collective, "verify": verify,
            "graph_replay_ms_per_step": graph_ms,
            "xr_bitround_input_GBps": bitround_gbs,
            "default_kwargs_step_GBps": nan_masked_gbs,  # masked_value=NaN, set_zero_insignificant=True (1 GPU)
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "configs": configs,
            "gpu_launches": int(launches), "clocks": clocks.summary(),
            "info_bits_preview": [float(v) for v in mi[:12]],
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_e2e(args, x_dev, rank, world, dev, sync_all, dist_max):
    """Same metric through the public host-buffer API: this rank's slab as a numpy array in pinned host memory ->
    sharded_bitinformation (xbi_bitinformation_host: slabs H2D + K1, then all-reduce + K2) -> 32 doubles back.
    Also: the same call on a pageable numpy array (what a numpy / xarray user passes), every rank's own H2D rate,
    and -- N > 1 -- a plain pinned -> device copy of 1 GiB per rank, each rank alone and all ranks together, to tell
    the host's share of an e2e number that does not scale from the library's."""
    import numpy as np
    import torch
    import torch.distributed as dist

    from xbitinfo_b200 import distributed as xdist

    try:
        import psutil

        avail = psutil.virtual_memory().available
    except Exception:
        avail = 64 << 30
    full = x_dev.numel() * 4
    budget = int(avail * 0.35 / max(world if args.scaling == "weak" else 1, 1))
    rows = x_dev.shape[0] if full <= budget else max(1, int(x_dev.shape[0] * budget / full))
    rows = min(rows, x_dev.shape[0])
    try:
        host = torch.empty((rows,) + tuple(x_dev.shape[1:]), dtype=torch.float32, pin_memory=True)
        pinned = True
    except Exception:
        host = torch.empty((rows,) + tuple(x_dev.shape[1:]), dtype=torch.float32)
        pinned = False
    host.copy_(x_dev[:rows])
    torch.cuda.synchronize()
    xh = host.numpy()
    steps = max(1, min(args.e2e_steps, args.steps))
    xdist.sharded_bitinformation(xh, -1, masked_value=None, device=dev.index)  # warm (staging buffers)
    sync_all()
    own = 0.0
    t0 = time.perf_counter()
    for _ in range(steps):
        t1 = time.perf_counter()
        xdist.sharded_bitinformation(xh, -1, masked_value=None, device=dev.index)
        own += time.perf_counter() - t1
    torch.cuda.synchronize()
    dt = dist_max((time.perf_counter() - t0) / steps)
    out = {
        "value": xh.nbytes * world / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(xh.nbytes),
        "d2h_bytes_per_step": int(32 * 8 + 32 * 4 * 8), "steps": steps, "pinned": pinned,
        "sample": f"({rows},{x_dev.shape[1]},{x_dev.shape[2]}) float32 per GPU"
                  + ("" if rows == x_dev.shape[0] else " (host RAM bound: time-slab of the workload)"),
        "api": "xbitinfo_b200.distributed.sharded_bitinformation(numpy) -> xbi_bitinformation_host (C ABI)",
    }
    rate = xh.nbytes * steps / own / 1e9
    if world > 1:
        rates = [None] * world
        dist.all_gather_object(rates, rate)
        out["h2d_GBps_per_rank"] = [round(float(r), 2) for r in rates]
    else:
        out["h2d_GBps_per_rank"] = [round(rate, 2)]

    # pageable input: an ordinary numpy array (up to 4 GiB of the same slab), through the pinned ring + copy threads
    prow = max(1, min(rows, (4 << 30) // (x_dev.shape[1] * x_dev.shape[2] * 4)))
    xp = np.empty((prow,) + tuple(x_dev.shape[1:]), dtype=np.float32)
    np.copyto(xp, xh[:prow])
    xdist.sharded_bitinformation(xp, -1, masked_value=None, device=dev.index)
    sync_all()
    t0 = time.perf_counter()
    xdist.sharded_bitinformation(xp, -1, masked_value=None, device=dev.index)
    torch.cuda.synchronize()
    dtp = dist_max(time.perf_counter() - t0)
    out["pageable"] = {"value": xp.nbytes * world / dtp / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(xp.nbytes),
                       "sample": f"({prow},{x_dev.shape[1]},{x_dev.shape[2]}) float32 per GPU, ordinary (pageable) numpy memory"}
    del xp

    if rank == 0 or world > 1:
        # the table that came back through the host path must be the one the device path printed
        _, c_e2e = xdist.sharded_bitinformation(xh, -1, masked_value=None, device=dev.index, return_counts=True)
        if rows == x_dev.shape[0]:
            out["counts_checksum"] = _checksum(c_e2e)

    if world > 1:
        # where does a flat e2e curve come from?  1 GiB pinned -> device per rank, alone and together
        n = min(host.numel(), (1 << 30) // 4)
        src, dst = host.view(-1)[:n], torch.empty(n, dtype=torch.float32, device=dev)

        def copy_rate():
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            return 3 * n * 4 / (time.perf_counter() - t0) / 1e9

        copy_rate()
        alone = 0.0
        for r in range(world):
            dist.barrier()
            if r == rank:
                alone = copy_rate()
        dist.barrier()
        together = copy_rate()
        a, t = [None] * world, [None] * world
        dist.all_gather_object(a, alone)
        dist.all_gather_object(t, together)
        out["h2d_probe_1GiB_pinned"] = {"alone_GBps": [round(float(v), 1) for v in a],
                                        "together_GBps": [round(float(v), 1) for v in t],
                                        "together_sum_GBps": round(float(sum(t)), 1)}
        del src, dst
        spread = max(t) / max(min(t), 1e-9)
        if os.environ.get("XBI_BENCH_FORCE_WEIGHTED") == "1":  # (exercise the weighted path on a box with equal links)
            t = [v * (1.0 + 0.5 * (i % 2)) for i, v in enumerate(t)]
            spread = 1.5
        if args.scaling == "strong" and spread > 1.15 and rows == x_dev.shape[0]:
            # The GPUs of this box do not all get the same share of the host's PCIe / memory fabric when they copy
            # together.  Equal time slabs then finish with the slowest link; slabs PROPORTIONAL to each rank's measured
            # rate finish together.  (Any partition of the time axis gives the same table: the checksum says so.)
            del host, xh
            T = args.time_steps
            w = np.asarray(t, dtype=np.float64)
            edges = np.concatenate([[0], np.rint(np.cumsum(w) / w.sum() * T)]).astype(np.int64)
            edges[-1] = T
            lo_w, hi_w = int(edges[rank]), int(edges[rank + 1])
            hostw = torch.empty((hi_w - lo_w,) + tuple(x_dev.shape[1:]), dtype=torch.float32, pin_memory=pinned)
            for b_lo in range(lo_w, hi_w, 4 * BLOCK_ROWS):
                b_hi = min(hi_w, b_lo + 4 * BLOCK_ROWS)
                hostw[b_lo - lo_w: b_hi - lo_w].copy_(synth_rows(b_lo, b_hi, dev))
            torch.cuda.synchronize()
            xw = hostw.numpy()
            xdist.sharded_bitinformation(xw, -1, masked_value=None, device=dev.index)
            sync_all()
            t0 = time.perf_counter()
            for _ in range(steps):
                _, c_w = xdist.sharded_bitinformation(xw, -1, masked_value=None, device=dev.index, return_counts=True)
            torch.cuda.synchronize()
            dtw = dist_max((time.perf_counter() - t0) / steps)
            total = T * x_dev.shape[1] * x_dev.shape[2] * 4
            weighted = {"value": total / dtw / 1e9, "unit": UNIT, "rows_per_rank": [int(edges[i + 1] - edges[i]) for i in range(world)],
                        "counts_checksum": _checksum(c_w),
                        "note": "time slabs proportional to each rank's measured H2D rate with all ranks copying"}
            out["equal_shards"] = {"value": out["value"], "h2d_bytes_per_step": out["h2d_bytes_per_step"]}
            out["weighted_shards"] = weighted
            if weighted["value"] > out["value"]:
                out["value"] = weighted["value"]
                out["h2d_bytes_per_step"] = int(xw.nbytes)  # this rank's share
                out["sample"] = f"({hi_w - lo_w},{x_dev.shape[1]},{x_dev.shape[2]}) float32 on rank {rank}; time slabs weighted by link rate"
                out["counts_checksum"] = weighted["counts_checksum"]
    return out


# --------------------------------------------------------------------------------------
# BASELINE configs[2], [3], [4] on one GPU
# --------------------------------------------------------------------------------------
def _synth_nd(shape, dtype, axis, seed, device, piece_bytes=1 << 30):
    """base + cumulative sum of float32 noise along `axis` (not the first one), generated in pieces along the
    first axis, cast to `dtype`."""
    import torch

    g = torch.Generator(device=device).manual_seed(seed)
    x = torch.empty(shape, dtype=dtype, device=device)
    per = max(1, piece_bytes // max(1, x[0].numel() * 4))
    for lo in range(0, shape[0], per):
        k = min(per, shape[0] - lo)
        blk = torch.empty((k,) + tuple(shape[1:]), dtype=torch.float32, device=device)
        blk.normal_(0.0, 0.3, generator=g)
        torch.cumsum(blk, dim=axis, out=blk)
        blk += 12.0
        x[lo: lo + k] = blk.to(dtype)
    return x


def run_configs(args, dev, peak):
    import numpy as np
    import torch

    from oracle import fast_count as fc  # the checker, outside every timed region
    from xbitinfo_b200 import _native as nat
    from xbitinfo_b200.chunked import keepbits_of_information
    from xbitinfo_b200.device import PairCounter, bitround

    K = max(3, args.config_steps)
    sync = torch.cuda.synchronize
    out = {}

    def timed(fn):
        for _ in range(3):
            fn()
        return _time_steps(fn, K, sync, float)

    # ---- c2: float64 ocean field, land mask, dim='latitude' (strided adjacency) ----
    try:
        days = 24
        shape = (days, 75, 1080, 1440)
        x = _synth_nd(shape, torch.float64, 2, 77, dev, piece_bytes=1 << 29)
        land = torch.rand((1080, 1440), generator=torch.Generator(device=dev).manual_seed(5), device=dev)
        # ~30 % land as a few continents: white noise box-filtered twice at 255 points and thresholded (a handful of
        # coast-line crossings per column, like a real 1/4-degree ocean grid; not salt and pepper)
        for _ in range(2):
            land = torch.nn.functional.avg_pool2d(land[None, None], 255, stride=1, padding=127)[0, 0]
        land = land < torch.quantile(land.flatten()[::7], 0.30)
        res = {}
        pc0 = PairCounter(x.dtype, dev)
        ms0 = timed(lambda: pc0.bitinformation(x, 2, masked_value=None))  # the same field before the land goes in
        no_mask_gbs = x.numel() * 8 / (ms0 * 1e-3) / 1e9
        del pc0
        for name, mv in (("nan", float("nan")), ("minus999", -999.0)):
            x.masked_fill_(land, mv)
            pc = PairCounter(x.dtype, dev)
            nat.profile_enable(True)
            ms = timed(lambda: pc.bitinformation(x, 2, masked_value=mv, set_zero_insignificant=True, confidence=0.99))
            kms, kn, kb = nat.profile_collect()
            nat.profile_enable(False)
            day = x[:1].contiguous()
            pd = PairCounter(x.dtype, dev)
            pd.bitinformation(day, 2, masked_value=mv)
            exact = bool(np.array_equal(pd.counts.cpu().numpy(), fc.count_pairs(day.cpu().numpy(), 2, masked_value=mv)))
            gbs = x.numel() * 8 / (ms * 1e-3) / 1e9
            res[name] = {"value": gbs, "unit": UNIT, "ms_per_step": ms, "frac_of_measured_peak": gbs / peak,
                         "kernel_GBps": (kb / (kms * 1e-3) / 1e9) if kn else None,
                         "bit_exact_vs_fast_count": exact, "checked": "first day (75,1080,1440)"}
            del pc, pd, day
        out["c2"] = {
            "workload": f"configs[2]: float64 ocean field, {days} of 365 days resident ({days},75,1080,1440) = "
                        f"{x.numel() * 8 / 1e9:.1f} GB (the full 340 GB field streams through the same accumulate-only "
                        "counts in time slabs), ~30 % land mask constant over time/level, dim='latitude' (inner = 1440), "
                        "masked_value + set_zero_insignificant=True, confidence=0.99",
            "land_fraction": float(land.float().mean()),
            "coast_crossings_per_column": float((land[1:] != land[:-1]).float().sum(0).mean()), "kernel": "k_pairs_cols<f64, masked sparse flavour>",
            "masked_value_nan": res["nan"], "masked_value_minus999": res["minus999"],
            "same_field_without_mask_GBps": no_mask_gbs,
        }
        del x, land
    except Exception as e:
        out["c2"] = {"error": f"{type(e).__name__}: {e}"}
    torch.cuda.empty_cache()

    # ---- c3: one float32 (720,137,256,512) variable along all four dims, significance filter on ----
    try:
        shape = (720, 137, 256, 512)
        x = _synth_nd(shape, torch.float32, 3, 91, dev)
        from xbitinfo_b200.device import MultiAxisCounter

        kw = dict(masked_value=float("nan"), set_zero_insignificant=True, confidence=0.99)
        mc = MultiAxisCounter(x.dtype, 4, dev)
        ms = timed(lambda: mc.bitinformation(x, [0, 1, 2, 3], **kw))       # xbi_bitinformation_dev_multi
        pc = PairCounter(x.dtype, dev)
        per_axis = []
        for a in range(4):
            m = timed(lambda a=a: pc.bitinformation(x, a, **kw))            # one single-axis pass, for scale
            per_axis.append({"axis": a, "ms": m, "GBps": x.numel() * 4 / (m * 1e-3) / 1e9})
        m_xy = timed(lambda: mc2.bitinformation(x, [2, 3], **kw)) if (mc2 := MultiAxisCounter(x.dtype, 2, dev)) else None
        slab = x[:8].contiguous()
        hs = slab.cpu().numpy()
        ms_slab = MultiAxisCounter(x.dtype, 4, dev)
        ms_slab.bitinformation(slab, [0, 1, 2, 3], masked_value=float("nan"))
        got = ms_slab.counts.cpu().numpy()
        exact = all(bool(np.array_equal(got[a], fc.count_pairs(hs, a, masked_value=float("nan")))) for a in range(4))
        nb = x.numel() * 4
        out["c3"] = {
            "workload": "configs[3]: one of the 10 float32 variables, (720,137,256,512) = 51.7 GB resident, analysed along "
                        "every dim in one call (get_bitinformation(ds) with dim=None -> xbi_bitinformation_dev_multi: the two "
                        "innermost dims in one read, the outer two one pass each), masked_value=NaN (API default), "
                        "set_zero_insignificant=True, confidence=0.99",
            "ms_all_dims": ms, "array_GBps": nb / (ms * 1e-3) / 1e9,       # the array's bytes once per all-dims job
            "value": 4 * nb / (ms * 1e-3) / 1e9, "unit": UNIT,              # (variable, dim) passes x bytes, SURVEY 8(d)
            "frac_of_measured_peak": 4 * nb / (ms * 1e-3) / 1e9 / peak,
            "passes_equivalent": ms / min(p_["ms"] for p_ in per_axis),  # in units of the fastest single-axis pass
            "ms_one_pass_per_axis": sum(p_["ms"] for p_ in per_axis),
            "two_innermost_axes_in_one_read": {"ms": m_xy, "array_GBps": x.numel() * 4 / (m_xy * 1e-3) / 1e9,
                                               "kernel": "k_pairs_cols_xy<f32>"},
            "per_axis": per_axis, "bit_exact_vs_fast_count": exact, "checked": "first 8 time steps, all four axes",
        }
        del x, pc, mc, mc2, ms_slab, slab
    except Exception as e:
        out["c3"] = {"error": f"{type(e).__name__}: {e}"}
    torch.cuda.empty_cache()

    # ---- c4: bitinfo -> keepbits -> bitround on one GPU's share (5e9 elements), float16 and float32 ----
    try:
        res = {}
        n = 5_000_000_000
        for name, tdt, ndt in (("float16", torch.float16, "float16"), ("float32", torch.float32, "float32")):
            rows = n // 100_000
            x = _synth_nd((rows, 100_000), tdt, 1, 13, dev).view(-1)
            y = torch.empty_like(x)
            pc = PairCounter(tdt, dev)
            state = {}

            def pipeline():
                mi = pc.bitinformation(x, 0, masked_value=float("nan"), set_zero_insignificant=True, confidence=0.99)
                info = mi.cpu().numpy()                                   # 16 / 32 doubles back: the host decides
                keep = max(0, keepbits_of_information(info, np.dtype(ndt), 0.99))
                bitround(x, keep, out=y)                                  # new array, input untouched
                state["keep"] = keep

            ms = timed(pipeline)
            m_count = timed(lambda: pc.bitinformation(x, 0, masked_value=float("nan"), set_zero_insignificant=True, confidence=0.99))
            m_round = timed(lambda: bitround(x, state["keep"], out=y))
            # checks on a 2^24-element slab: counts vs fast_count, rounded bits vs the bitround oracle
            from oracle import bitround_oracle as bro

            k = 1 << 24
            hs = x[:k].cpu().numpy()
            p = PairCounter(tdt, dev)
            p.bitinformation(x[:k], 0, masked_value=float("nan"))
            u = f"u{np.dtype(ndt).itemsize}"
            exact = bool(np.array_equal(p.counts.cpu().numpy(), fc.count_pairs(hs, 0, masked_value=float("nan"))))
            exact_round = bool(np.array_equal(y[:k].cpu().numpy().view(u), bro.bitround(hs, state["keep"]).view(u)))
            isz = x.element_size()
            res[name] = {
                "ms_pipeline": ms, "elements_per_s": n / (ms * 1e-3), "keepbits": state["keep"],
                "value": 3 * n * isz / (ms * 1e-3) / 1e9, "unit": UNIT,  # 1x read (count) + read + write (round), SURVEY 8(d)
                "frac_of_measured_peak": 3 * n * isz / (ms * 1e-3) / 1e9 / peak,
                "count_GBps": n * isz / (m_count * 1e-3) / 1e9, "bitround_GBps_read_plus_write": 2 * n * isz / (m_round * 1e-3) / 1e9,
                "bit_exact_counts_vs_fast_count": exact, "bit_exact_rounding_vs_oracle": exact_round,
                "checked": "first 2^24 elements",
            }
            del x, y, pc, p
            torch.cuda.empty_cache()
        out["c4"] = {
            "workload": "configs[4]: bitinfo -> get_keepbits(0.99) -> bitround on ONE GPU's share of the 4e10-element arrays "
                        "(5e9 elements: float16 10 GB, float32 20 GB), API default keywords; the 8-GPU job runs this on "
                        "every rank with the all-reduce of the counts before the keepbits decision",
            "float16": res["float16"], "float32": res["float32"],
        }
    except Exception as e:
        out["c4"] = {"error": f"{type(e).__name__}: {e}"}
    torch.cuda.empty_cache()

    # ---- chunks (docs/chunking.ipynb cells 8-10): get_bitinformation on every block of a chunked variable
    # on its own -> keepbits per block -> bitround per block, one month of the configs[1] field cut into 7 x 8 blocks ----
    try:
        from xbitinfo_b200.chunked import keepbits_of_information_batched
        from xbitinfo_b200.device import bitround_chunked_, count_pairs_chunked, information_batched

        shape, chunk = (744, 721, 1440), (744, 103, 180)
        x = _synth_nd(shape, torch.float32, 2, 21, dev)
        nb = x.numel() * x.element_size()
        res = {}
        for axis, name in ((0, "dim_time"), (2, "dim_longitude")):
            ms = timed(lambda: count_pairs_chunked(x, chunk, axis, masked_value=float("nan")))
            res[name] = {"ms": ms, "value": nb / (ms * 1e-3) / 1e9, "unit": UNIT, "frac_of_measured_peak": nb / (ms * 1e-3) / 1e9 / peak}
        y = x.clone()
        state = {}

        def workflow():
            counts = count_pairs_chunked(x, chunk, 0, masked_value=float("nan"))
            info = information_batched(counts, True, 0.99).cpu().numpy().reshape(-1, 32)
            keep = np.clip(keepbits_of_information_batched(info, np.dtype("float32"), 0.99), 0, 23).astype("int32")
            y.copy_(x)
            bitround_chunked_(y, chunk, torch.from_numpy(keep).to(dev))
            state["counts"], state["keep"] = counts, keep

        ms_flow = timed(workflow)
        # check: three blocks (first, one in the middle, the ragged last one) against fast_count on the host
        got = state["counts"].cpu().numpy()
        exact = True
        for (cy, cx) in ((0, 0), (3, 4), (6, 7)):
            blk = x[:, cy * 103:(cy + 1) * 103, cx * 180:(cx + 1) * 180].contiguous().cpu().numpy()
            exact = exact and bool(np.array_equal(got[0, cy, cx], fc.count_pairs(blk, 0, masked_value=float("nan"))))
        out["chunks"] = {
            "workload": "per-chunk workflow of docs/chunking.ipynb on the device: float32 (744,721,1440) = 3.1 GB in 56 blocks of "
                        "(744,103,180), masked_value=NaN; xbi_count_pairs_chunked alone along time / longitude, then count -> "
                        "information -> keepbits per block (host) -> copy + bitround per block",
            "count_pairs_chunked": res,
            "workflow_ms": ms_flow, "workflow_bytes_GBps": 5 * nb / (ms_flow * 1e-3) / 1e9,  # count: 1 read; copy: read + write; round in place: read + write
            "keepbits_min_max": [int(state["keep"].min()), int(state["keep"].max())],
            "bit_exact_vs_fast_count": exact, "checked": "blocks (0,0), (3,4), (6,7) along time",
        }
        del x, y, state
    except Exception as e:
        out["chunks"] = {"error": f"{type(e).__name__}: {e}"}
    torch.cuda.empty_cache()
    return out


def run_pipeline_sharded(args, dev, peak, rank, world, peers, sync_all, dist_max):
    """configs[4] across the GPUs of the node: every rank holds 5e9 elements of the (world x 5e9) array; one step =
    count the local share -> all-reduce the int64[nbits,2,2] counts (inside the finish kernel when the peer exchange
    is up, else NCCL) -> K2 on every rank -> 16 / 32 doubles to the host -> get_keepbits(0.99) -> bitround the
    local share into a second array.  Timed with CUDA events, barrier on both sides, max over ranks."""
    import numpy as np
    import torch
    import torch.distributed as dist

    from oracle import bitround_oracle as bro
    from oracle import fast_count as fc
    from xbitinfo_b200 import distributed as xdist
    from xbitinfo_b200.chunked import keepbits_of_information
    from xbitinfo_b200.device import PairCounter, bitround

    K = max(3, args.config_steps)
    n = 5_000_000_000
    res = {}
    try:
        for name, tdt in (("float16", torch.float16), ("float32", torch.float32)):
            x = _synth_nd((n // 100_000, 100_000), tdt, 1, 13 + rank, dev).view(-1)
            y = torch.empty_like(x)
            pc = PairCounter(tdt, dev)
            kw = dict(masked_value=float("nan"), set_zero_insignificant=True, confidence=0.99)
            state = {}

            def pipeline():
                if peers is not None:
                    mi = pc.bitinformation(x, 0, peers=peers, **kw)
                else:
                    pc.bitinformation(x, 0, masked_value=float("nan"), with_information=False)
                    xdist.allreduce_counts(pc.counts)
                    mi = pc.information(True, 0.99)
                keep = max(0, keepbits_of_information(mi.cpu().numpy(), np.dtype(name), 0.99))  # the same on every rank
                bitround(x, keep, out=y)
                state["keep"] = keep

            for _ in range(3):
                pipeline()
            ms = _time_steps(pipeline, K, sync_all, dist_max)
            # checks (outside the timed region): the global table = sum of per-rank tables of fast_count on a slab is not
            # affordable at 5e9 elements per rank; the local slab's counts and the rounded bits are checked instead, and
            # every rank must have taken the same keepbits decision
            k = 1 << 24
            hs = x[:k].cpu().numpy()
            p = PairCounter(tdt, dev)
            p.bitinformation(x[:k], 0, masked_value=float("nan"))
            u = f"u{np.dtype(name).itemsize}"
            ok = np.array_equal(p.counts.cpu().numpy(), fc.count_pairs(hs, 0, masked_value=float("nan"))) and \
                np.array_equal(y[:k].cpu().numpy().view(u), bro.bitround(hs, state["keep"]).view(u))
            pairs_ok = int(pc.counts[0].sum().item()) == world * (n - 1)
            flags = torch.tensor([int(ok and pairs_ok), state["keep"], -state["keep"]], device=dev)
            dist.all_reduce(flags, op=dist.ReduceOp.MIN)
            isz = x.element_size()
            res[name] = {
                "ms_pipeline": ms, "elements_per_s": world * n / (ms * 1e-3), "keepbits": state["keep"],
                "value": 3 * world * n * isz / (ms * 1e-3) / 1e9, "unit": UNIT,
                "frac_of_measured_peak": 3 * n * isz / (ms * 1e-3) / 1e9 / peak,
                "bit_exact_slab_on_every_rank": bool(flags[0].item()),
                "same_keepbits_on_every_rank": bool(flags[1].item() == -flags[2].item()),
                "global_pairs_ok": pairs_ok,
            }
            del x, y, pc, p
            torch.cuda.empty_cache()
        res["workload"] = (f"configs[4]: bitinfo -> all-reduce -> get_keepbits(0.99) -> bitround on a ({world} x 5e9)-element "
                           f"array, one 5e9-element share per GPU ({world * 5e9:.1e} elements in all), float16 and float32, "
                           "API default keywords")
        res["collective"] = "peer exchange in the finish kernel" if peers is not None else "NCCL all_reduce"
    except Exception as e:
        res["error"] = f"{type(e).__name__}: {e}"
    return res


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
