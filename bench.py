#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native stroke tessellator / flattener.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W   (the CPU arm: oracle restatement of lyon)

Workload (BASELINE.json configs[2], "C3"): 1M synthetic paths x 64 mixed line/quad/cubic segments, stroke width 2,
round joins + round caps, tolerance 0.1.  STRONG scaling: the 1M-path batch is partitioned across the N GPUs by
cumulative segment count; the only collective is an all-gather of per-rank vertex/index totals (global index
rebasing: every rank emits with local ids in one pass, then adds its base to its indices).  A "step" = pre-pass +
fused count/scan/emit kernel (+ all-gather + index rebase at N>1) over the rank's whole slice, inputs already resident in
HBM.  `value` = output vertices of all ranks / max-over-ranks device time.

Also reported on the same JSON line:
  e2e           the same metric through the host-buffer API (pinned host inputs, H2D, kernels, D2H of all output)
  roofline      emit kernel: (32 B/vertex + 4 B/index) / CUDA-event time of the emit launch vs measured HBM copy peak
  cpu_baseline  CPU restatement of lyon (oracle/) on the box's cores, bounded sample, rank 0, N=1 only
  flatten       configs[1] ("C2"): 10M random cubics, tolerance 0.01, flatten segments/s (N=1 only)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import numpy as np  # noqa: E402

TOTAL_PATHS = 1_000_000
N_SEGMENTS = 64
C2_CURVES = 10_000_000


def stroke_options():
    from lyon_b200 import LineCap, LineJoin, StrokeOptions
    return StrokeOptions(line_width=2.0, line_join=LineJoin.Round, tolerance=0.1).with_line_cap(LineCap.Round)


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 50 ms during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(sample_paths: int, threads: int):
    """CPU restatement of lyon's StrokeTessellator (oracle/), std::thread over paths, warm pre-sized buffers."""
    import oracle_ffi as O
    import workloads as W
    batch = W.c3_batch_range(0, sample_paths, N_SEGMENTS)
    secs, nv, ni = O.stroke_batch_timed(batch, stroke_options(), threads)
    return {"value": nv / secs, "unit": "vertices/s", "cores": threads, "kind": "port",
            "sample": f"first {sample_paths} paths of the same C3 batch ({nv} vertices, {secs:.2f} s, warm buffers)",
            "sample_vertices": int(nv), "sample_indices": int(ni)}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path, timed on the host cores.  The Rust crate
    cannot be built in this image (no cargo), so this is the oracle port of lyon (kind: port)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle_ffi as O
    import workloads as W
    threads = os.cpu_count() or 1
    sample = args.ref_sample
    batch = W.c3_batch_range(0, sample, N_SEGMENTS)
    opts = stroke_options()
    for _ in range(args.warmup):
        O.stroke_batch_timed(batch.slice(0, min(sample, 5000)), opts, threads)
    secs, nv = 0.0, 0
    for _ in range(args.steps):
        s, v, _i = O.stroke_batch_timed(batch, opts, threads)
        secs += s
        nv += v
    value = nv / secs
    line = {
        "impl": "reference", "metric": "stroke_output_vertices_per_s", "value": value, "unit": "vertices/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus, sample_note=f"each step = first {sample} paths of the 1M-path batch"),
        "cpu_baseline": {"value": value, "unit": "vertices/s", "cores": threads, "kind": "port",
                         "sample": f"first {sample} paths per step, {threads} threads, warm pre-sized buffers"},
        "e2e": {"value": value, "unit": "vertices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(n_gpus, sample_note=None):
    cfg = {
        "workload": "C3: 1M synthetic paths x 64 mixed line/quad/cubic segments, stroke width 2, round joins+caps, tol 0.1",
        "total_paths": TOTAL_PATHS, "segments_per_path": N_SEGMENTS, "parallelism": f"paths partitioned over {n_gpus} GPU(s)",
        "l2": "working set (inputs+intermediates+outputs, tens of GB per step) far exceeds the 126 MB L2; no flush needed",
    }
    if sample_note:
        cfg["sample"] = sample_note
    return cfg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--paths", type=int, default=TOTAL_PATHS, help="total paths of the batch (default: the full C3 size)")
    ap.add_argument("--cpu-sample", type=int, default=200_000)
    ap.add_argument("--ref-sample", type=int, default=200_000)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-flatten", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the C4 / C5 sub-benchmarks")
    ap.add_argument("--e2e-steps", type=int, default=3)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    import workloads as W
    from lyon_b200 import DeviceTessellator, HostPipeline
    from lyon_b200.sharding import exchange_totals

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: int) -> int:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return int(t.item())

    # ---- this rank's slice: every path has the same number of segments, so the cumulative-segment partition
    # (workloads.partition_by_cost) is an equal split of path indices; each rank generates only its own slice.
    total_paths = args.paths
    per = total_paths // world
    lo, hi = rank * per, (rank + 1) * per if rank < world - 1 else total_paths
    batch = W.c3_batch_range(lo, hi, N_SEGMENTS)
    opts = stroke_options()
    tz = DeviceTessellator(local_rank)
    dbatch = tz.upload(batch)
    torch.cuda.synchronize()
    if world > 1:  # the library's own communicator: rank 0 creates the NCCL id, torch.distributed only carries it
        ids = [tz.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        tz.comm_init(ids[0], rank, world)

    verts = idx = None
    tz.enable_timing(True)
    kernel_ms = []

    def step(ev=None, record=False):
        """ONE pass -- lyon_cuda_stroke_fused (pre-pass, then count+scan+emit fused in the tile kernel) into buffers kept
        across steps.  N>1: the rank's slice is emitted with local ids, then the all-gather of {V_r, I_r} and a streaming
        rebase of the rank's indices (lyon_cuda_rebase_indices) make the ids global."""
        nonlocal verts, idx
        if ev:
            ev[0].record()
        if world == 1:
            v, i = tz.stroke_fused(dbatch, opts)
            nv, ni = int(v.shape[0]), int(i.shape[0])
        else:
            # lyon_cuda_stroke_fused_sharded: local ids, then the one collective (ncclAllGather of {V_r, I_r, status} on the
            # context's stream) and the index rebase with the base read on the device -- no host hop between the three
            v, i, _base, _tot = tz.stroke_fused_sharded(dbatch, opts)
            nv, ni = int(v.shape[0]), int(i.shape[0])
        if ev:
            ev[1].record()
        if record:
            kernel_ms.append(tz.last_kernel_ms())
        return nv, ni

    for _ in range(max(args.warmup, 3)):
        nv, ni = step()
    torch.cuda.synchronize()
    tz.check()

    launches0 = tz.launch_count()
    events = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(args.steps)]
    sampler = ClockSampler(local_rank)
    barrier()
    torch.cuda.synchronize()
    sampler.start()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        nv, ni = step(events[k], record=True)
    torch.cuda.synchronize()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    tz.check()
    launches = tz.launch_count() - launches0
    dev_ms = events[0][0].elapsed_time(events[-1][1])
    ms_per_step = max_over_ranks(dev_ms / args.steps)
    total_v = sum_over_ranks(nv)
    total_i = sum_over_ranks(ni)
    value = total_v / (ms_per_step * 1e-3)

    # roofline of the dominant kernel (k_stroke_tile<FUSED>, at every N: the sharded entry runs the same single pass), timed by CUDA events the library
    # records around that launch on its stream.  Algorithmic bytes = inputs read once + every vertex and index written once.
    peak, peak_src = measured_peaks()
    in_bytes = dbatch.input_bytes()
    alg_bytes = in_bytes + 32 * nv + 4 * ni
    k_ms = sum(kernel_ms) / len(kernel_ms)
    k_gbs = alg_bytes / (k_ms * 1e-3) / 1e9
    roofline = {
        "bound": "hbm", "kernel": "k_stroke_tile<FUSED>", "achieved": k_gbs,
        "peak": peak, "unit": "GB/s", "frac": k_gbs / peak, "peak_source": peak_src,
        # dram__bytes_read.sum + dram__bytes_write.sum of this kernel: `ncu --set full` on the first 200k paths of the same
        # batch (profiles/r02_ncu_summary.txt) measured 8.449 GB written + 0.501 GB read = 8.951 GB for 8.686 GB of
        # algorithmic bytes.  ESTIMATE for this launch: that ratio times the launch's algorithmic bytes (a profiler cannot
        # run inside the timed region).
        "traffic": int(alg_bytes * (8.951 / 8.686)), "traffic_source": "estimated: ncu ratio of a 200k-path sample x algorithmic bytes",
        "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": k_ms, "step_ms": dev_ms / args.steps,
        "kernel_share_of_step": k_ms / (dev_ms / args.steps),
        "whole_step_gbs": alg_bytes / (dev_ms / args.steps * 1e-3) / 1e9,
        "whole_step_frac": alg_bytes / (dev_ms / args.steps * 1e-3) / 1e9 / peak,
        "frac_of_nominal_8TBs": k_gbs / 8000.0,
    }

    # ---- e2e: host buffers in, host buffers out, PCIe copies inside the timed region -------------------------------
    e2e = None
    if not args.no_e2e:
        verts = idx = None
        tz._own_v = tz._own_i = None
        torch.cuda.empty_cache()
        # >= 16 chunks per rank, so that the two alternating streams reach their steady state at every N
        from lyon_b200.numa import bind_to_device_node
        numa_node = bind_to_device_node(local_rank)  # pinned buffers on the GPU's own memory node (best effort)
        pipe = HostPipeline(local_rank, chunk_paths=max(4096, min(25000, (batch.n_paths + 15) // 16)))
        pipe.stage(batch)
        pipe.run(opts)  # warm-up: sizes + pins the output arrays
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            ev, ei = pipe.run(opts)
        torch.cuda.synchronize()
        barrier()
        e2e_s = max_over_ranks((time.perf_counter() - t0) / args.e2e_steps)
        e2e = {"value": sum_over_ranks(ev) / e2e_s, "unit": "vertices/s", "h2d_bytes_per_step": pipe.h2d_bytes,
               "d2h_bytes_per_step": pipe.d2h_bytes, "ms_per_step": e2e_s * 1e3, "api": "lyon_cuda_tessellate_paths_into (C ABI; "
               "pinned host inputs -> chunked H2D -> fused stroke kernels -> D2H of every vertex and index, two streams)",
               "chunk_paths": pipe.chunk_paths, "numa_node": numa_node,
               "d2h_gbs_per_gpu": pipe.d2h_bytes / e2e_s / 1e9}
        del pipe
        torch.cuda.empty_cache()

    flatten = None
    if world == 1 and not args.no_flatten:
        flatten = bench_flatten(tz, torch, dev, peak, args)

    configs = None
    if world == 1 and not args.no_configs:
        configs = bench_configs(tz, torch, peak)

    cpu = None
    if world == 1 and rank == 0 and not args.no_cpu:
        sample = min(args.cpu_sample, total_paths)
        cpu = cpu_baseline(sample, os.cpu_count() or 1)
        # the GPU arm on the very same sample must produce the very same totals as the CPU arm
        gv, gi = tz.stroke_fused(tz.upload(batch.slice(0, sample)), opts)
        cpu["gpu_same_sample"] = {"vertices": int(gv.shape[0]), "indices": int(gi.shape[0])}
        assert (int(gv.shape[0]), int(gi.shape[0])) == (cpu["sample_vertices"], cpu["sample_indices"]), \
            f"GPU and CPU arms disagree on the first {sample} paths: {gv.shape[0]}/{gi.shape[0]} vs {cpu['sample_vertices']}/{cpu['sample_indices']}"
        del gv, gi

    if rank == 0:
        line = {
            "metric": "stroke_output_vertices_per_s", "value": value, "unit": "vertices/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(world), "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
            "roofline": roofline, "cpu_baseline": cpu, "flatten": flatten, "configs": configs,
            "output": {"vertices": total_v, "indices": total_i, "bytes": 32 * total_v + 4 * total_i},
            "wall_ms_per_step": 1e3 * t_wall / args.steps, "workspace_bytes": tz.workspace_bytes(),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _time_fused(tz, torch, d, o, reps=3):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    best = None
    for _ in range(reps + 1):
        ev[0].record()
        v, i = tz.stroke_fused(d, o)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1])
        best = ms if best is None else min(best, ms)
    tz.check()
    return best, int(v.shape[0]), int(i.shape[0])


def bench_configs(tz, torch, peak):
    """The other configs of BASELINE.json on one GPU (device-resident inputs, whole lyon_cuda_stroke_fused call incl. its
    pre-pass): C4 = 100k polylines x 1k segments, miter limit 4, butt caps, fixed and variable line width; C5 = the two
    logo fixtures x 1000 jittered copies, tolerance sweep 1.0 .. 0.001; and the host-side packing rate of
    PathBatch.from_paths."""
    import workloads as W
    from lyon_b200 import LineJoin, PathBatch, StrokeOptions
    out = {}
    o4 = StrokeOptions(line_join=LineJoin.Miter, miter_limit=4.0)
    for vw in (False, True):
        batch = W.c4_polylines(100_000, 1000, with_attribute=vw)
        d = tz.upload(batch)
        ms, nv, ni = _time_fused(tz, torch, d, o4.with_variable_line_width(0) if vw else o4)
        alg = d.input_bytes() + 32 * nv + 4 * ni
        out["C4_variable_width" if vw else "C4_fixed_width"] = {
            "workload": "100k polylines x 1k segments, miter joins (limit 4), butt caps" + (", width attribute 0" if vw else ""),
            "ms": ms, "vertices": nv, "indices": ni, "value": nv / (ms * 1e-3), "unit": "vertices/s",
            "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg / (ms * 1e-3) / 1e9 / peak, "algorithmic_bytes": alg, "scope": "whole call"}}
        del d, batch
        tz._own_v = tz._own_i = None
        torch.cuda.empty_cache()
    paths = W.fixture_paths(ROOT / "tests" / "golden")
    t0 = time.perf_counter()
    batch = W.c5_logo_copies(paths, 1000)
    d = tz.upload(batch)
    sweep = []
    for tol in (1.0, 0.1, 0.01, 0.003, 0.001):
        ms, nv, ni = _time_fused(tz, torch, d, StrokeOptions(tolerance=tol))
        alg = d.input_bytes() + 32 * nv + 4 * ni
        sweep.append({"tolerance": tol, "ms": ms, "vertices": nv, "indices": ni, "value": nv / (ms * 1e-3),
                      "frac": alg / (ms * 1e-3) / 1e9 / peak})
    out["C5_tolerance_sweep"] = {"workload": "logo.path + rust_logo.path x 1000 jittered copies, width 1, miter, butt",
                                 "unit": "vertices/s", "sweep": sweep}
    # packing: lyon Paths -> the SoA batch the device reads (host side of the boundary)
    plist = [batch.path(i) for i in range(min(batch.n_paths, 2000))]
    t0 = time.perf_counter()
    PathBatch.from_paths(plist)
    dt = time.perf_counter() - t0
    out["packing"] = {"api": "PathBatch.from_paths (numpy concatenation)", "paths": len(plist), "paths_per_s": len(plist) / dt,
                      "verbs_per_s": float(sum(p.verbs.shape[0] for p in plist)) / dt}
    del d
    tz._own_v = tz._own_i = None
    torch.cuda.empty_cache()
    return out


def bench_flatten(tz, torch, dev, peak, args):
    """C2: 10M random cubics, tolerance 0.01 -> flatten segments/s on one B200.  A step = ONE fused kernel (Wang counts,
    decoupled look-back scan, emit) into point buffers kept across steps; the two-pass count + emit form (exact sizes
    first) is timed next to it."""
    import workloads as W
    ctrl = W.c2_cubics(C2_CURVES)
    d = torch.from_numpy(ctrl).to(dev)
    tol = 0.01
    for _ in range(3):
        off, pts = tz.flatten_curves_fused(d, 4, tol)
    torch.cuda.synchronize()
    total = int(pts.shape[0])
    steps = max(args.steps, 3)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    kms = []
    ev[0].record()
    for _ in range(steps):
        off, pts = tz.flatten_curves_fused(d, 4, tol)
        kms.append(tz.last_kernel_ms())
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / steps
    k_ms = sum(kms) / len(kms)
    # two-pass form
    for _ in range(2):
        tz.flatten_curves(d, 4, tol)
    e2 = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    e2[0].record()
    for _ in range(steps):
        tz.flatten_curves(d, 4, tol)
    e2[1].record()
    torch.cuda.synchronize()
    two_pass_ms = e2[0].elapsed_time(e2[1]) / steps
    alg = 32 * C2_CURVES + 8 * total + 8 * (C2_CURVES + 1)  # control points in, points + offsets out
    counts = (off[1:] - off[:-1])
    return {"metric": "flatten_segments_per_s", "value": total / (ms * 1e-3), "unit": "segments/s",
            "workload": "C2: 10M random cubics (seed 1234), tolerance 0.01", "segments": total, "ms_per_step": ms,
            "two_pass_ms_per_step": two_pass_ms, "mean_segments_per_curve": total / C2_CURVES,
            "max_segments_per_curve": int(counts.max()), "lut_fraction": float((counts <= 24).float().mean()),
            "roofline": {"bound": "hbm", "kernel": "k_flat_emit<4, fused>", "achieved": alg / (k_ms * 1e-3) / 1e9, "peak": peak,
                         "unit": "GB/s", "frac": alg / (k_ms * 1e-3) / 1e9 / peak, "kernel_ms": k_ms,
                         "algorithmic_bytes_per_launch": alg,
                         # ncu --set full, same launch: 0.320 GB read + 1.974 GB written (profiles/r01k_ncu_summary.txt)
                         "traffic": 2294213176}}


if __name__ == "__main__":
    main()
