#!/usr/bin/env python
"""bench.py -- hide+filter throughput of the B200 hider on BASELINE.json's synthetic scenes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 2] [--impl reference]

A "step" is one pass of the hot path (project+bust, bin, sample, composite, filter, expose,
quantise, and for N>1 the NCCL gather of the image) over one frame of synthetic shaded grids.
`value` is measured with the grids already resident in HBM; `e2e` goes through the public
C ABI with pinned HOST buffers (H2D of the grids and D2H of the image inside the timed region).
N>1: one process per GPU under torchrun, image strips dealt round-robin (strong scaling of one
frame), grids replicated to the ranks whose strips they touch, final image gathered to rank 0.

--impl reference times the reference's own CPU implementation for the same metric: aqsis'
libs/core hider sources compiled in place (oracle/_ref/libaqsis_refhider.so), one
single-threaded process per host core, each step a bounded, same-density sample of the workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    1: "config1: 640x480, PixelSamples 4 4, gaussian 2x2, 10k bilinear patches (0.64M micropolygons)",
    2: "config2: 1920x1080, PixelSamples 8 8, ShadingRate 1, ~20M micropolygons, opaque, catmull-rom 3x3",
    3: "config3: 1920x1080, PixelSamples 8 8, motion blur (shutter 0 1) + depth of field, ~20M micropolygons",
    4: "config4: 3840x2160, PixelSamples 16 16, ShadingRate 0.25, 4 layers, semi-transparent, gaussian 2x2",
}
# same-density reduced copies of the workloads for the CPU legs (linear image scale)
CPU_SAMPLE_SCALE = {1: 1.0, 2: 0.25, 3: 0.08, 4: 0.05}


def make_scene(config, scale=1.0):
    from aqsis_b200 import scenes
    fn = {1: scenes.config1, 2: scenes.config2, 3: scenes.config3, 4: scenes.config4}[config]
    return fn(scale=scale)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = threading.Event()
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag.is_set():
                    break
                parts = [x.strip() for x in line.split(",")]
                if len(parts) >= 7:
                    self.samples.append(parts)
        except Exception:
            pass

    def finish(self):
        self.stop_flag.set()
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass
        sm, mx, reasons = [], 0, set()
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx = max(mx, float(s[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # under load = samples within 25% of the top observed clock or all if few
        med = float(np.median(sm)) if sm else None
        return {"sm_mhz": med, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


class _CudaArray:
    """Expose a raw device pointer to torch through __cuda_array_interface__."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 2}


def _ref_worker(_):
    import orc
    t0 = time.perf_counter()
    orc.render_reference(_REF_SCENE[0], _REF_SCENE[1])
    return time.perf_counter() - t0


_REF_SCENE = None


def cpu_reference(config, procs, rounds=1):
    """Time the reference's own CPU hider on a bounded, same-density sample of the workload.

    kind "reference": oracle/_ref/libaqsis_refhider.so -- aqsis' libs/core hider sources compiled in
    place (single-threaded, process-global state, exactly like aqsis).  To use every host core the way a
    render farm uses aqsis, `procs` independent processes each render the sample frame; throughput =
    procs * micropolygons / wall time.  kind "port": the oracle restatement with buckets over threads
    (only when the reference library did not travel to this machine)."""
    global _REF_SCENE
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import multiprocessing as mp
    import orc
    scale = CPU_SAMPLE_SCALE[config]
    params, grids = make_scene(config, scale)
    nmp = grids.n_micropolygons
    nsamp = params.xres * params.yres * params.xsamples * params.ysamples
    what = (f"{WORKLOADS[config].split(':')[0]} at linear scale {scale} ({params.xres}x{params.yres}, {grids.n_grids} grids, "
            f"{nmp} micropolygons, same density and options)")
    out = {"unit": "Mpolys/s", "cores": procs}
    if orc.refhider() is not None:
        _REF_SCENE = (params, grids)
        t0 = time.perf_counter()
        orc.render_reference(params, grids)
        t1 = time.perf_counter() - t0
        out["value_1process"] = nmp / t1 / 1e6
        if procs > 1:
            ctx = mp.get_context("fork")
            with ctx.Pool(procs) as pool:
                t0 = time.perf_counter()
                for _ in range(rounds):
                    pool.map(_ref_worker, range(procs))
                dt = (time.perf_counter() - t0) / rounds
            out["value"] = procs * nmp / dt / 1e6
        else:
            dt = t1
            out["value"] = out["value_1process"]
        out.update(kind="reference", ms_per_frame=dt * 1e3, msamples_per_s=procs * nsamp / dt / 1e6,
                   sample=what + f"; aqsis' own libs/core hider compiled in place (oracle/_ref), {procs} independent "
                                 f"single-threaded processes each rendering the sample frame")
    else:
        orc.build_oracle()
        t0 = time.perf_counter()
        for _ in range(rounds):
            orc.render(params, grids, procs)
        dt = (time.perf_counter() - t0) / rounds
        out.update(kind="port", value=nmp / dt / 1e6, ms_per_frame=dt * 1e3, msamples_per_s=nsamp / dt / 1e6,
                   sample=what + f"; oracle/oracle_hider.cpp, buckets over {procs} threads")
    return out


def run_reference(args):
    """The reference arm: aqsis' own CPU hider on all host cores, each step a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    for _ in range(max(0, args.warmup - 1)):
        cpu_reference(args.config, cores, 1)
    cb = cpu_reference(args.config, cores, max(1, args.steps))
    line = {
        "impl": "reference", "metric": "hide+filter throughput", "value": cb["value"], "unit": "Mpolys/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_frame"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "msamples_per_s": cb["msamples_per_s"],
        "config": {"workload": WORKLOADS[args.config], "sample": cb["sample"]},
        "cpu_baseline": {k: cb[k] for k in cb if k in ("value", "unit", "cores", "kind", "sample", "value_1process")},
        "e2e": {"value": cb["value"], "unit": "Mpolys/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def cpu_baseline(config):
    cb = cpu_reference(config, os.cpu_count() or 1, 1)
    return {k: cb[k] for k in cb if k in ("value", "unit", "cores", "kind", "sample", "value_1process")}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4])
    ap.add_argument("--scale", type=float, default=1.0, help="linear image scale of the workload (1.0 = as named)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--strip-rows", type=int, default=0, help="strip height; 0 = balanced (world*k near-equal strips)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    real_stdout = os.dup(1)
    os.dup2(2, 1)            # libraries (NCCL version banner, torchrun notices) must not pollute the one JSON line
    import torch
    import torch.distributed as dist
    from aqsis_b200 import Hider, build, scenes, sharding
    from aqsis_b200.hider import display_info

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hider has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        build.build()
    if world > 1:
        dist.barrier()

    if args.config == 4 and world > 1:
        # the 5.4 GB scene is generated layer by layer and sharded on the fly: N ranks never hold N full copies
        def shard(p, block):
            p.rank, p.world_size, p.strip_rows = rank, world, args.strip_rows
            return sharding.split_grids_for_rank(p, block, rank, world)
        params, mine = scenes.config4(scale=args.scale, shard=shard)
        params.rank, params.world_size, params.strip_rows = rank, world, args.strip_rows
        n_mp_total = int(mine.total_micropolygons)
        px_bytes = (params.crop_xmax - params.crop_xmin) * (params.crop_ymax - params.crop_ymin)
        b_alg_total = int(mine.total_vbytes) + scenes.algorithmic_bytes(params, mine) - int(mine.n_verts) * 36
        b_alg_mine = scenes.algorithmic_bytes(params, mine)
    else:
        params, grids = make_scene(args.config, args.scale)
        params.rank, params.world_size, params.strip_rows = rank, world, args.strip_rows
        n_mp_total = grids.n_micropolygons
        b_alg_total = scenes.algorithmic_bytes(params, grids)
        mine = sharding.split_grids_for_rank(params, grids, rank, world)
        b_alg_mine = scenes.algorithmic_bytes(params, mine) if world > 1 else b_alg_total
        del grids
    n_samples_total = (params.crop_xmax - params.crop_xmin) * (params.crop_ymax - params.crop_ymin) * params.xsamples * params.ysamples

    stream = torch.cuda.current_stream()
    h = Hider(local_rank, stream=stream.cuda_stream)
    dev_grids = mine.to_torch(device=dev)
    pin_grids = mine.to_torch(pin=True)
    h2d_bytes = sum(int(t.numel() * t.element_size()) for t in (pin_grids.P, pin_grids.Ci, pin_grids.Oi) if t is not None)

    # ---- gather plumbing: rows owned by each rank (index tensors built once)
    h.begin_frame(params)
    h.add_grid_block(dev_grids)
    h.render_device()
    assert h.strips() == sharding.strips_for_rank(params, rank)
    dtype_d, nch_d, es_d = display_info(params, 0) if params.n_displays else (np.dtype("uint8"), 0, 0)
    gather = sharding.ImageGather(params, rank, world, dev, dist if world > 1 else None)

    def device_images():
        pc, _ = h.device_channels()
        ch = torch.as_tensor(_CudaArray(pc, (params.yres, params.xres * 9), "<f4"), device=dev)
        dsp = None
        if params.n_displays:
            pd, _ = h.device_display(0)
            dsp = torch.as_tensor(_CudaArray(pd, (params.yres, params.xres * es_d), "|u1"), device=dev)
        return ch, dsp

    final = {}

    def gather_image():
        """Final image to rank 0 over NCCL (the only collective of the path)."""
        ch, dsp = device_images()
        res = gather([ch] if dsp is None else [ch, dsp])
        if res is not None:
            final["channels"] = res[0]
            final["display"] = res[1] if len(res) > 1 else None

    def step_resident():
        h.render_device()
        gather_image()

    def step_e2e():
        h.begin_frame(params)
        h.add_grid_block(pin_grids)
        h.end_frame(fetch=False)     # the image lands in the library's pinned host buffers (D2H inside the call)
        gather_image()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stage = {"project_bust_ms": 0.0, "render_mpgs_ms": 0.0, "filter_ms": 0.0, "launches": 0}
        e0.record()
        for _ in range(steps):
            fn()
            s = h.stats()
            for k in ("project_bust_ms", "render_mpgs_ms", "filter_ms"):
                stage[k] += s[k]
            stage["launches"] += s["gpu_launches"]
        e1.record()
        sync_all()
        ms = e0.elapsed_time(e1) / steps
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        for k in ("project_bust_ms", "render_mpgs_ms", "filter_ms"):
            stage[k] /= steps
        return ms, stage

    # resident-input measurement
    h.begin_frame(params)
    h.add_grid_block(dev_grids)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    ms, stage = timed(step_resident, args.steps, args.warmup)
    clocks = sampler.finish() if sampler else None
    stats = h.stats()

    # the opt-in tile-partials filter mode (not bit-exact), for information
    from aqsis_b200 import abi as _abi
    params.filter_mode = _abi.FILTER_TILE_PARTIALS
    h.begin_frame(params)
    h.add_grid_block(dev_grids)
    tiled_ms, tiled_stage = timed(step_resident, max(3, args.steps // 2), 2)
    params.filter_mode = _abi.FILTER_REFERENCE_ORDER

    e2e = None
    if not args.no_e2e:
        e_ms, _ = timed(step_e2e, max(3, args.steps // 2), 2)
        s2 = h.stats()
        e2e = {"value": n_mp_total / (e_ms * 1e-3) / 1e6, "unit": "Mpolys/s", "ms_per_step": e_ms,
               "h2d_bytes_per_step": int(s2["h2d_bytes"]), "d2h_bytes_per_step": int(s2["d2h_bytes"]),
               "last_step_ms": {k: round(float(s2[k]), 3) for k in ("prepare_ms", "upload_ms", "device_total_ms", "download_ms") if k in s2}}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)"
        hide_ms = stage["render_mpgs_ms"]
        # dominant kernel: k_hide.  Algorithmic bytes of one launch = the vertex data of the grids this
        # rank hides, V*(12K+24) (sample state stays on chip and is not compulsory traffic).
        nvb = b_alg_mine - (params.crop_xmax - params.crop_xmin) * (params.crop_ymax - params.crop_ymin) * (36 + es_d)
        achieved = nvb / (hide_ms * 1e-3) / 1e9 if hide_ms > 0 else 0.0
        frame_achieved = b_alg_total / (ms * 1e-3) / 1e9
        traffic, ncu_note = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(str(args.config))
            if tj and world == 1 and args.scale == 1.0:
                traffic = int(tj["dram_bytes_per_launch"])
                ncu_note = {k: tj[k] for k in ("sm_issue_active_pct", "warp_instructions", "source") if k in tj}
        except Exception:
            pass
        line = {
            "metric": "hide+filter throughput", "value": n_mp_total / (ms * 1e-3) / 1e6, "unit": "Mpolys/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "msamples_per_s": n_samples_total / (ms * 1e-3) / 1e6, "frame_ms": ms,
            "config": {"workload": WORKLOADS[args.config] + ("" if args.scale == 1.0 else f" [scale {args.scale}]"),
                       "micropolygons": n_mp_total, "samples": n_samples_total,
                       "l2_policy": "inputs (>= 0.8 GB) and sample planes (>= 4 GB) exceed the 126 MB L2",
                       "parallelism": f"{world} rank(s), {len(sharding.strips_for_rank(params, 0))} strip(s) of pixel rows per rank dealt round-robin, grids replicated to the ranks they touch, NCCL gather to rank 0"},
            "stages_ms": {k: round(v, 4) for k, v in stage.items() if k.endswith("_ms")},
            "gpu_launches": int(stage["launches"]),
            "roofline": {"bound": "hbm", "kernel": "k_hide", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_launch": int(nvb), "ncu": ncu_note,
                         "note": "issue-slot bound by design of the path (SURVEY 8d): the HBM fraction is reported as the contract asks; ncu issue utilisation is the figure that moves",
                         "frame_achieved_gbs": frame_achieved, "frame_frac": frame_achieved / peak,
                         "algorithmic_bytes_frame": int(b_alg_total)},
            "filter_mode": "reference-order (bit-exact sums)",
            "tile_partials_mode": {"ms_per_step": tiled_ms, "stages_ms": {k: round(v, 4) for k, v in tiled_stage.items() if k.endswith("_ms")}},
            "clocks": clocks, "e2e": e2e,
            "counters": {k: int(stats[k]) for k in ("n_grids", "n_vertices", "n_micropolygons", "n_bin_entries", "n_deep_hits")},
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(args.config)
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
