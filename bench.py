#!/usr/bin/env python
"""bench.py -- hide+filter throughput of the B200 hider on BASELINE.json's synthetic scenes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 4] [--sub 1,2,3,5] [--impl reference]

A "step" is one pass of the hot path (project+bust, bin, sample, composite, filter, expose, quantise, and for N>1 the
NCCL gather of the strips) over one frame of synthetic shaded grids.  The headline workload is config 4 of
BASELINE.json (3840x2160, PixelSamples 16 16, four layers, semi-transparent): the configuration the north-star target
is quoted on; it fits one B200.  `value` is measured with the grids already resident in HBM; `e2e` goes through the
public C ABI with pinned HOST buffers and the library's capture display as callbacks (H2D of the grids, D2H of the image
and the per-bucket DspyImageData-style delivery inside the timed region), warm (frame tables cached) and `e2e_cold`
(caches cleared: the host replay of the renderer's random stream is paid again, like the first frame of a process).
At N=1 the line also carries `sub`: the same measurements for configs 1, 2, 3 and the 30-filter sweep of config 5.

N>1: one process per GPU under torchrun; one contiguous strip of pixel rows per rank (strong scaling of one frame),
grids replicated to the ranks whose strips they touch (aqh_grid_rank_masks), finished strips gathered to rank 0 by the
library's own grouped ncclSend/ncclRecv (aqh_gather).

--impl reference times the reference's own CPU implementation for the same metric and config: aqsis' libs/core hider
sources compiled in place (oracle/_ref/libaqsis_refhider.so), one single-threaded process per host core
(aqsis' hider cannot thread), each step a bounded, same-density sample of the workload.  That arm never maps the product
library: its parameter blocks are built in pure python and the pixel filter is the reference's own, chosen by name.
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
    5: "config5: PixelFilter sweep (box, triangle, gaussian, catmull-rom, sinc at widths 1-6) on the config-2 scene",
}
L2_POLICY = "inputs (>= 0.8 GB of grids per frame at configs 2-5) and the resolved-sample planes (>= 4 GB) exceed the 126 MB L2"
# same-density reduced copies of the workloads for the CPU legs (linear image scale)
CPU_SAMPLE_SCALE = {1: 1.0, 2: 0.25, 3: 0.08, 4: 0.05}


def config_dict(config):
    """The `config` object of the JSON line: the same for the product arm and the reference arm."""
    return {"workload": WORKLOADS[config], "l2_policy": L2_POLICY}


def make_scene(config, scale=1.0, **kw):
    from aqsis_b200 import scenes
    fn = {1: scenes.config1, 2: scenes.config2, 3: scenes.config3, 4: scenes.config4}[config]
    return fn(scale=scale, **kw)


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
        med = float(np.median(sm)) if sm else None
        return {"sm_mhz": med, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------------------------
# CPU legs: the reference's own hider (oracle/_ref) or, when it did not travel, the oracle port.  tests/orc.py is the
# ctypes face of both; nothing here touches the product library.
def _ref_worker(_):
    import orc
    t0 = time.perf_counter()
    orc.render_reference(_REF_SCENE[0], _REF_SCENE[1])
    return time.perf_counter() - t0


_REF_SCENE = None


def cpu_reference(config, procs, rounds=1, scale=None):
    """Time the reference's own CPU hider on a bounded, same-density sample of the workload.

    kind "reference": oracle/_ref/libaqsis_refhider.so -- aqsis' libs/core hider sources compiled in place
    (single-threaded, process-global state, exactly like aqsis).  To use every host core the way a render farm uses
    aqsis, `procs` independent processes each render the sample frame; throughput = procs * micropolygons / wall time.
    kind "port": the oracle restatement with buckets over threads (only when the reference library did not travel)."""
    global _REF_SCENE
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import multiprocessing as mp
    from aqsis_b200 import hider as _hider
    _hider.PURE = True                 # parameter blocks in pure python: the product library is never mapped here
    try:
        import orc
        scale = CPU_SAMPLE_SCALE[config] if scale is None else scale
        params, grids = make_scene(config, scale)
    finally:
        _hider.PURE = False
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


def cpu_filter_baseline(procs):
    """config 5's CPU leg: the FilterBucket stage alone (oracle port: the reference's hider has no per-stage entry point;
    the port's filter loop restates bucketprocessor.cpp:584-664 and is pinned bit for bit to it), catmull-rom 4x4."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from aqsis_b200 import hider as _hider
    _hider.PURE = True
    try:
        import orc
        params, grids = make_scene(2, 0.25, filter=("catmull-rom", 4.0, 4.0))
    finally:
        _hider.PURE = False
    orc.build_oracle()
    _, _, st = orc.render(params, grids, procs)
    taps = 25 * params.xsamples * params.ysamples
    px = params.xres * params.yres
    # filter_s is summed over the bucket threads: wall time of the stage = filter_s / threads
    wall = st["filter_s"] / max(1, st["threads"])
    return {"value": px * taps / wall / 1e6, "unit": "Mtaps/s", "cores": procs, "kind": "port",
            "sample": f"config2 scene at linear scale 0.25 ({params.xres}x{params.yres}), catmull-rom 4x4 ({taps} taps per pixel): "
                      f"FilterBucket stage of oracle/oracle_hider.cpp, buckets over {procs} threads"}


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
        "config": config_dict(args.config),
        "cpu_baseline": {k: cb[k] for k in cb if k in ("value", "unit", "cores", "kind", "sample", "value_1process")},
        "e2e": {"value": cb["value"], "unit": "Mpolys/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "product_library_mapped": any("libaqsis_b200_hider" in l for l in open("/proc/self/maps")),
    }
    print(json.dumps(line))
    return 0


def cpu_baseline(config):
    cb = cpu_reference(config, os.cpu_count() or 1, 1)
    return {k: cb[k] for k in cb if k in ("value", "unit", "cores", "kind", "sample", "value_1process")}


# --------------------------------------------------------------------------------------------------------------------
class Bench:
    """One process of the product arm (rank `rank` of `world`)."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from aqsis_b200 import Hider, build
        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the hider has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        if self.rank == 0:
            build.build()
        if self.world > 1:
            dist.barrier()
        self.stream = torch.cuda.current_stream()
        self.h = Hider(self.local_rank, stream=self.stream.cuda_stream)
        if self.world > 1:
            # the library's own communicator for the gather of the strips; the id travels over torch.distributed
            from aqsis_b200.hider import comm_unique_id
            idt = torch.zeros(128, dtype=torch.uint8, device=self.dev)
            if self.rank == 0:
                idt.copy_(torch.frombuffer(bytearray(comm_unique_id()), dtype=torch.uint8))
            dist.broadcast(idt, 0)
            self.h.comm_init(bytes(idt.cpu().numpy().tobytes()), self.rank, self.world)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        self.peak = float(peaks.get("hbm_gbs", 6650.0))
        self.peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)" if "hbm_gbs" in peaks else "6650 GB/s (fallback of B200_PROFILING.md)"

    def sync_all(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def timed(self, fn, steps, warmup, before_each=None):
        torch, h = self.torch, self.h
        for _ in range(warmup):
            if before_each:
                before_each()
            fn()
        self.sync_all()
        stage = {"project_bust_ms": 0.0, "render_mpgs_ms": 0.0, "filter_ms": 0.0, "gather_ms": 0.0, "launches": 0}
        total = 0.0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if before_each is None:
            e0.record()
        for _ in range(steps):
            if before_each:                       # untimed preparation between steps (cache clearing)
                before_each()
                self.sync_all()
                e0.record()
            fn()
            s = h.stats()
            for k in ("project_bust_ms", "render_mpgs_ms", "filter_ms"):
                stage[k] += s[k]
            stage["launches"] += s["gpu_launches"]
            if before_each:
                e1.record()
                self.sync_all()
                total += e0.elapsed_time(e1)
        if before_each is None:
            e1.record()
            self.sync_all()
            total = e0.elapsed_time(e1)
        ms = total / steps
        if self.world > 1:
            t = torch.tensor([ms], device=self.dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t.item())
        for k in ("project_bust_ms", "render_mpgs_ms", "filter_ms"):
            stage[k] /= steps
        stage["gather_ms"] = float(h.stats().get("gather_ms", 0.0))
        return ms, stage

    def measure(self, config, steps, warmup, scale=1.0, e2e=True, cold=True, partials=False, scene_kw=None, strip_rows=None):
        """Resident and end-to-end measurements of one workload; returns the record (rank 0) or None."""
        from aqsis_b200 import abi as _abi, scenes, sharding
        from aqsis_b200.hider import display_info
        torch, h, rank, world = self.torch, self.h, self.rank, self.world
        scene_kw = scene_kw or {}
        if config == 4 and world > 1:
            # the 5.4 GB scene is generated layer by layer and sharded on the fly: N ranks never hold N full copies
            def shard(p, block):
                p.rank, p.world_size, p.strip_rows = rank, world, (-1 if strip_rows is None else strip_rows)
                return sharding.split_grids_for_rank(p, block, rank, world)
            params, mine = scenes.config4(scale=scale, shard=shard)
            params.rank, params.world_size, params.strip_rows = rank, world, (-1 if strip_rows is None else strip_rows)
            n_mp_total = int(mine.total_micropolygons)
            b_alg_total = int(mine.total_vbytes) + scenes.algorithmic_bytes(params, mine) - int(mine.n_verts) * 36
        else:
            params, grids = make_scene(config, scale, **scene_kw)
            params.rank, params.world_size = rank, world
            if world > 1:
                if strip_rows is None:
                    sharding.balance_strips(params, [grids])        # one contiguous strip per rank, equal estimated work
                else:
                    params.strip_rows = strip_rows
            n_mp_total = grids.n_micropolygons
            b_alg_total = scenes.algorithmic_bytes(params, grids)
            mine = sharding.split_grids_for_rank(params, grids, rank, world)
            del grids
        b_alg_mine = scenes.algorithmic_bytes(params, mine)
        crop_px = (params.crop_xmax - params.crop_xmin) * (params.crop_ymax - params.crop_ymin)
        n_samples_total = crop_px * params.xsamples * params.ysamples
        dev_grids = mine.to_torch(device=self.dev)
        pin_grids = mine.to_torch(pin=True) if e2e else None
        es_d = display_info(params, 0)[2] if params.n_displays else 0

        def step_resident():
            h.render_device()
            if world > 1:
                h.gather(0)                        # finished strips to rank 0 over NCCL (the only collective of the path)

        # host images the capture display fills bucket by bucket (what a framebuffer driver does with DspyImageData)
        cap_ch = np.zeros((params.yres, params.xres, 9), np.float32) if e2e else None
        cap_d = [np.zeros((params.yres, params.xres, es_d), np.uint8)] if (e2e and params.n_displays) else []

        def step_e2e():
            h.begin_frame(params)
            h.add_grid_block(pin_grids)
            h.end_frame_capture(cap_ch, cap_d)     # gathers (N>1), downloads on rank 0, delivers every bucket in reference order

        h.begin_frame(params)
        h.add_grid_block(dev_grids)
        sampler = ClockSampler(self.local_rank) if (rank == 0 and config == self.args.config) else None
        if sampler:
            sampler.start()
        ms, stage = self.timed(step_resident, steps, warmup)
        clocks = sampler.finish() if sampler else None
        stats = h.stats()
        rec = {"workload": WORKLOADS[config] + ("" if scale == 1.0 else f" [scale {scale}]"),
               "value": n_mp_total / (ms * 1e-3) / 1e6, "unit": "Mpolys/s", "ms_per_step": ms, "frame_ms": ms,
               "msamples_per_s": n_samples_total / (ms * 1e-3) / 1e6, "micropolygons": n_mp_total, "samples": n_samples_total,
               "stages_ms": {k: round(v, 4) for k, v in stage.items() if k.endswith("_ms")},
               "gpu_launches": int(stage["launches"]), "bands": int(stats.get("n_bands", 1)),
               "device_bytes": int(stats.get("device_bytes", 0)),
               "counters": {k: int(stats[k]) for k in ("n_grids", "n_vertices", "n_micropolygons", "n_bin_entries", "n_deep_hits")}}
        hide_ms = stage["render_mpgs_ms"]
        # dominant kernel: k_hide.  Algorithmic bytes of its launches in one frame = the vertex data of the grids this
        # rank hides, V*(12K+24) (sample state stays on chip and is not compulsory traffic); the duration
        # is the sum of the frame's k_hide launches (one per band of tile rows), CUDA events on the launch stream.
        nvb = b_alg_mine - crop_px * (36 + es_d)
        achieved = nvb / (hide_ms * 1e-3) / 1e9 if hide_ms > 0 else 0.0
        traffic, ncu_note = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(str(config))
            if tj and world == 1 and scale == 1.0:
                # DRAM bytes of the frame's k_hide launches: one captured launch x the launches of a frame (bands)
                if tj.get("dram_bytes_per_launch"):
                    traffic = int(tj["dram_bytes_per_launch"]) * int(tj.get("launches_per_frame", 1))
                ncu_note = {k: tj[k] for k in ("sm_issue_active_pct", "warp_instructions", "source", "launches_per_frame", "dram_bytes_per_launch") if k in tj}
        except Exception:
            pass
        rec["roofline"] = {"bound": "hbm", "kernel": "k_hide", "achieved": achieved, "peak": self.peak, "unit": "GB/s",
                           "frac": achieved / self.peak, "traffic": traffic, "peak_source": self.peak_src,
                           "algorithmic_bytes_launch": int(nvb), "k_hide_ms_per_frame": hide_ms, "ncu": ncu_note,
                           "note": "issue-slot bound by design of the path (SURVEY 8d): the HBM fraction is reported as the contract asks; ncu issue utilisation is the figure that moves",
                           "frame_achieved_gbs": b_alg_total / (ms * 1e-3) / 1e9, "frame_frac": b_alg_total / (ms * 1e-3) / 1e9 / self.peak,
                           "algorithmic_bytes_frame": int(b_alg_total)}
        if partials:
            params.filter_mode = _abi.FILTER_TILE_PARTIALS
            h.begin_frame(params)
            h.add_grid_block(dev_grids)
            tiled_ms, tiled_stage = self.timed(step_resident, max(3, steps // 2), 2)
            params.filter_mode = _abi.FILTER_REFERENCE_ORDER
            rec["tile_partials_mode"] = {"ms_per_step": tiled_ms, "stages_ms": {k: round(v, 4) for k, v in tiled_stage.items() if k.endswith("_ms")},
                                         "note": "opt-in filter mode, not bit-exact (different association of the filter sums)"}
        if e2e:
            e_ms, _ = self.timed(step_e2e, max(3, steps // 2), 2)
            s2 = h.stats()
            rec["e2e"] = {"value": n_mp_total / (e_ms * 1e-3) / 1e6, "unit": "Mpolys/s", "ms_per_step": e_ms,
                          "h2d_bytes_per_step": int(s2["h2d_bytes"]), "d2h_bytes_per_step": int(s2["d2h_bytes"]),
                          "callbacks": "aqh_capture_on_bucket + aqh_capture_on_data per bucket, reference bucket order (rank 0)",
                          "last_step_ms": {k: round(float(s2[k]), 3) for k in ("prepare_ms", "upload_ms", "device_total_ms", "download_ms") if k in s2}}
            if cold:
                c_ms, _ = self.timed(step_e2e, 3, 1, before_each=h.clear_caches)
                rec["e2e_cold"] = {"value": n_mp_total / (c_ms * 1e-3) / 1e6, "unit": "Mpolys/s", "ms_per_step": c_ms,
                                   "prepare_ms": round(float(h.stats()["prepare_ms"]), 3),
                                   "note": "frame-table cache cleared before every step: the host replay of the renderer's random stream is inside the timed region"}
        rec["_clocks"] = clocks
        rec["_strips"] = len(sharding.strips_for_rank(params, 0))
        del dev_grids, pin_grids
        torch.cuda.empty_cache()
        return rec

    def filter_sweep(self, steps):
        """config 5: the 30 PixelFilter combinations on the resident config-2 scene; filter-only Mtaps/s."""
        from aqsis_b200 import scenes
        h = self.h
        rows, dev_grids = [], None
        for name, w in scenes.config5_filters():
            p, g = scenes.config2(filter=(name, w, w))
            if dev_grids is None:
                dev_grids = g.to_torch(device=self.dev)
            del g
            h.begin_frame(p)
            h.add_grid_block(dev_grids)
            for _ in range(2):
                h.render_device()
            acc = {"project_bust_ms": 0.0, "render_mpgs_ms": 0.0, "filter_ms": 0.0}
            for _ in range(steps):
                h.render_device()
                s = h.stats()
                for k in acc:
                    acc[k] += s[k] / steps
            shift = int(w // 2)
            taps = (2 * shift + 1) ** 2 * p.xsamples * p.ysamples
            px = p.xres * p.yres
            rows.append({"filter": name, "width": w, "taps_per_pixel": taps, "frame_ms": round(sum(acc.values()), 3),
                         "hide_ms": round(acc["render_mpgs_ms"], 3), "filter_ms": round(acc["filter_ms"], 3),
                         "filter_mtaps_per_s": round(px * taps / acc["filter_ms"] / 1e3, 1)})
        del dev_grids
        self.torch.cuda.empty_cache()
        tot_taps = sum(r["taps_per_pixel"] for r in rows) * 1920 * 1080
        tot_ms = sum(r["filter_ms"] for r in rows)
        return {"workload": WORKLOADS[5], "value": tot_taps / tot_ms / 1e3, "unit": "Mtaps/s",
                "ms_per_step": sum(r["frame_ms"] for r in rows), "filter_ms_total": tot_ms, "steps": steps,
                "note": "value = taps of all 30 filters / their k_filter_spans time (filter stage only, CUDA events); ms_per_step = the 30 whole frames (hide + filter)",
                "filters": rows}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", type=int, default=4, choices=[1, 2, 3, 4])
    ap.add_argument("--sub", default="1,2,3,5", help="sub-records at N=1 (comma separated configs; empty = none)")
    ap.add_argument("--scale", type=float, default=1.0, help="linear image scale of the workload (1.0 = as named)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--strip-rows", type=int, default=None, help="strip dealing (AqhFrameParams::strip_rows); default: one contiguous strip per rank")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    real_stdout = os.dup(1)
    os.dup2(2, 1)            # libraries (NCCL version banner, torchrun notices) must not pollute the one JSON line
    B = Bench(args)
    world, rank = B.world, B.rank
    if world != args.gpus and world > 1:
        args.gpus = world
    t_start = time.time()
    head = B.measure(args.config, args.steps, args.warmup, scale=args.scale, e2e=not args.no_e2e, cold=not args.no_e2e,
                     partials=False, strip_rows=args.strip_rows)
    sub = {}
    if world == 1 and args.scale == 1.0:
        for c in [int(x) for x in args.sub.split(",") if x.strip()]:
            if c == args.config:
                continue
            t0 = time.time()
            if c == 5:
                r = B.filter_sweep(3)
                if not args.no_cpu_baseline:
                    r["cpu_baseline"] = cpu_filter_baseline(os.cpu_count() or 1)
            else:
                r = B.measure(c, 5 if c == 3 else args.steps, 3, e2e=not args.no_e2e, cold=not args.no_e2e, partials=(c == 2))
                r.pop("_clocks", None), r.pop("_strips", None)
                if not args.no_cpu_baseline:
                    r["cpu_baseline"] = cpu_baseline(c)
            r["bench_seconds"] = round(time.time() - t0, 1)
            sub[f"config{c}"] = r
    if rank == 0:
        clocks, nstrips = head.pop("_clocks"), head.pop("_strips")
        line = {
            "metric": "hide+filter throughput", "value": head["value"], "unit": "Mpolys/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "msamples_per_s": head["msamples_per_s"], "frame_ms": head["frame_ms"],
            "config": config_dict(args.config),
            "workload": {"name": head["workload"], "micropolygons": head["micropolygons"], "samples": head["samples"],
                         "parallelism": f"{world} rank(s), {nstrips} contiguous strip(s) of pixel rows per rank, grids replicated to the ranks they "
                                        f"touch (aqh_grid_rank_masks), finished strips to rank 0 by grouped ncclSend/ncclRecv (aqh_gather)"},
            "stages_ms": head["stages_ms"], "gpu_launches": head["gpu_launches"], "bands": head["bands"],
            "device_bytes": head["device_bytes"],
            "roofline": head["roofline"],
            "filter_mode": "reference-order (bit-exact sums)",
            "clocks": clocks, "e2e": head.get("e2e"), "e2e_cold": head.get("e2e_cold"),
            "counters": head["counters"],
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(args.config)
        if sub:
            line["sub"] = sub
        line["bench_seconds"] = round(time.time() - t_start, 1)
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        B.dist.barrier()
        B.dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
