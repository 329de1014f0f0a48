#!/usr/bin/env python
"""bench.py — samples/s (and Mrays/s) of the B200 path tracer on the reference's headline scene.

Contract: `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line (rank 0).

Workload (BASELINE.json configs):
  N = 1 : configs[2]  flying_unicorn at 1920x1080, 256 spp (the BVH-heavy scene the target is quoted on)
  N > 1 : configs[3]  flying_unicorn at 3840x2160, tile-sharded over the ranks, 64*N spp
          (weak scaling: 530.8 M samples per GPU at every N, the same count as the N = 1 frame),
          RGB8 shards all_gathered with NCCL and scattered into scan-line order.
A "step" is one complete frame: generation -> wavefront iterations until every path has ended ->
per-sub-pixel resolve (clamp, gamma, RGB8).  Nothing is cached between steps (accumulators are
cleared, every sample is re-traced; the seed changes per step).

  value    : whole-job samples/s, scene + BVH resident in HBM, frame left in device memory
  e2e      : same metric through the host-buffer C-ABI call (rtb_scene_upload from pinned memory +
             rtb_render into a host buffer: H2D of the flattened scene and D2H of the frame inside
             the timed region)
  roofline : the kernel with the largest share of the step (k_shade or k_traverse) against the HBM peak
             by its algorithmic queue bytes; `fp32` is the whole step against the measured FP32 FMA
             peak with SURVEY §8(d)'s flop model; `roofline_other` is the second kernel
  cpu_baseline / --impl reference : oracle/ (f64 C++ port of the reference, octree-faithful, live
             NEE) on the host cores, bounded sample of the same frame.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
SCENES = os.path.join(ROOT, "tests", "golden", "scenes")
SCENE = "flying_unicorn"
METRIC = "samples/sec (Mrays/sec alongside) on flying_unicorn"
UNIT = "samples/s"
# algorithmic HBM bytes of the two hot kernels per unit (DESIGN.md "Kernels and rooflines"):
#   k_shade    per path-queue entry read (camera rays + re-queued paths; a path stays in registers from vertex
#              to vertex while its next hit is analytic): hit 8 + origin/direction/throughput 3 x 16 = 56 B
#              per path-queue entry written: origin/direction/throughput/hit             = 56 B
#              per queued shadow ray written: origin/direction/contribution              = 48 B
#              per NEE contribution added directly: one 16-byte RED                       = 16 B
#   k_traverse per BVH extension ray: origin/direction 32 + hit read 8 + hit write 8     = 48 B
#              per BVH shadow ray   : origin/direction/contribution 48 + one RED 16      = 64 B
#              (+ the LBVH itself, 4 MB, which lives in L1/L2)
SHADE_B_VERTEX, SHADE_B_EXT, SHADE_B_SHQ, SHADE_B_RED = 56.0, 56.0, 48.0, 16.0
TRAV_B_EXT, TRAV_B_SH = 48.0, 64.0


def workload(n_gpus: int):
    if n_gpus == 1:
        return {"width": 1920, "height": 1080, "spp": 256, "name": "flying_unicorn 1920x1080 256spp (configs[2])"}
    return {"width": 3840, "height": 2160, "spp": 64 * n_gpus,
            "name": f"flying_unicorn 3840x2160 {64 * n_gpus}spp tile-sharded over {n_gpus} GPUs (configs[3], 64 spp per GPU)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_run(width, height, spp, seconds_target=20.0, threads=None):
    """Times the oracle port (reference algorithm: octree-faithful traversal, live NEE, f64) on the host
    cores over a bounded sample of the frame: every `stride`-th row at `spp_cpu` spp."""
    from oracle import oracle as O

    threads = threads or os.cpu_count() or 1
    sc = O.OracleScene.from_toml(os.path.join(SCENES, SCENE + ".toml"))
    sc.set_modes(O.ACCEL_OCTREE_FAITHFUL, O.EST_NEE)
    # calibrate on a thin sample, then size the real one for ~seconds_target
    stride = max(1, height // 16)
    t0 = time.perf_counter()
    r = sc.render(width, height, 4, seed=1, nthreads=-threads, row_stride=stride)
    dt = time.perf_counter() - t0
    rate = r["samples"] / max(dt, 1e-6)
    want = rate * seconds_target
    spp_cpu = 16
    rows = max(threads, int(want / (width * spp_cpu)))
    stride = max(1, height // rows)
    t0 = time.perf_counter()
    r = sc.render(width, height, spp_cpu, seed=2, nthreads=-threads, row_stride=stride)
    dt = time.perf_counter() - t0
    n_rows = (height + stride - 1) // stride
    return {"value": r["samples"] / dt, "mrays": r["rays"] / dt / 1e6, "seconds": dt, "cores": threads,
            "sample": f"{n_rows} of {height} rows (every {stride}th) x {width} px at {spp_cpu} spp, "
                      f"{r['samples']} samples, scaled linearly (samples are i.i.d.)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = workload(args.gpus)
    steps = max(1, args.steps)
    per_step = max(3.0, min(20.0, 90.0 / (steps + args.warmup)))
    vals = []
    for _ in range(args.warmup):
        cpu_reference_run(wl["width"], wl["height"], wl["spp"], seconds_target=min(per_step, 3.0))
    t0 = time.perf_counter()
    for _ in range(steps):
        vals.append(cpu_reference_run(wl["width"], wl["height"], wl["spp"], seconds_target=per_step))
    total = time.perf_counter() - t0
    v = sum(x["value"] for x in vals) / len(vals)
    last = vals[-1]
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": total / steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "reference scene fixtures (tests/golden/scenes), CPU",
        "config": {"workload": wl["name"], "width": wl["width"], "height": wl["height"], "spp": wl["spp"]},
        "mrays_per_s": sum(x["mrays"] for x in vals) / len(vals),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": last["cores"], "kind": "port", "sample": last["sample"]},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="rtb200", choices=["rtb200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--spp", type=int, default=0, help="override spp (debug only; invalidates the bench line)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import ctypes as C

    import numpy as np
    import torch

    import raytracer_server_b200 as R
    from raytracer_server_b200 import _abi, sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            print(f"bench.py: --gpus {args.gpus} needs torchrun (WORLD_SIZE={world})", file=sys.stderr)
            return 2
    if not torch.cuda.is_available():
        print("bench.py: no CUDA device (there is no CPU fallback for the rtb200 arm)", file=sys.stderr)
        return 2
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)

    wl = workload(world)
    W, H, SPP = wl["width"], wl["height"], args.spp or wl["spp"]
    scene = R.Scene.from_toml(os.path.join(SCENES, SCENE + ".toml"), device=local_rank)
    stride = sharding.shard_stride(W, H, world)
    shard = torch.zeros(stride, dtype=torch.uint8, device=dev)
    gathered = torch.empty(world * stride, dtype=torch.uint8, device=dev) if world > 1 else None
    frame = torch.empty((H, W, 3), dtype=torch.uint8, device=dev)
    samples_per_step_total = W * H * (SPP // 4) * 4

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step(i, totals):
        p = R.make_params(W, H, SPP, seed=1000 + i, rank=rank, world=world)
        scene.render_device(p, shard.data_ptr())
        st = scene.stats()
        for k in ("samples", "rays_primary", "rays_extension", "rays_shadow", "kernel_launches", "iterations", "rays_bvh",
                  "shadow_bvh", "paths_queued"):
            totals[k] = totals.get(k, 0) + st[k]
        for k in ("render_ms", "extend_ms", "shade_ms", "resolve_ms"):
            totals[k] = totals.get(k, 0.0) + st[k]
        if world > 1:
            dist.all_gather_into_tensor(gathered, shard)
            torch.cuda.synchronize(dev)
            if rank == 0:
                R.host._check(_abi.lib().rtb_untile_device(C.byref(p), C.c_void_p(gathered.data_ptr()), stride,
                                                           C.c_void_p(frame.data_ptr()), local_rank))
            totals["kernel_launches"] += 1

    for i in range(args.warmup):
        step(i, {})
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    totals = {}
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(args.warmup + i, totals)
    barrier()
    elapsed = time.perf_counter() - t0
    clk = clocks.stop() if rank == 0 else None

    # max over ranks of the bracketed time; sums of the counters
    if dist is not None:
        t = torch.tensor([elapsed], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed = float(t.item())
        keys = ["samples", "rays_primary", "rays_extension", "rays_shadow", "kernel_launches"]
        c = torch.tensor([float(totals[k]) for k in keys], dtype=torch.float64, device=dev)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        job = dict(zip(keys, [float(x) for x in c.tolist()]))
    else:
        job = {k: float(v) for k, v in totals.items()}
    value = job["samples"] / elapsed
    rays = job["rays_primary"] + job["rays_extension"] + job["rays_shadow"]

    out = None
    if rank == 0:
        # ---- e2e: host-buffer C-ABI call, H2D scene upload + D2H frame inside the timed region (this rank's GPU;
        # at N > 1 every rank would do the same on its shard, so the per-GPU figure is scaled by N)
        host_frame = np.zeros((H, W, 3), dtype=np.uint8)
        e2e_steps = max(1, min(args.steps, 2))
        h2d = scene.upload()
        scene.render(W, H, SPP, seed=5, rank=rank, world=world, out=host_frame)
        t1 = time.perf_counter()
        e2e_samples = 0
        for i in range(e2e_steps):
            h2d = scene.upload()
            scene.render(W, H, SPP, seed=2000 + i, rank=rank, world=world, out=host_frame)
            e2e_samples += scene.stats()["samples"]
        e2e_dt = time.perf_counter() - t1
        d2h = int(sharding.local_pixels(W, H, rank, world) * 3) if world > 1 else W * H * 3
        e2e_value = e2e_samples / e2e_dt * world

        # ---- rooflines of the two hot kernels, measured live over the timed region (CUDA events around every launch,
        # recorded on the render stream inside librtb200 and summed); the one with the larger share is `roofline`
        n_launch = max(1.0, totals["iterations"])
        vertices = totals["rays_primary"] + totals["rays_extension"]      # path vertices shaded == closest-hit rays
        shadow_direct = totals["rays_shadow"] - totals["shadow_bvh"]       # upper bound of the REDs issued by k_shade
        entries_read = totals["rays_primary"] + totals["paths_queued"]     # what k_shade pulls out of the path queue
        shade_bytes = (entries_read * SHADE_B_VERTEX + totals["paths_queued"] * SHADE_B_EXT + totals["shadow_bvh"] * SHADE_B_SHQ
                       + shadow_direct * SHADE_B_RED)
        trav_bytes = totals["rays_bvh"] * TRAV_B_EXT + totals["shadow_bvh"] * TRAV_B_SH
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        traffic = {}
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic = json.load(f)
        except Exception:
            pass

        # measured DRAM traffic (ncu --set full, profiles/traffic.json) is stored per unit and scaled to this run's
        # units per launch, because the path pool (hence a launch) is sized per frame
        traffic_per_launch = {
            "k_shade": traffic.get("k_shade_dram_bytes_per_vertex", 0.0) * vertices / n_launch or None,
            "k_traverse": traffic.get("k_traverse_dram_bytes_per_bvh_ray", 0.0) * (totals["rays_bvh"] + totals["shadow_bvh"]) / n_launch or None,
        }

        def roof(kernel, ms, nbytes):
            gbs = nbytes / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
            return {"kernel": kernel, "bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                    "traffic": traffic_per_launch[kernel],
                    # neither hot kernel is HBM bound (DESIGN.md "Rooflines"): what limits them is instruction issue, so the
                    # ncu-measured issue utilisation and SIMD lanes per instruction of the committed profile ride along
                    "issue_active_pct_ncu": traffic.get(kernel + "_issue_active_pct"),
                    "lanes_per_instruction_ncu": traffic.get(kernel + "_lanes_per_instruction"),
                    "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650",
                    "algorithmic_bytes_per_launch": nbytes / n_launch, "avg_launch_ms": ms / n_launch,
                    "share_of_step": ms / max(totals["render_ms"], 1e-9)}

        roofs = [roof("k_shade", totals["shade_ms"], shade_bytes), roof("k_traverse", totals["extend_ms"], trav_bytes)]
        roofs.sort(key=lambda r: -r["share_of_step"])
        # FP32 view: flop model of SURVEY §8(d) with node visits / triangle tests measured by a counting pass
        pc = R.make_params(640, 360, 16, seed=9, count_work=True)
        cnt_frame = torch.zeros(sharding.shard_stride(640, 360, 1), dtype=torch.uint8, device=dev)
        scene.render_device(pc, cnt_frame.data_ptr())
        cs = scene.stats()
        crays = cs["rays_primary"] + cs["rays_extension"] + cs["rays_shadow"]
        n_node, n_tri = cs["bvh_node_visits"] / crays, cs["bvh_tri_tests"] / crays
        info = scene.info
        flops_per_ray = 14 * info.n_planes + 20 * info.n_spheres + 48 * n_node + 46 * n_tri
        flops_per_vertex = 120.0
        fp32_peak = R.fp32_peak_tflops(local_rank)
        step_s = totals["render_ms"] * 1e-3
        fp32_achieved = ((totals["rays_primary"] + totals["rays_extension"] + totals["rays_shadow"]) * flops_per_ray
                         + vertices * flops_per_vertex) / step_s / 1e12

        cpu = None
        if not args.no_cpu_baseline and world == 1:
            c = cpu_reference_run(W, H, SPP, seconds_target=15.0)
            cpu = {"value": c["value"], "unit": UNIT, "cores": c["cores"], "kind": "port", "sample": c["sample"],
                   "mrays_per_s": c["mrays"]}

        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed / max(1, args.steps) * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "reference scene fixtures (tests/golden/scenes), Philox seeds per step",
            "config": {"workload": wl["name"], "width": W, "height": H, "spp": SPP, "samples_per_step": samples_per_step_total,
                       "parallelism": f"tiles32x32 interleaved x{world}" if world > 1 else "single GPU",
                       "l2": "no flush needed: every step re-traces the whole frame, streaming ~270 GB through HBM (path / shadow queue entries of the paths that leave registers + the 127 MiB accumulator buffer, cleared per step), far more than the 126 MB L2; only the 4 MB scene + LBVH is meant to stay cache-resident"},
            "mrays_per_s": rays / elapsed / 1e6,
            "rays_per_sample": rays / max(1.0, job["samples"]),
            "device_ms_per_step": totals["render_ms"] / max(1, args.steps),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_dt / e2e_steps * 1e3, "api": "rtb_scene_upload + rtb_render (host RGB8 frame)"},
            "gpu_launches": int(job["kernel_launches"]),
            "clocks": clk,
            "roofline": dict(roofs[0], fp32={"achieved": fp32_achieved, "peak": fp32_peak, "unit": "TFLOP/s",
                                             "frac": fp32_achieved / fp32_peak, "flops_per_ray": flops_per_ray,
                                             "flops_per_vertex": flops_per_vertex, "bvh_nodes_per_ray": n_node,
                                             "tri_tests_per_ray": n_tri, "scope": "whole step, flop model of SURVEY 8(d)",
                                             "peak_source": "rtb_fp32_peak FMA chain, this run"}),
            "roofline_other": roofs[1],
            "bvh_ray_fraction": {"extension": totals["rays_bvh"] / max(1.0, vertices),
                                 "shadow": totals["shadow_bvh"] / max(1.0, totals["rays_shadow"])},
            "cpu_baseline": cpu,
        }
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
