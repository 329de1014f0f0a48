"""Measures every BASELINE.json config on one GPU plus the CPU port beside it -> profiles/r1_configs.json."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import raytracer_server_b200 as R
from oracle import oracle as O

SC = os.path.join(ROOT, "tests/golden/scenes")
NCPU = os.cpu_count() or 1
out = {"host_cores": NCPU, "configs": []}

def gpu(scene, w, h, spp, mis=False, reps=3):
    g = R.Scene.from_toml(os.path.join(SC, scene + ".toml"))
    g.render(w, h, min(spp, 16), use_mis=mis)
    best = None
    for i in range(reps):
        t0 = time.perf_counter(); g.render(w, h, spp, seed=10 + i, use_mis=mis); dt = time.perf_counter() - t0
        st = g.stats()
        rays = st["rays_primary"] + st["rays_extension"] + st["rays_shadow"]
        r = {"wall_s": dt, "samples_per_s": st["samples"] / dt, "mrays_per_s": rays / dt / 1e6, "rays_per_sample": rays / st["samples"],
             "iterations": st["iterations"]}
        if best is None or r["wall_s"] < best["wall_s"]:
            best = r
    return best

def cpu(scene, w, h, mis, threads, seconds=8.0):
    sc = O.OracleScene.from_toml(os.path.join(SC, scene + ".toml"))
    sc.set_modes(O.ACCEL_OCTREE_FAITHFUL, O.EST_MIS_DEAD if mis else O.EST_NEE)
    stride = max(1, h // 12)
    t0 = time.perf_counter(); r = sc.render(w, h, 4, seed=1, nthreads=-threads if threads > 1 else 1, row_stride=stride); dt = time.perf_counter() - t0
    rate = r["samples"] / dt
    rows = max(threads, int(rate * seconds / (w * 8)))
    stride = max(1, h // rows)
    t0 = time.perf_counter(); r = sc.render(w, h, 8, seed=2, nthreads=-threads if threads > 1 else 1, row_stride=stride); dt = time.perf_counter() - t0
    return {"samples_per_s": r["samples"] / dt, "mrays_per_s": r["rays"] / dt / 1e6, "threads": threads, "sampled": f"every {stride}th row at 8 spp"}

cfgs = [("C1 cornell_box 600x450 64spp", "cornell_box", 600, 450, 64, False),
        ("C2 cubes 600x450 256spp MIS off", "cubes", 600, 450, 256, False),
        ("C2 cubes 600x450 256spp MIS on (dead branch)", "cubes", 600, 450, 256, True),
        ("C3 flying_unicorn 1920x1080 256spp", "flying_unicorn", 1920, 1080, 256, False),
        ("C4 flying_unicorn 3840x2160 4096spp (1 GPU)", "flying_unicorn", 3840, 2160, 4096, False)]
for name, scene, w, h, spp, mis in cfgs:
    reps = 1 if spp >= 4096 else 3
    g = gpu(scene, w, h, spp, mis, reps)
    c1 = cpu(scene, w, h, mis, 1, seconds=6.0)
    cn = cpu(scene, w, h, mis, NCPU, seconds=6.0)
    row = {"config": name, "gpu": g, "cpu_1_thread": c1, "cpu_all_threads": cn}
    out["configs"].append(row)
    print(f"{name}: GPU {g['samples_per_s']/1e6:.1f} Msamples/s {g['mrays_per_s']:.0f} Mrays/s ({g['wall_s']*1e3:.1f} ms) | CPU 1 thr {c1['samples_per_s']/1e6:.3f} | CPU {NCPU} thr {cn['samples_per_s']/1e6:.3f} Msamples/s", flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r1_configs.json"), "w"), indent=1)
