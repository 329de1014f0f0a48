"""Generates the converged (4096 spp) golden frames with the CPU oracle (f64, exact accel).

BASELINE.json gate: "converged 4096-spp images must match within 1 % mean relative error per channel and
PSNR >= 40 dB, with MIS on and off".  The oracle needs ~25 minutes per frame on 8 cores, so the frames are
rendered once here and committed (uint8 RGB, npz-compressed); tests/test_gpu_converged.py renders the same
seed on the GPU (same RNG contract) and applies the gate.  Resumable: existing files are skipped.

`faithful` adds the reference's REAL mesh behaviour: `to_scene` always calls `mesh.accelerate()`
(src/scene.rs:430-432) and `_intersect_recurse` returns the first child with any hit (src/geometry.rs:1263-1273),
so the Rust binary renders the octree_faithful image, not the exact one.  Those frames (same seed, same RNG
contract, 1024 spp: the octree costs 2.2x the exact BVH on the CPU) are the "reference-defect floor" that
tests/test_gpu_converged.py measures the exact oracle and the GPU against.

usage: python tests/golden/make_converged.py [threads] [faithful [spp]]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "converged")
W, H, SPP, SEED = 600, 450, 4096, 7     # the reference server's frame size (src/server.rs:29-30)
JOBS = [("cornell_box", 0), ("cubes", 0), ("cubes", 1), ("flying_unicorn", 0), ("cornell_box", 1), ("flying_unicorn", 1)]

FAITHFUL_JOBS = [("cubes", 0), ("flying_unicorn", 0)]   # the scenes with meshes; live estimator

if __name__ == "__main__":
    threads = int(sys.argv[1]) if len(sys.argv) > 1 else (os.cpu_count() or 1)
    faithful = len(sys.argv) > 2 and sys.argv[2] == "faithful"
    spp = int(sys.argv[3]) if len(sys.argv) > 3 else (1024 if faithful else SPP)
    os.makedirs(OUT, exist_ok=True)
    for scene, est in (FAITHFUL_JOBS if faithful else JOBS):
        tag = "octree_" if faithful else ""
        path = os.path.join(OUT, f"{scene}_{tag}{'mis' if est else 'nee'}_{W}x{H}_{spp}spp_seed{SEED}.npz")
        if os.path.exists(path):
            continue
        sc = O.OracleScene.from_toml(os.path.join(ROOT, "tests", "golden", "scenes", scene + ".toml"))
        sc.set_modes(O.ACCEL_OCTREE_FAITHFUL if faithful else O.ACCEL_EXACT, est)
        t0 = time.time()
        r = sc.render(W, H, spp, seed=SEED, nthreads=-threads)
        np.savez_compressed(path, rgb8=r["rgb8"], width=W, height=H, spp=spp, seed=SEED, estimator=est,
                            accel=0 if faithful else 1, rays=r["rays"], samples=r["samples"], seconds=time.time() - t0)
        print(f"{path}: {time.time() - t0:.0f} s, {r['samples']} samples", flush=True)
