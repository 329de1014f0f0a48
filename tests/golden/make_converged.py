"""Generates the converged (4096 spp) golden frames with the CPU oracle (f64, exact accel).

BASELINE.json gate: "converged 4096-spp images must match within 1 % mean relative error per channel and
PSNR >= 40 dB, with MIS on and off".  The oracle needs ~25 minutes per frame on 8 cores, so the frames are
rendered once here and committed (uint8 RGB, npz-compressed); tests/test_gpu_converged.py renders the same
seed on the GPU (same RNG contract) and applies the gate.  Resumable: existing files are skipped.

usage: python tests/golden/make_converged.py [threads]
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

if __name__ == "__main__":
    threads = int(sys.argv[1]) if len(sys.argv) > 1 else (os.cpu_count() or 1)
    os.makedirs(OUT, exist_ok=True)
    for scene, est in JOBS:
        path = os.path.join(OUT, f"{scene}_{'mis' if est else 'nee'}_{W}x{H}_{SPP}spp_seed{SEED}.npz")
        if os.path.exists(path):
            continue
        sc = O.OracleScene.from_toml(os.path.join(ROOT, "tests", "golden", "scenes", scene + ".toml"))
        sc.set_modes(O.ACCEL_EXACT, est)
        t0 = time.time()
        r = sc.render(W, H, SPP, seed=SEED, nthreads=-threads)
        np.savez_compressed(path, rgb8=r["rgb8"], width=W, height=H, spp=SPP, seed=SEED, estimator=est,
                            rays=r["rays"], samples=r["samples"], seconds=time.time() - t0)
        print(f"{path}: {time.time() - t0:.0f} s, {r['samples']} samples", flush=True)
