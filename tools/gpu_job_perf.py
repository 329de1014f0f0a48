"""Streaming-job measurements: time to first record / job throughput vs the blocking render, and BASELINE config 5
(progressive cornell_box, 1 sample per pixel per frame) in frames/s.  Run under gpurun."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import raytracer_server_b200 as R
SC = os.path.join(ROOT, "tests/golden/scenes")
out = {}
g = R.Scene.from_toml(os.path.join(SC, "flying_unicorn.toml"))
for (w, h, spp) in ((1920, 1080, 256), (600, 450, 256)):
    g.render(w, h, 8)
    t0 = time.time(); g.render(w, h, spp, seed=3); blocking = (time.time() - t0) * 1e3
    for workers, growth in (("1", "8"), ("2", "8"), ("3", "8"), ("2", "1"), ("2", "4"), ("2", "16")):
        os.environ["RTB_JOB_WORKERS"] = workers
        os.environ["RTB_BAND_GROWTH_CAP"] = growth
        best = None
        for rep in range(2):
            job = R.RenderJob(g, w, h, spp, seed=3)
            n = sum(1 for _ in job.messages())
            st = job.stats(); job.close()
            if best is None or st["wall_ms"] < best["wall_ms"]:
                best = st
        print(f"flying_unicorn {w}x{h}x{spp} workers {workers} growth cap {growth}: blocking {blocking:.1f} ms | job wall {best['wall_ms']:.1f} ms first record {best['first_record_ms']:.1f} ms "
              f"({100*best['first_record_ms']/best['wall_ms']:.1f} %) | {best['samples']/best['wall_ms']/1e3:.1f} Msamples/s | iterations {best['iterations']}", flush=True)
        out[f"job_{w}x{h}x{spp}_w{workers}_g{growth}"] = {"blocking_ms": blocking, **{k: best[k] for k in ("wall_ms", "first_record_ms", "samples", "iterations")}}
    os.environ.pop("RTB_JOB_WORKERS")
    os.environ.pop("RTB_BAND_GROWTH_CAP")
c = R.Scene.from_toml(os.path.join(SC, "cornell_box.toml"))
W, H = 600, 450
for passes in (256, 1024):
    job = R.RenderJob(c, W, H, passes, seed=1, passes=passes)
    t0 = time.time(); n = 0
    for idx, f in job.frames():
        n += 1
    dt = time.time() - t0
    st = job.stats(); job.close()
    print(f"cornell_box progressive {W}x{H}: {n} frames in {dt:.3f} s = {n/dt:.0f} frames/s (first frame after {st['first_record_ms']:.2f} ms)", flush=True)
    out[f"progressive_{passes}"] = {"frames": n, "seconds": dt, "fps": n / dt, "first_frame_ms": st["first_record_ms"]}
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r2_job_perf.json"), "w"), indent=1)
