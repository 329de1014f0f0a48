"""Ad-hoc GPU bring-up check: parity vs the oracle + first timings.  Run under gpurun."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import raytracer_server_b200 as R
from oracle import oracle as O

SC = os.path.join(ROOT, "tests/golden/scenes")
out = {}
for name in ("cornell_box", "cubes", "flying_unicorn"):
    path = os.path.join(SC, name + ".toml")
    t0 = time.time(); g = R.Scene.from_toml(path); t1 = time.time()
    o = O.OracleScene.from_toml(path); t2 = time.time()
    print(f"== {name}: gpu load {t1-t0:.3f}s (bvh build {g.info.build_ms:.3f} ms, nodes {g.info.bvh_nodes}), oracle load {t2-t1:.3f}s", flush=True)
    W, H = 320, 240
    org, dirs = o.primary_rays(W, H, 0, 0, 0.0, 0.0)
    ro = o.trace_rays(org, dirs)
    rg = g.trace_primary(W, H, 0, 0, 0.0, 0.0)
    mism = (ro["obj"] != rg["obj"]) | (ro["tri"] != rg["tri"])
    hit = ro["obj"] >= 0
    rel = np.abs(rg["t"][hit & ~mism] - ro["t"][hit & ~mism]) / ro["t"][hit & ~mism]
    print(f"primary ids: {mism.sum()} / {mism.size} mismatches; t rel err max {rel.max():.3e} mean {rel.mean():.3e}", flush=True)
    if mism.sum():
        idx = np.nonzero(mism)[0][:8]
        for i in idx: print("   px", i % W, i // W, "oracle", ro["obj"][i], ro["tri"][i], ro["t"][i], "gpu", rg["obj"][i], rg["tri"][i], rg["t"][i])
    # path-level: same RNG
    rng = np.random.default_rng(1)
    n = 4000; spp = 16
    px = rng.integers(0, W, n); py = rng.integers(0, H, n); si = rng.integers(0, spp, n)
    Lo = o.sample_radiance(W, H, spp, 7, px, py, si)
    Lg = g.sample_radiance(W, H, spp, px, py, si, seed=7)
    d = np.abs(Lg - Lo).max(axis=1); s = np.abs(Lo).max(axis=1) + 1e-3
    relp = d / s
    print(f"path radiance: median rel {np.median(relp):.2e}  frac(rel>1e-3) {np.mean(relp>1e-3):.4f} frac(rel>1e-1) {np.mean(relp>1e-1):.4f}  mean Lo {Lo.mean():.5f} Lg {Lg.mean():.5f}", flush=True)
    # small image, sub-pixel means
    w, h, spp = 96, 72, 64
    t0 = time.time(); io = o.render(w, h, spp, seed=3, nthreads=-os.cpu_count(), want_sub=False); t1 = time.time()
    ig = g.render(w, h, spp, seed=3); t2 = time.time()
    df = np.abs(ig.astype(int) - io["rgb8"].astype(int))
    print(f"image {w}x{h}x{spp}: oracle {t1-t0:.2f}s gpu {t2-t1:.3f}s  |diff| mean {df.mean():.3f} max {df.max()} frac>2 {np.mean(df>2):.4f}; mean oracle {io['rgb8'].mean():.2f} gpu {ig.mean():.2f}", flush=True)
    st = g.stats(); print("   stats", {k: (round(v,3) if isinstance(v,float) else v) for k,v in st.items()}, "oracle rays", io["rays"], flush=True)
    for use_mis in (True,):
        o.set_modes(O.ACCEL_EXACT, O.EST_MIS_DEAD)
        io = o.render(w, h, 16, seed=3, nthreads=-os.cpu_count())
        ig = g.render(w, h, 16, seed=3, use_mis=True)
        df = np.abs(ig.astype(int) - io["rgb8"].astype(int))
        print(f"MIS-dead image: |diff| mean {df.mean():.3f} max {df.max()} frac>2 {np.mean(df>2):.4f}; mean oracle {io['rgb8'].mean():.2f} gpu {ig.mean():.2f}", flush=True)
        o.set_modes(O.ACCEL_EXACT, O.EST_NEE)
    # throughput
    for (w, h, spp) in ((600, 450, 64), (1920, 1080, 16)):
        g.render(w, h, 4)
        t0 = time.time(); g.render(w, h, spp); dt = time.time() - t0
        st = g.stats()
        rays = st["rays_primary"] + st["rays_extension"] + st["rays_shadow"]
        print(f"throughput {w}x{h}x{spp}: wall {dt:.3f}s dev {st['render_ms']:.1f} ms ext_ms {st['extend_ms']:.1f} -> {st['samples']/st['render_ms']/1e3:.2f} Msamples/s {rays/st['render_ms']/1e3:.1f} Mrays/s iters {st['iterations']}", flush=True)
print("fp32 peak TFLOP/s", R.fp32_peak_tflops(0))
