"""Tree quality vs traversal time: the builders behind RTB_BVH on the bench scene.  Run under gpurun."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raytracer_server_b200 as R
os.environ["RTB_NO_GRAPH"] = "1"          # direct launches: per-kernel times in the stats
W, H, SPP = 1920, 1080, int(os.environ.get("SPP", "128"))
for mode in sys.argv[1:] or ("", "lbvh", "sah"):
    # "sah:BINS:SWEEP:LEAF:CT" sets the host builder's knobs
    parts = mode.split(":")
    if parts[0]: os.environ["RTB_BVH"] = parts[0]
    else: os.environ.pop("RTB_BVH", None)
    for k, v in zip(("RTB_SAH_BINS", "RTB_SAH_SWEEP", "RTB_SAH_LEAF", "RTB_SAH_CT"), parts[1:] + [""] * 4):
        if v: os.environ[k] = v
        else: os.environ.pop(k, None)
    t0 = time.perf_counter()
    g = R.Scene.from_toml(os.path.join(ROOT, "tests/golden/scenes/flying_unicorn.toml"))
    load = time.perf_counter() - t0
    g.render(W, H, 8, seed=2)
    best = None
    for rep in range(3):
        t0 = time.perf_counter(); g.render(W, H, SPP, seed=1); dt = time.perf_counter() - t0
        st = g.stats(); st["wall"] = dt
        if best is None or dt < best["wall"]: best = st
    i = g.info
    print(f"RTB_BVH={mode or 'sah (default)'}: load {load*1e3:.0f} ms nodes {i.bvh_nodes} depth {i.bvh_depth} | frame {best['wall']*1e3:.1f} ms traverse {best['extend_ms']:.1f} shade {best['shade_ms']:.1f} -> {best['samples']/best['wall']/1e6:.1f} Msamples/s", flush=True)
    del g
