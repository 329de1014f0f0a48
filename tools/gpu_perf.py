"""Quick per-kernel timing breakdown (CUDA events inside librtb200).  Run under gpurun."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raytracer_server_b200 as R
SC = os.path.join(ROOT, "tests/golden/scenes")
cfgs = [("flying_unicorn", 1920, 1080, 32), ("cornell_box", 1920, 1080, 32), ("cubes", 600, 450, 256)]
if len(sys.argv) > 1:
    cfgs = [c for c in cfgs if c[0] in sys.argv[1:]]
for name, w, h, spp in cfgs:
    g = R.Scene.from_toml(os.path.join(SC, name + ".toml"))
    if not os.environ.get("RTB_PERF_NOWARM"):
        g.render(w, h, 8)
    t0 = time.time(); g.render(w, h, spp, seed=1); dt = time.time() - t0
    st = g.stats()
    rays = st["rays_primary"] + st["rays_extension"] + st["rays_shadow"]
    it = st["iterations"]
    print(f"{name} {w}x{h}x{spp}: wall {dt*1e3:.1f} ms dev {st['render_ms']:.1f} ms | extend {st['extend_ms']:.1f} shade {st['shade_ms']:.1f} bin {st['bin_ms']:.1f} other {st['render_ms']-st['extend_ms']-st['shade_ms']-st['bin_ms']:.1f} | iters {it} "
          f"| {st['samples']/st['render_ms']/1e3:.1f} Msamples/s {rays/st['render_ms']/1e3:.0f} Mrays/s | bvh rays {st['rays_bvh']/max(1,st['rays_primary']+st['rays_extension']):.3f} shadow bvh {st['shadow_bvh']/max(1,st['rays_shadow']):.3f}"
          f" | per iter: ext {st['extend_ms']/it*1e3:.0f} us shade {st['shade_ms']/it*1e3:.0f} us bin {st['bin_ms']/it*1e3:.0f} us", flush=True)
# small frames (launch-bound regime): BASELINE configs[0] and [1]
for name, w, h, spp in (("cornell_box", 600, 450, 64), ("cubes", 600, 450, 256)):
    if len(sys.argv) > 1 and name not in sys.argv[1:]:
        continue
    g = R.Scene.from_toml(os.path.join(SC, name + ".toml"))
    g.render(w, h, spp, seed=2)
    t0 = time.time(); g.render(w, h, spp, seed=1); dt = time.time() - t0
    st = g.stats()
    print(f"{name} {w}x{h}x{spp} (config): wall {dt*1e3:.1f} ms dev {st['render_ms']:.1f} ms iters {st['iterations']} -> {st['samples']/dt/1e6:.1f} Msamples/s wall", flush=True)
