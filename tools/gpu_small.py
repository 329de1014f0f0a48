import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raytracer_server_b200 as R
g = R.Scene.from_toml(os.path.join(ROOT, "tests/golden/scenes/cornell_box.toml"))
for w, h, spp in ((600, 450, 4), (600, 450, 4), (600, 450, 16), (300, 225, 4), (64, 48, 4)):
    t0 = time.perf_counter(); g.render(w, h, spp, seed=1); dt = time.perf_counter() - t0
    st = g.stats()
    print(f"{w}x{h}x{spp}: wall {dt*1e3:.2f} ms dev {st['render_ms']:.2f} ms resolve {st['resolve_ms']:.3f} iters {st['iterations']} -> {st['render_ms']/st['iterations']*1e3:.1f} us/iter launches {st['kernel_launches']}", flush=True)
os.environ["RTB_NO_GRAPH"] = "1"
for w, h, spp in ((600, 450, 4), (600, 450, 4)):
    t0 = time.perf_counter(); g.render(w, h, spp, seed=1); dt = time.perf_counter() - t0
    st = g.stats()
    print(f"no graph {w}x{h}x{spp}: wall {dt*1e3:.2f} ms dev {st['render_ms']:.2f} ms iters {st['iterations']} -> {st['render_ms']/st['iterations']*1e3:.1f} us/iter", flush=True)
