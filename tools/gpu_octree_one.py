"""One octree-mode frame (for an ncu capture of k_traverse_octree).  Run under gpurun."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raytracer_server_b200 as R
os.environ["RTB_NO_GRAPH"] = "1"
g = R.Scene.from_toml(os.path.join(ROOT, "tests/golden/scenes/flying_unicorn.toml"))
g.render(1920, 1080, int(sys.argv[1]) if len(sys.argv) > 1 else 16, seed=1, accel=R.ACCEL_OCTREE_REFERENCE)
st = g.stats()
print(f"octree frame: dev {st['render_ms']:.1f} ms traverse {st['extend_ms']:.1f} bin {st['bin_ms']:.1f} shade {st['shade_ms']:.1f} iters {st['iterations']}")
