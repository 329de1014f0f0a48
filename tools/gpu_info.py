"""Scene / LBVH sizes and build times of the three reference scenes.  Run under gpurun."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raytracer_server_b200 as R
for n in ("flying_unicorn", "cubes", "cornell_box"):
    g = R.Scene.from_toml(os.path.join(ROOT, "tests/golden/scenes", n + ".toml"))
    i = g.info
    print(n, "nodes", i.bvh_nodes, "leaves", i.bvh_leaves, "tris", i.n_triangles, "build_ms", round(i.build_ms, 2), "planes", i.n_planes, "spheres", i.n_spheres)
