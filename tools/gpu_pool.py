import os, sys
sys.path.insert(0, os.getcwd())
import raytracer_server_b200 as R
g = R.Scene.from_toml("tests/golden/scenes/flying_unicorn.toml")
w,h,spp=1920,1080,256
for P in (1<<24, 1<<25, 3<<24, 1<<26, 1<<25):
    g.render(w,h,16,pool_paths=P)
    g.render(w,h,spp,seed=1,pool_paths=P)
    st=g.stats()
    print(f"pool {P>>20}M: {st['samples']/st['render_ms']/1e3:.1f} Msamples/s iters {st['iterations']} traverse {st['extend_ms']:.0f} ms shade {st['shade_ms']:.0f} ms total {st['render_ms']:.0f} ms", flush=True)
