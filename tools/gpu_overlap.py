"""Experiment: do two concurrent renders (two streams, half the CTAs each) beat one render at full occupancy?
Run under gpurun; knobs RTB_TRAV_CTAS / RTB_SHADE_CTAS are read when a render context is created."""
import os, sys, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raytracer_server_b200 as R
SC = os.path.join(ROOT, "tests/golden/scenes")
name, w, h, spp = "flying_unicorn", 1920, 1080, 64
nthr = int(sys.argv[1]) if len(sys.argv) > 1 else 2
g = R.Scene.from_toml(os.path.join(SC, name + ".toml"))
def one(seed, s):
    g.render(w, h, s, seed=seed)
for rep in range(2):
    ts = [threading.Thread(target=one, args=(10 + i, spp // nthr)) for i in range(nthr)]
    t0 = time.time()
    for t in ts: t.start()
    for t in ts: t.join()
    dt = time.time() - t0
    print(f"{nthr} concurrent renders of {spp // nthr} spp: wall {dt*1e3:.1f} ms -> {w*h*spp/dt/1e6:.1f} Msamples/s (TRAV_CTAS={os.environ.get('RTB_TRAV_CTAS')}, SHADE_CTAS={os.environ.get('RTB_SHADE_CTAS')})", flush=True)
