"""torchrun check: N ranks render interleaved tiles, NCCL all_gather, untile; frame must equal the 1-rank frame."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import raytracer_server_b200 as R
from raytracer_server_b200 import sharding

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
scene = R.Scene.from_toml(os.path.join(ROOT, "tests/golden/scenes/flying_unicorn.toml"), device=lr)
W, H, SPP = 1000, 700, 16
frame = sharding.render_sharded(scene, W, H, SPP, seed=3).cpu().numpy()
ok = True
if rank == 0:
    whole = scene.render(W, H, SPP, seed=3)
    d = np.abs(frame.astype(int) - whole.astype(int))
    ok = d.max() <= 1
    print(f"world {world}: sharded frame vs single-GPU frame: max |diff| {d.max()}, mean {frame.mean():.2f} -> {'OK' if ok else 'MISMATCH'}", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
