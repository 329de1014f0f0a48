"""Multi-GPU frame assembly: one process per GPU, image tiles interleaved over ranks.

The path shards with no data-path collective (pixels are independent): every rank
renders the 32x32 tiles t with t % world == rank into a tile-ordered RGB8 shard; ONE collective
per frame (all_gather over NCCL / NVLink, gloo in the CPU tests) collects the shards and a scatter
kernel (``rtb_untile_device``) — or numpy on the CPU — puts them into scan-line order.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _abi
from .host import Scene, _check, make_params


def local_pixels(width: int, height: int, rank: int, world: int) -> int:
    p = make_params(width, height, 4, rank=rank, world=world)
    return _check(_abi.lib().rtb_local_pixels(C.byref(p)))


def shard_stride(width: int, height: int, world: int) -> int:
    """Bytes of the largest shard (all_gather needs equal sizes; smaller shards are zero padded)."""
    return max(local_pixels(width, height, r, world) for r in range(world)) * 3


def tile_map(width: int, height: int, rank: int, world: int) -> np.ndarray:
    """[n_local, 2] int32 (x, y) per local pixel slot; -1 for slots outside the frame."""
    p = make_params(width, height, 4, rank=rank, world=world)
    n = local_pixels(width, height, rank, world)
    xy = np.full((max(n, 1), 2), -1, dtype=np.int32)
    _check(_abi.lib().rtb_tile_map(C.byref(p), xy.ctypes.data_as(C.POINTER(C.c_int32)), n))
    return xy[:n]


def untile_numpy(shards: np.ndarray, width: int, height: int, world: int) -> np.ndarray:
    """CPU twin of rtb_untile_device: shards [world, stride] uint8 -> frame [h, w, 3]."""
    frame = np.zeros((height, width, 3), dtype=np.uint8)
    for r in range(world):
        xy = tile_map(width, height, r, world)
        px = shards[r, : xy.shape[0] * 3].reshape(-1, 3)
        ok = xy[:, 0] >= 0
        frame[xy[ok, 1], xy[ok, 0]] = px[ok]
    return frame


def gather_frame_cpu(local_shard: np.ndarray, width: int, height: int, group=None) -> np.ndarray:
    """all_gather of tile-ordered shards on CPU tensors (gloo) + numpy untile; every rank gets the frame."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    stride = shard_stride(width, height, world)
    mine = torch.zeros(stride, dtype=torch.uint8)
    mine[: local_shard.size] = torch.from_numpy(np.ascontiguousarray(local_shard).reshape(-1))
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    return untile_numpy(torch.stack(parts).numpy(), width, height, world)


def broadcast_scene(path: str | None, assets_dir: str | None = None, *, device: int, src: int = 0, group=None) -> Scene:
    """The one-time scene + BVH broadcast of a multi-GPU job: rank `src` loads the TOML / OBJ files and builds the LBVH on its
    GPU, exports both (rtb_scene_export) and broadcasts the blob (NCCL on CUDA tensors; gloo works on CPU tensors for handles
    with device = -1); every other rank imports it (rtb_scene_import) — no parsing, no BVH build, bit-identical tables."""
    import torch
    import torch.distributed as dist

    rank = dist.get_rank(group)
    cuda = device >= 0 and dist.get_backend(group) == "nccl"
    tdev = torch.device("cuda", device) if cuda else torch.device("cpu")
    scene = None
    size = torch.zeros(1, dtype=torch.int64, device=tdev)
    if rank == src:
        scene = Scene.from_toml(path, assets_dir, device)
        blob = torch.from_numpy(scene.export())      # (a host-only handle exports the objects alone: importers build their own LBVH)
        size[0] = blob.numel()
    dist.broadcast(size, src, group=group)
    payload = blob.to(tdev) if rank == src else torch.empty(int(size.item()), dtype=torch.uint8, device=tdev)
    dist.broadcast(payload, src, group=group)
    if rank != src:
        scene = Scene.from_export(payload.cpu().numpy(), device, name=os.path.splitext(os.path.basename(path))[0] if path else "")
    return scene


def render_sharded(scene: Scene, width: int, height: int, spp: int, *, seed: int = 0, use_mis: bool = False,
                   pool_paths: int = 0, group=None):
    """Renders this rank's tiles on its GPU, all_gathers the RGB8 shards (NCCL) and returns the full
    scan-line frame as a CUDA uint8 tensor [h, w, 3] on every rank."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = torch.device("cuda", scene.info.device)
    stride = shard_stride(width, height, world)
    mine = torch.zeros(stride, dtype=torch.uint8, device=dev)
    p = make_params(width, height, spp, use_mis=use_mis, seed=seed, rank=rank, world=world, pool_paths=pool_paths)
    torch.cuda.synchronize(dev)
    scene.render_device(p, mine.data_ptr())
    gathered = torch.empty(world * stride, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(gathered, mine, group=group)
    frame = torch.empty((height, width, 3), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize(dev)
    _check(_abi.lib().rtb_untile_device(C.byref(p), C.c_void_p(gathered.data_ptr()), stride, C.c_void_p(frame.data_ptr()),
                                        scene.info.device))
    return frame
