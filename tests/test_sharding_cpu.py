"""Multi-GPU host logic on CPU: tile interleave, shard padding, gather + untile over a
world_size-2 (and 3) gloo process group."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_tiles_partition_the_frame(rtb):
    from raytracer_server_b200 import sharding as S

    for (w, h, world) in ((600, 450, 1), (600, 450, 2), (100, 70, 3), (3840, 2160, 8), (33, 31, 4), (5, 5, 8)):
        seen = np.zeros((h, w), dtype=np.int32)
        for r in range(world):
            xy = S.tile_map(w, h, r, world)
            assert xy.shape[0] == S.local_pixels(w, h, r, world) and xy.shape[0] % 1024 == 0
            ok = xy[:, 0] >= 0
            assert ((xy[ok, 0] < w) & (xy[ok, 1] < h)).all()
            np.add.at(seen, (xy[ok, 1], xy[ok, 0]), 1)
        assert (seen == 1).all()                      # every pixel owned by exactly one rank
        sizes = [S.local_pixels(w, h, r, world) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1024         # interleave balances to within one tile


def test_warp_sized_blocks_inside_a_tile(rtb):
    from raytracer_server_b200 import sharding as S

    xy = S.tile_map(64, 64, 0, 1)
    first = xy[:32]
    assert first[:, 0].max() - first[:, 0].min() == 7 and first[:, 1].max() - first[:, 1].min() == 3   # 8x4 pixels per warp


def _worker(rank, world, port, w, h, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import raytracer_server_b200  # noqa: F401
    from raytracer_server_b200 import sharding as S

    frame = np.random.default_rng(7).integers(0, 256, (h, w, 3), dtype=np.uint8)   # same on every rank
    xy = S.tile_map(w, h, rank, world)
    ok = xy[:, 0] >= 0
    shard = np.zeros((xy.shape[0], 3), dtype=np.uint8)
    shard[ok] = frame[xy[ok, 1], xy[ok, 0]]       # what this rank's GPU would have produced, tile order
    got = S.gather_frame_cpu(shard, w, h)
    ret[rank] = bool(np.array_equal(got, frame))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,w,h", [(2, 600, 450), (3, 100, 70)])
def test_gather_frame_gloo(world, w, h):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, w, h, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert all(ret.get(r) for r in range(world))


def _bcast_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import raytracer_server_b200  # noqa: F401
    from raytracer_server_b200 import sharding as S

    # only rank 0 touches the files; host-only handles (device = -1) ship the objects, a GPU rank would ship its LBVH too
    path = os.path.join(root, "tests", "golden", "scenes", "cubes.toml")
    sc = S.broadcast_scene(path if rank == 0 else None, device=-1, src=0)
    ret[rank] = (sc.info.n_objects, sc.info.n_triangles, sc.light_source, sc.object(6)["bb_min"], sc.octree_stats(6)["tri_refs"],
                 float(sc.triangles().sum()))
    dist.barrier()
    dist.destroy_process_group()


def test_broadcast_scene_gloo():
    # the scene (+ BVH) broadcast of a multi-GPU job, host side: rank 0 loads, everyone else imports the blob
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_bcast_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret[0] == ret[1] and ret[0][:3] == (9, 24, 8) and ret[0][4] == 44
