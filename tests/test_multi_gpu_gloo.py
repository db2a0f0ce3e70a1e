"""World-size-2 CPU tests (gloo) of the multi-GPU orchestration in cloud_merger_b200/multi_gpu.py: frame sharding with no
data-path collective, and the single-giant-cloud voxel-key range partition with its one all-to-all. The per-rank compute
is injected; here it is the numpy oracle (tests only), on the GPU box it is the CUDA library (test_multi_gpu_nccl)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cloud_merger_b200 import multi_gpu, synth
from oracle import np_oracle as npo

from helpers import cloud_dict


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _init(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)


def _frame_summary(f):
    clouds, mats = synth.frame_clouds("cfg1", 1000, f)
    clouds = [c[:4096] for c in clouds]
    cds = [cloud_dict(p, m[:3]) for p, m in zip(clouds, mats)]
    r = npo.merge_frame(cds, synth.ROI_BOX, [0.1] * 3, 2, True, True)
    return dict(frame=f, survivors=len(r["survivor_src"]), voxels=len(r["voxel"]["idx"]),
                checksum=int(r["voxel"]["idx"].sum() % (1 << 61)))


def _worker_frames(rank, world, port, n_frames, q):
    _init(rank, world, port)
    out = multi_gpu.run_frame_sharded(n_frames, _frame_summary, rank, world)
    if rank == 0:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_frame_sharding_world2():
    n_frames, world = 5, 2
    assert multi_gpu.frames_of_rank(n_frames, 0, world) == [0, 2, 4] and multi_gpu.frames_of_rank(n_frames, 1, world) == [1, 3]
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_frames, args=(r, world, port, n_frames, q)) for r in range(world)]
    [p.start() for p in procs]
    got = q.get()
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    want = [_frame_summary(f) for f in range(n_frames)]
    assert got == want


def _np_backend(leaf, min_points):
    def run(points, min_p, max_p):
        r = npo.voxelgrid(points.cpu().numpy(), leaf, min_points, True, True, bounds=(min_p, max_p))
        return dict(idx=r["idx"], count=r["count"], centroid=r["centroid_f64"], n_voxels=len(r["idx"]))
    return run


def _worker_giant(rank, world, port, n, leaf, min_points, q):
    _init(rank, world, port)
    whole = synth.map_cloud(4, n, extent=(60.0, 60.0, 6.0), n_boxes=40)
    whole[::997, 0] = np.nan                      # a few invalid points: VoxelGrid skips them
    lo, hi = rank * n // world, (rank + 1) * n // world   # block distribution
    local = torch.from_numpy(whole[lo:hi].copy())
    out = multi_gpu.giant_cloud_voxelgrid(local, [leaf] * 3, min_points, _np_backend([leaf] * 3, min_points), rank, world)
    q.put((rank, out["idx"], out["count"], out["centroid"], out["points_received"], out["points_sent_away"]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("leaf,min_points", [(0.25, 1), (0.5, 2)])
def test_giant_cloud_partition_world2(leaf, min_points):
    n, world = 60000, 2
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_giant, args=(r, world, port, n, leaf, min_points, q)) for r in range(world)]
    [p.start() for p in procs]
    parts = sorted([q.get() for _ in range(world)], key=lambda t: t[0])
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    whole = synth.map_cloud(4, n, extent=(60.0, 60.0, 6.0), n_boxes=40)
    whole[::997, 0] = np.nan
    ref = npo.voxelgrid(whole, [leaf] * 3, min_points, True, True)
    idx = np.concatenate([p[1] for p in parts])
    cnt = np.concatenate([p[2] for p in parts])
    cen = np.concatenate([p[3] for p in parts])
    assert (np.diff(idx) > 0).all(), "rank order must be global voxel order, no voxel on two ranks"
    assert (idx == ref["idx"]).all() and (cnt == ref["count"]).all()
    np.testing.assert_allclose(cen, ref["centroid_f64"], rtol=1e-9, atol=1e-9)
    received = [p[4] for p in parts]
    assert sum(received) == n and min(received) > 0.3 * n, "splitters must balance the ranks: %r" % received
    assert sum(p[5] for p in parts) > 0, "some points must have crossed ranks"


def test_splitters_and_keys_single_process():
    x = torch.from_numpy(synth.uniform_cloud(3, 20000, extent=(40.0, 40.0, 4.0)))
    min_p, max_p, min_b, div_b = multi_gpu.global_grid(x, [0.5] * 3)
    keys = multi_gpu.voxel_keys(x, [0.5] * 3, min_b, div_b)
    ref = npo.voxelgrid(x.numpy(), [0.5] * 3, 1, True, True)
    assert (keys.numpy() == ref["point_idx"]).all()
    assert min_b.tolist() == ref["min_b"].tolist() and div_b.tolist() == ref["div_b"].tolist()
    sp = multi_gpu.pick_splitters(keys, int(np.prod(div_b)), 4)
    dest = torch.searchsorted(sp, keys, right=True)
    counts = torch.bincount(dest, minlength=4).numpy()
    assert counts.sum() == len(x) and counts.min() > 0.15 * len(x), counts
