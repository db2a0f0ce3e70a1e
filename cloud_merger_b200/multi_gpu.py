"""Multi-GPU drivers of the merge hot path (one process per GPU, torch.distributed for the plumbing only).

Two modes:

* frame-sharded stream (BASELINE configs 2 and 3): frames are independent, so frame f goes to rank f mod G and there is
  NO data-path collective; only a small gather of per-frame summaries re-orders the results on rank 0.
* single giant cloud (config 4): the cloud is block-distributed; every rank builds the SAME voxel grid from the
  all-reduced bounding box, a coarse histogram of the voxel keys is all-reduced to pick G-1 balanced splitters, one
  all-to-all moves every point to the rank owning its key range (a voxel never straddles ranks), and each rank then
  runs the ordinary single-GPU VoxelGrid. Concatenating the rank outputs in rank order is the global PCL order.

The per-rank compute is injected as a callable so that the same orchestration runs on NCCL with the CUDA library and,
in the CPU tests, on gloo with a checker backend; with a CudaRouter the routing steps (bounding box, key histogram,
grouping by destination) run on this package's kernels too and torch only carries the collectives.
"""
from __future__ import annotations

import sys
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

F32 = np.float32


# ---- frame-sharded stream ---------------------------------------------------------------------------------------------------
def frames_of_rank(n_frames: int, rank: int, world: int) -> List[int]:
    """Round-robin ownership: frame f is merged by rank f mod world."""
    return list(range(rank, n_frames, world))


def run_frame_sharded(n_frames: int, merge_frame: Callable[[int], dict], rank: int, world: int, group=None) -> Optional[List[dict]]:
    """Every rank merges its own frames with `merge_frame(f) -> summary dict`; rank 0 gets the summaries of all frames in
    frame order (no point data crosses ranks)."""
    mine = [(f, merge_frame(f)) for f in frames_of_rank(n_frames, rank, world)]
    if world == 1:
        return [s for _, s in sorted(mine, key=lambda t: t[0])]
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(mine, gathered, dst=0, group=group)
    if rank != 0:
        return None
    flat = [item for part in gathered for item in part]
    flat.sort(key=lambda t: t[0])
    assert [f for f, _ in flat] == list(range(n_frames)), "every frame must be merged exactly once"
    return [s for _, s in flat]


# ---- single giant cloud: voxel-key range partition -----------------------------------------------------------------------------
def global_grid(local_xyzi: torch.Tensor, leaf: Sequence[float], group=None) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """All-reduced bounding box (pcl::getMinMax3D over the whole cloud) and the grid PCL derives from it.
    Returns (min_p, max_p, min_b, div_b); float32 arithmetic as in PCL 1.8.1 VoxelGrid::applyFilter."""
    xyz = local_xyzi[:, :3]
    fin = torch.isfinite(xyz).all(dim=1)
    big = torch.finfo(torch.float32).max
    if bool(fin.any()):
        v = xyz[fin]
        mn, mx = v.min(dim=0).values, v.max(dim=0).values
    else:
        mn = torch.full((3,), big, dtype=torch.float32, device=xyz.device)
        mx = -mn
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(mn, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    min_p, max_p = mn.cpu().numpy().astype(F32), mx.cpu().numpy().astype(F32)
    inv = (F32(1.0) / np.asarray(leaf, F32)).astype(F32)
    min_b = np.floor((min_p * inv).astype(F32)).astype(np.int64)
    max_b = np.floor((max_p * inv).astype(F32)).astype(np.int64)
    return min_p, max_p, min_b, max_b - min_b + 1


def voxel_keys(xyzi: torch.Tensor, leaf: Sequence[float], min_b: np.ndarray, div_b: np.ndarray) -> torch.Tensor:
    """64-bit voxel index idx = i + j*div_x + k*div_x*div_y of every point (float32 multiply, floor, as PCL); -1 for
    non-finite points. Used only to ROUTE points; the owning rank recomputes the keys in its CUDA kernels."""
    inv = torch.tensor((F32(1.0) / np.asarray(leaf, F32)).astype(F32), device=xyzi.device)
    xyz = xyzi[:, :3]
    fin = torch.isfinite(xyz).all(dim=1)
    cell = torch.floor(torch.where(fin[:, None], xyz, torch.zeros_like(xyz)) * inv).to(torch.int64)
    cell = cell - torch.tensor(min_b, device=xyzi.device, dtype=torch.int64)
    d0, d1 = int(div_b[0]), int(div_b[1])
    key = cell[:, 0] + cell[:, 1] * d0 + cell[:, 2] * (d0 * d1)
    return torch.where(fin, key, torch.full_like(key, -1))


def pick_splitters(keys: torch.Tensor, n_cells: int, world: int, bins: int = 1 << 16, group=None) -> torch.Tensor:
    """G-1 key splitters that balance the point counts: all-reduce a coarse histogram of the keys (bins of equal key
    width), cut it at multiples of total/G. Returns int64 keys s_1 < ... ; rank r owns keys in [s_r, s_{r+1})."""
    width = max(1, -(-int(n_cells) // bins))
    valid = keys >= 0
    hist = torch.bincount((keys[valid] // width).clamp_(0, bins - 1), minlength=bins).to(torch.int64)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(hist, group=group)
    cum = torch.cumsum(hist, dim=0)
    total = int(cum[-1].item())
    targets = torch.tensor([total * r // world for r in range(1, world)], device=keys.device, dtype=torch.int64)
    cut_bins = torch.searchsorted(cum, targets, right=False) + 1  # first bin boundary at or after the target
    return (cut_bins.clamp_(0, bins) * width).to(torch.int64)


def exchange_by_key_range(local_xyzi: torch.Tensor, keys: torch.Tensor, splitters: torch.Tensor, rank: int, world: int,
                          group=None) -> torch.Tensor:
    """One all-to-all: every point goes to the rank owning its key range (invalid points, key -1, stay where they are:
    VoxelGrid skips them). Within a destination the source order is kept."""
    if world == 1:
        return local_xyzi
    dest = torch.searchsorted(splitters, keys, right=True)
    dest = torch.where(keys < 0, torch.full_like(dest, rank), dest)
    order = torch.argsort(dest, stable=True)
    send = local_xyzi[order].contiguous()
    send_counts = torch.bincount(dest, minlength=world).to(torch.int64)
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts, group=group)
    sc, rc = send_counts.tolist(), recv_counts.tolist()
    recv = torch.empty((int(sum(rc)), 4), dtype=local_xyzi.dtype, device=local_xyzi.device)
    dist.all_to_all_single(recv, send, output_split_sizes=rc, input_split_sizes=sc, group=group)
    return recv


class CudaRouter:
    """The routing steps of giant_cloud_voxelgrid on this package's CUDA kernels instead of torch ops: bounding box
    (cm_dev_bounds), coarse key histogram (cm_dev_key_histogram), grouping by destination (cm_dev_route_by_key, the
    zone-slicing count / scan / scatter). torch only carries the collectives. The leaf must have been set on the merger
    (cm_set_voxel) -- cuda_voxelgrid_backend does that."""

    def __init__(self, merger, bins: int = 1 << 14):
        self.m = merger
        self.bins = bins

    def global_grid(self, local_xyzi: torch.Tensor, leaf: Sequence[float], group=None):
        stream = torch.cuda.current_stream().cuda_stream
        mn, mx, _ = self.m.dev_bounds(local_xyzi.data_ptr(), int(local_xyzi.shape[0]), stream=stream)
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            t = torch.tensor(np.concatenate([mn, -mx]), device=local_xyzi.device)  # one MIN all-reduce for both bounds
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
            t = t.cpu().numpy()
            mn, mx = t[:3].astype(F32), (-t[3:]).astype(F32)
        inv = (F32(1.0) / np.asarray(leaf, F32)).astype(F32)
        min_b = np.floor((mn * inv).astype(F32)).astype(np.int64)
        max_b = np.floor((mx * inv).astype(F32)).astype(np.int64)
        return mn, mx, min_b, max_b - min_b + 1

    def exchange(self, local_xyzi: torch.Tensor, min_p, max_p, rank: int, world: int, group=None):
        """histogram -> all-reduce -> splitters -> group by destination -> one all-to-all. Returns (received points,
        points sent to other ranks, splitters). With CM_GIANT_TRACE=1 rank 0 prints a per-step time breakdown."""
        import os
        import time
        trace = os.environ.get("CM_GIANT_TRACE") and rank == 0
        marks = []

        def mark(name):
            if trace:
                torch.cuda.synchronize()
                marks.append((name, time.perf_counter()))
        mark("start")
        dev = local_xyzi.device
        n = int(local_xyzi.shape[0])
        stream = torch.cuda.current_stream().cuda_stream
        hist = torch.empty(self.bins, dtype=torch.int64, device=dev)
        width = self.m.dev_key_histogram(local_xyzi.data_ptr(), n, min_p, max_p, self.bins, hist.data_ptr(), stream=stream)
        mark("key histogram")
        dist.all_reduce(hist, group=group)
        cum = torch.cumsum(hist, dim=0)
        total = int(cum[-1].item())
        targets = torch.tensor([total * r // world for r in range(1, world)], device=dev, dtype=torch.int64)
        cut_bins = torch.searchsorted(cum, targets, right=False) + 1
        splitters = (cut_bins.clamp_(0, self.bins) * width).to(torch.int64)
        sp = splitters.cpu().tolist()
        mark("all-reduce + splitters")
        self.m.dev_route_by_key(local_xyzi.data_ptr(), n, min_p, max_p, sp, rank, stream=stream)
        xyzi_ptr, _, begin = self.m.zone_out_raw()
        mark("group by destination")
        send = torch.empty((begin[-1], 4), dtype=torch.float32, device=dev)
        self.m.memcpy_d2d(send.data_ptr(), xyzi_ptr, begin[-1] * 16, stream=stream)
        mark("copy to send buffer")
        sc = [begin[r + 1] - begin[r] for r in range(world)]
        send_counts = torch.tensor(sc, dtype=torch.int64, device=dev)
        recv_counts = torch.empty_like(send_counts)
        dist.all_to_all_single(recv_counts, send_counts, group=group)
        rc = recv_counts.tolist()
        recv = torch.empty((int(sum(rc)), 4), dtype=torch.float32, device=dev)
        mark("counts exchange")
        dist.all_to_all_single(recv, send, output_split_sizes=rc, input_split_sizes=sc, group=group)
        mark("all-to-all")
        if trace:
            print("[giant] " + ", ".join("%s %.2f ms" % (b[0], (b[1] - a[1]) * 1e3) for a, b in zip(marks, marks[1:])), file=sys.stderr)
        return recv, n - sc[rank], splitters


def giant_cloud_voxelgrid(local_xyzi: torch.Tensor, leaf: Sequence[float], min_points: int,
                          voxelgrid_local: Callable[[torch.Tensor, np.ndarray, np.ndarray], dict], rank: int, world: int,
                          group=None, router: Optional[CudaRouter] = None) -> dict:
    """VoxelGrid of one cloud spread over `world` ranks. `voxelgrid_local(points, min_p, max_p) -> dict(idx, count,
    centroid)` is the single-GPU VoxelGrid run with the GLOBAL bounding box (cm_set_voxel_bounds + cm_dev_voxelgrid).
    Returns this rank's voxels (ascending idx; ranks hold ascending, disjoint key ranges) plus exchange statistics."""
    if router is not None:  # the product path: CUDA kernels of this package, torch for the collectives only
        min_p, max_p, min_b, div_b = router.global_grid(local_xyzi, leaf, group)
        if world > 1:
            recv, sent_away, splitters = router.exchange(local_xyzi, min_p, max_p, rank, world, group)
        else:
            splitters, recv, sent_away = torch.zeros(0, dtype=torch.int64), local_xyzi, 0
        out = voxelgrid_local(recv, min_p, max_p)
        out.update(points_received=int(recv.shape[0]), points_sent_away=sent_away, min_b=min_b, div_b=div_b,
                   splitters=splitters.cpu().numpy())
        return out
    min_p, max_p, min_b, div_b = global_grid(local_xyzi, leaf, group)
    n_cells = int(div_b[0]) * int(div_b[1]) * int(div_b[2])
    if world > 1:
        keys = voxel_keys(local_xyzi, leaf, min_b, div_b)
        splitters = pick_splitters(keys, n_cells, world, group=group).to(keys.device)
        recv = exchange_by_key_range(local_xyzi, keys, splitters, rank, world, group)
        dest = torch.searchsorted(splitters, keys, right=True)
        sent_away = int(((dest != rank) & (keys >= 0)).sum().item())
    else:
        splitters, recv, sent_away = torch.zeros(0, dtype=torch.int64), local_xyzi, 0
    out = voxelgrid_local(recv, min_p, max_p)
    out.update(points_received=int(recv.shape[0]), points_sent_away=sent_away, min_b=min_b, div_b=div_b,
               splitters=splitters.cpu().numpy())
    return out


def cuda_voxelgrid_backend(merger, leaf, min_points: int, download: bool = True):
    """The product backend of giant_cloud_voxelgrid: the CUDA VoxelGrid of this package on the received points.
    download=False leaves the voxels on the device (cm_get_device_out) and returns only their number."""
    merger.set_voxel(leaf, min_points, True)

    def run(points: torch.Tensor, min_p: np.ndarray, max_p: np.ndarray) -> dict:
        assert points.is_cuda and points.dtype == torch.float32 and points.is_contiguous()
        merger.set_voxel_bounds(min_p, max_p)
        stream = torch.cuda.current_stream().cuda_stream
        merger.dev_voxelgrid(points.data_ptr(), int(points.shape[0]), True, stream=stream)
        st = merger.stats()
        v = int(st.voxels_out)
        res = dict(n_voxels=v, gpu_ms=float(st.gpu_ms), key_bits=int(st.key_bits), idx=np.zeros(0, np.int64),
                   count=np.zeros(0, np.uint32), centroid=np.zeros((0, 4), np.float32))
        if download:
            o = merger.device_out()
            step_f = merger.out_point_step // 4
            cen = merger.download(o.voxel_xyzi, np.float32, v * step_f).reshape(v, step_f)
            res.update(idx=merger.download(o.voxel_idx, np.uint64, v).astype(np.int64),
                       count=merger.download(o.voxel_count, np.uint32, v),
                       centroid=cen[:, :4] if step_f == 4 else np.concatenate([cen[:, 0:3], cen[:, 4:5]], axis=1))
        return res
    return run
