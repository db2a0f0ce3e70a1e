"""ctypes binding of oracle/libcm_oracle.so (TEST INFRASTRUCTURE ONLY; parity unpinned -- see cm_oracle.h).

Builds the library with oracle/Makefile on first use if it is missing.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class CmoPass(C.Structure):
    _fields_ = [("axis", C.c_int32), ("lo", C.c_float), ("hi", C.c_float), ("negative", C.c_int32)]


class CmoCloud(C.Structure):
    _fields_ = [("data", C.c_void_p), ("n_points", C.c_int64), ("point_step", C.c_int32), ("off_x", C.c_int32),
                ("off_y", C.c_int32), ("off_z", C.c_int32), ("off_i", C.c_int32), ("is_dense", C.c_int32),
                ("m", C.c_float * 12)]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libcm_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("cm_oracle.cpp", "cm_oracle.h", "Makefile")]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.cmo_version.restype = C.c_char_p
        L.cmo_passthrough.restype = C.c_int64
        L.cmo_radius_outlier.restype = C.c_int64
        L.cmo_voxelgrid.restype = C.c_int64
        L.cmo_merge_frame.restype = C.c_int64
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def unpack(data: np.ndarray, n: int, point_step: int, off_x: int, off_y: int, off_z: int, off_i: int) -> np.ndarray:
    data = np.ascontiguousarray(data).view(np.uint8).reshape(-1)
    out = np.empty((n, 4), np.float32)
    lib().cmo_unpack(_p(data), C.c_int64(n), point_step, off_x, off_y, off_z, off_i, _p(out))
    return out


def transform(xyzi: np.ndarray, m, is_dense: bool = True) -> np.ndarray:
    xyzi = np.ascontiguousarray(xyzi, np.float32)
    m = np.ascontiguousarray(m, np.float32).reshape(-1)[:12].copy()
    out = np.empty_like(xyzi)
    lib().cmo_transform(_p(xyzi), C.c_int64(len(xyzi)), _p(m), int(is_dense), _p(out))
    return out


def passthrough(xyzi: np.ndarray, axis: int, lo: float, hi: float, negative: bool = False) -> np.ndarray:
    xyzi = np.ascontiguousarray(xyzi, np.float32)
    idx = np.empty(len(xyzi), np.int32)
    k = lib().cmo_passthrough(_p(xyzi), C.c_int64(len(xyzi)), axis, C.c_float(lo), C.c_float(hi), int(negative), _p(idx))
    return idx[:k].copy()


def radius_outlier(xyzi: np.ndarray, radius: float, min_pts: int, negative: bool = False) -> np.ndarray:
    """pcl::RadiusOutlierRemoval as configured by outlierRemoval() (pc_preprocessing_main.cpp:184-192): indices kept."""
    xyzi = np.ascontiguousarray(xyzi, np.float32).reshape(-1, 4)
    idx = np.empty(max(len(xyzi), 1), np.int32)
    k = lib().cmo_radius_outlier(_p(xyzi), C.c_int64(len(xyzi)), C.c_double(radius), int(min_pts), int(negative), _p(idx))
    return idx[:k].copy()


def mt19937_at(seed: int, i: int) -> int:
    f = lib().cmo_mt19937_at
    f.restype = C.c_uint32
    return int(f(C.c_uint32(seed), C.c_int64(i)))


def plane_score(xyzi: np.ndarray, sample, threshold: float, sum_order: int = 0):
    """One RANSAC hypothesis (three point indices): (coefficients [4] float32, inlier count or -1 for a bad sample)."""
    xyzi = np.ascontiguousarray(xyzi, np.float32).reshape(-1, 4)
    smp = np.ascontiguousarray(sample, np.int32)
    coeff = np.zeros(4, np.float32)
    f = lib().cmo_plane_score
    f.restype = C.c_int64
    k = f(_p(xyzi), C.c_int64(len(xyzi)), _p(smp), C.c_double(threshold), int(sum_order), _p(coeff))
    return coeff, int(k)


def plane_ransac(xyzi: np.ndarray, threshold: float, probability: float = 0.99, max_iterations: int = 1000,
                 optimize: bool = True, seed: int = 12345, sum_order: int = 0) -> dict:
    """pcl::SACSegmentation (SACMODEL_PLANE, SAC_RANSAC) as removeGround() runs it (pc_preprocessing_main.cpp:95-108)."""
    xyzi = np.ascontiguousarray(xyzi, np.float32).reshape(-1, 4)
    n = len(xyzi)
    info = np.zeros(7, np.int32)
    c_r = np.zeros(4, np.float32)
    c_o = np.zeros(4, np.float32)
    inl = np.empty(max(n, 1), np.int32)
    f = lib().cmo_plane_ransac
    f.restype = C.c_int64
    k = f(_p(xyzi), C.c_int64(n), C.c_double(threshold), C.c_double(probability), int(max_iterations), int(bool(optimize)),
          C.c_uint32(seed), int(sum_order), _p(info), _p(c_r), _p(c_o), _p(inl))
    return {"found": bool(info[0]), "iterations": int(info[1]), "draws": int(info[2]), "best_count": int(info[3]),
            "sample": info[4:7].copy(), "coeff_ransac": c_r, "coeff": c_o, "inliers": inl[:k].copy()}


def zone_split(xyzi: np.ndarray, zones) -> list:
    """The reference's per-zone sequence, literally: for every zone run its PassThrough stages one after the other, each
    on the cloud the previous one copied out (getCloudPart followed by the z window of removeGround,
    pc_preprocessing_main.cpp:49-59, 80-92). zones: list of chains [(axis, lo, hi, negative), ...].
    Returns, per zone, (xyzi [k,4], src [k] = indices into the input cloud, in input order)."""
    xyzi = np.ascontiguousarray(xyzi, np.float32).reshape(-1, 4)
    out = []
    for chain in zones:
        cur = xyzi
        src = np.arange(len(xyzi), dtype=np.int64)
        for (axis, lo, hi, neg) in chain:
            keep = passthrough(cur, int(axis), float(np.float32(lo)), float(np.float32(hi)), bool(neg)).astype(np.int64)
            cur = np.ascontiguousarray(cur[keep])
            src = src[keep]
        out.append((cur, src.astype(np.uint32)))
    return out


def voxelgrid(xyzi: np.ndarray, leaf, min_points: int = 0, downsample_all: bool = True, force64: bool = True,
              is_dense: bool = True) -> dict:
    xyzi = np.ascontiguousarray(xyzi, np.float32)
    n = len(xyzi)
    leaf = np.ascontiguousarray(leaf, np.float32)
    cap = max(n, 1)
    o = np.empty((cap, 4), np.float32); os_ = np.empty((cap, 4), np.float32); od = np.empty((cap, 4), np.float64)
    cnt = np.empty(cap, np.uint32); idx = np.empty(cap, np.int64); pidx = np.empty(cap, np.int64)
    grid = np.zeros(9, np.int32); flags = np.zeros(1, np.int32)
    v = lib().cmo_voxelgrid(_p(xyzi), C.c_int64(n), int(is_dense), _p(leaf), C.c_uint32(min_points), int(downsample_all),
                            int(force64), _p(o), _p(os_), _p(od), _p(cnt), _p(idx), _p(pidx), _p(grid), _p(flags))
    overflow = bool(flags[0] & 1)
    if overflow and not force64:
        return dict(n=v, centroid=o[:v].copy(), pcl_overflow=True, returned_input=True, min_b=grid[0:3].copy(),
                    max_b=grid[3:6].copy(), div_b=grid[6:9].copy())
    return dict(n=v, centroid=o[:v].copy(), centroid_sort=os_[:v].copy(), centroid_f64=od[:v].copy(),
                count=cnt[:v].copy(), idx=idx[:v].copy(), point_idx=pidx[:n].copy(), min_b=grid[0:3].copy(),
                max_b=grid[3:6].copy(), div_b=grid[6:9].copy(), pcl_overflow=overflow, returned_input=False)


def make_clouds(clouds):
    """clouds: list of dict(data(np.uint8 buffer), n_points, point_step, off_x, off_y, off_z, off_i, is_dense, m)."""
    arr = (CmoCloud * len(clouds))()
    keep = []
    for i, c in enumerate(clouds):
        d = np.ascontiguousarray(c["data"]).view(np.uint8).reshape(-1)
        keep.append(d)
        arr[i].data = d.ctypes.data
        arr[i].n_points = int(c["n_points"]); arr[i].point_step = int(c["point_step"])
        arr[i].off_x = int(c["off_x"]); arr[i].off_y = int(c["off_y"]); arr[i].off_z = int(c["off_z"])
        arr[i].off_i = int(c["off_i"]); arr[i].is_dense = int(c["is_dense"])
        m = np.asarray(c["m"], np.float32).reshape(-1)[:12]
        for k in range(12):
            arr[i].m[k] = float(m[k])
    return arr, keep


def make_passes(passes):
    arr = (CmoPass * max(len(passes), 1))()
    for i, (axis, lo, hi, neg) in enumerate(passes):
        arr[i].axis = int(axis); arr[i].lo = float(np.float32(lo)); arr[i].hi = float(np.float32(hi))
        arr[i].negative = int(neg)
    return arr


def merge_frame(clouds, passes, leaf, min_points: int = 0, downsample_all: bool = True, force64: bool = True,
                threads: int = 1, want_outputs: bool = True) -> dict:
    """Whole reference path on the CPU. Returns dict(survivor_xyzi, survivor_src, n_survivors, voxel outputs...)."""
    carr, keep = make_clouds(clouds)
    parr = make_passes(passes)
    total = int(sum(int(c["n_points"]) for c in clouds))
    cap = max(total, 1)
    leaf = np.ascontiguousarray(leaf, np.float32)
    nsurv = C.c_int64(0)
    grid = np.zeros(9, np.int32); flags = np.zeros(1, np.int32)
    if want_outputs:
        sx = np.empty((cap, 4), np.float32); ss = np.empty(cap, np.uint32)
        o = np.empty((cap, 4), np.float32); od = np.empty((cap, 4), np.float64)
        cnt = np.empty(cap, np.uint32); idx = np.empty(cap, np.int64); pidx = np.empty(cap, np.int64)
    else:
        sx = ss = od = cnt = idx = pidx = None
        o = np.empty((cap, 4), np.float32)
    v = lib().cmo_merge_frame(carr, len(clouds), parr, len(passes), _p(leaf), C.c_uint32(min_points), int(downsample_all),
                              int(force64), int(threads), _p(sx), _p(ss), C.byref(nsurv), _p(o), _p(od), _p(cnt), _p(idx),
                              _p(pidx), _p(grid), _p(flags))
    m = nsurv.value
    res = dict(n_voxels=v, n_survivors=m, pcl_overflow=bool(flags[0] & 1), min_b=grid[0:3].copy(),
               max_b=grid[3:6].copy(), div_b=grid[6:9].copy())
    res["returned_input"] = res["pcl_overflow"] and not force64
    if want_outputs:
        res.update(survivor_xyzi=sx[:m].copy(), survivor_src=ss[:m].copy(), centroid=o[:v].copy())
        if not res["returned_input"]:
            res.update(centroid_f64=od[:v].copy(), count=cnt[:v].copy(), idx=idx[:v].copy(), point_idx=pidx[:m].copy())
    return res


def tf_to_matrix(quat_xyzw, origin_xyz) -> np.ndarray:
    q = np.ascontiguousarray(quat_xyzw, np.float64); t = np.ascontiguousarray(origin_xyz, np.float64)
    m = np.empty(12, np.float32)
    lib().cmo_tf_to_matrix(_p(q), _p(t), _p(m))
    return m
