"""ctypes binding of oracle/_ref/libcm_pcl_ref.so -- the REAL PCL behind the C wrapper of oracle/pcl_ref/cm_pcl_ref.cpp
(TEST INFRASTRUCTURE ONLY). available() is False wherever PCL could not be built (this repository's own image)."""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.normpath(os.path.join(_HERE, "..", "_ref"))
SO = os.path.join(REF_DIR, "libcm_pcl_ref.so")
_LIB = None
_WHY = ""


def try_build() -> bool:
    """Configure + build with cmake when PCL can be found; quiet no-op otherwise. Never raises."""
    global _WHY
    if os.path.exists(SO):
        return True
    if not shutil.which("cmake"):
        _WHY = "cmake not found"
        return False
    bdir = os.path.join(REF_DIR, "build")
    os.makedirs(bdir, exist_ok=True)
    r = subprocess.run(["cmake", "-S", _HERE, "-B", bdir], capture_output=True, text=True)
    if r.returncode != 0:
        _WHY = "PCL >= 1.8 not found by cmake (find_package(PCL) failed)"
        return False
    r = subprocess.run(["cmake", "--build", bdir, "--target", "cm_pcl_ref"], capture_output=True, text=True)
    if r.returncode != 0 or not os.path.exists(SO):
        _WHY = "building oracle/pcl_ref failed: " + (r.stdout + r.stderr)[-400:]
        return False
    return True


def available() -> bool:
    return os.path.exists(SO) or try_build()


def why_not() -> str:
    return _WHY or "oracle/_ref/libcm_pcl_ref.so is missing (build it with the recipe in oracle/pcl_ref/README.md)"


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        L = C.CDLL(SO)
        L.cmp_version.restype = C.c_char_p
        for f in ("cmp_passthrough", "cmp_concat", "cmp_voxelgrid", "cmp_radius_outlier", "cmp_plane_ransac"):
            getattr(L, f).restype = C.c_int64
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def version() -> str:
    return lib().cmp_version().decode()


def tf_to_matrix(q, t) -> np.ndarray:
    q = np.ascontiguousarray(q, np.float64); t = np.ascontiguousarray(t, np.float64)
    m = np.empty(12, np.float32)
    lib().cmp_tf_to_matrix(_p(q), _p(t), _p(m))
    return m


def transform(xyzi, m, is_dense=True) -> np.ndarray:
    xyzi = np.ascontiguousarray(xyzi, np.float32)
    m = np.ascontiguousarray(m, np.float32).reshape(-1)[:12].copy()
    out = np.empty_like(xyzi)
    lib().cmp_transform(_p(xyzi), C.c_int64(len(xyzi)), _p(m), int(is_dense), _p(out))
    return out


def passthrough(xyzi, axis, lo, hi, negative=False) -> np.ndarray:
    xyzi = np.ascontiguousarray(xyzi, np.float32)
    idx = np.empty(max(len(xyzi), 1), np.int32)
    k = lib().cmp_passthrough(_p(xyzi), C.c_int64(len(xyzi)), int(axis), C.c_float(lo), C.c_float(hi), int(negative), _p(idx))
    return idx[:k].copy()


def concat(a, a_dense, a_stamp, b, b_dense, b_stamp):
    a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32)
    out = np.empty((len(a) + len(b), 4), np.float32)
    stamp = C.c_uint64(0)
    meta = np.zeros(3, np.int32)
    n = lib().cmp_concat(_p(a), C.c_int64(len(a)), int(a_dense), C.c_uint64(a_stamp), _p(b), C.c_int64(len(b)), int(b_dense),
                         C.c_uint64(b_stamp), _p(out), C.byref(stamp), _p(meta))
    return out[:n], int(stamp.value), meta


def voxelgrid(xyzi, leaf, min_points, downsample_all=True, is_dense=True):
    xyzi = np.ascontiguousarray(xyzi, np.float32)
    leaf = np.ascontiguousarray(np.broadcast_to(np.asarray(leaf, np.float32), (3,)))
    out = np.empty((max(len(xyzi), 1), 4), np.float32)
    grid = np.zeros(9, np.int32)
    v = lib().cmp_voxelgrid(_p(xyzi), C.c_int64(len(xyzi)), int(is_dense), _p(leaf), C.c_uint32(min_points), int(downsample_all),
                            _p(out), _p(grid))
    return out[:v].copy(), grid


def radius_outlier(xyzi, radius, min_pts, negative=False) -> np.ndarray:
    xyzi = np.ascontiguousarray(xyzi, np.float32)
    idx = np.empty(max(len(xyzi), 1), np.int32)
    k = lib().cmp_radius_outlier(_p(xyzi), C.c_int64(len(xyzi)), C.c_double(radius), int(min_pts), int(negative), _p(idx))
    return idx[:k].copy()


def plane_ransac(xyzi, threshold, probability=0.99, max_iterations=1000, optimize=True, eps_angle=0.05):
    xyzi = np.ascontiguousarray(xyzi, np.float32)
    coeff = np.zeros(4, np.float32)
    inl = np.empty(max(len(xyzi), 1), np.int32)
    k = lib().cmp_plane_ransac(_p(xyzi), C.c_int64(len(xyzi)), C.c_double(threshold), C.c_double(probability), int(max_iterations),
                               int(bool(optimize)), C.c_float(eps_angle), _p(coeff), _p(inl))
    return coeff, inl[:k].copy()
