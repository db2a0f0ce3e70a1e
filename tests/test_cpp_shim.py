"""The reference-facing C++ shim (include/cloud_merger_shim.hpp): builds tests/cpp/test_shim and runs it.
CPU box: the binary must refuse to compute (exit 77: no CUDA device, no CPU fallback). GPU box: it replays the reference's
per-frame call sequence (transformPointCloud, getROI, getCloudPart, fusePointclouds, voxelgrid, and the fused path) on the
GPU and compares every cloud bit for bit with the oracle."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "tests", "cpp")


def _build():
    from cloud_merger_b200 import build as cm_build
    from oracle import cm_oracle_py
    cm_build.build()
    cm_oracle_py.build()
    subprocess.check_call(["make", "-C", CPP, "-s"])
    return os.path.join(CPP, "test_shim")


def test_shim_builds_and_refuses_without_gpu():
    exe = _build()
    from cloud_merger_b200 import _lib
    if _lib.load().cm_device_count() > 0:
        pytest.skip("a GPU is present")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 77, r.stdout + r.stderr
    assert "no CPU path" in r.stdout


@pytest.mark.gpu
def test_shim_matches_oracle_on_gpu(gpu_ok):
    exe = _build()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "shim ok" in r.stdout
