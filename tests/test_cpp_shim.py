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


def test_shim_handoff_is_race_free_under_tsan():
    """FusedFrame's hand-off between the six callback threads and the main loop (include/cloud_merger_shim.hpp), built with
    -fsanitize=thread against a host-only stub of the C ABI (tests/cpp/stub_cm.cpp): ThreadSanitizer reports no race (it
    would exit 66), every fused frame holds the required sensors, no delivery is lost. CPU only: CUDA does not run under TSAN."""
    _build()
    exe = os.path.join(CPP, "test_shim_threads")
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=1 exitcode=66")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "shim threads ok" in r.stdout and "ThreadSanitizer" not in r.stderr
