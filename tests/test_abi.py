"""CPU tests of the drop-in boundary (not gpu): the C-ABI library loads, exports every symbol include/cloud_merger_gpu.h
declares, its structs have the sizes the ctypes mirror assumes, and -- with no GPU -- it refuses loudly instead of
falling back to a CPU path. No compute call is made here."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "cloud_merger_gpu.h")


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"CM_API\s+[\w\s\*]+?\b(cm_\w+)\s*\(", text)))


def test_header_declares_a_sane_surface():
    names = declared_symbols()
    for must in ("cm_create", "cm_destroy", "cm_set_extrinsic", "cm_set_crop", "cm_set_voxel", "cm_submit_cloud",
                 "cm_merge_frame", "cm_run_batch", "cm_dev_transform_crop", "cm_dev_voxelgrid", "cm_get_stats",
                 "cm_strerror"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from cloud_merger_b200 import _lib
    lib = _lib.load()
    names = declared_symbols()
    assert sorted(_lib.SYMBOLS) == names, "ctypes table and header disagree"
    for n in names:
        assert hasattr(lib, n), n
    assert b"sm_100a" in lib.cm_version()
    assert lib.cm_strerror(_lib.CM_E_NO_DEVICE) == b"no usable CUDA device"


def test_struct_sizes_match_the_header():
    """Compile a tiny C program against the header and compare sizeof() with the ctypes mirror."""
    from cloud_merger_b200 import _lib
    src = r'''
#include <stdio.h>
#include "cloud_merger_gpu.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(cm_pass_t), sizeof(cm_layout_t), sizeof(cm_segment_t),
         sizeof(cm_config_t), sizeof(cm_stats_t), sizeof(cm_frame_info_t), sizeof(cm_device_out_t), sizeof(cm_frame_out_t),
         sizeof(cm_zone_t), sizeof(cm_zone_out_t), sizeof(cm_plane_cfg_t), sizeof(cm_plane_t), sizeof(cm_pc2_field_t),
         sizeof(cm_pc2_desc_t), sizeof(cm_proceed_cfg_t), sizeof(cm_proceed_out_t), sizeof(cm_giant_info_t), sizeof(cm_frame_view_t));
  return 0;
}'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(v) for v in subprocess.check_output([exe]).split()]
    mirror = [C.sizeof(t) for t in (_lib.CmPass, _lib.CmLayout, _lib.CmSegment, _lib.CmConfig, _lib.CmStats,
                                    _lib.CmFrameInfo, _lib.CmDeviceOut, _lib.CmFrameOut, _lib.CmZone, _lib.CmZoneOut,
                                    _lib.CmPlaneCfg, _lib.CmPlane, _lib.CmPc2Field, _lib.CmPc2Desc, _lib.CmProceedCfg,
                                    _lib.CmProceedOut, _lib.CmGiantInfo, _lib.CmFrameView)]
    assert sizes == mirror


def test_no_cpu_fallback_without_a_gpu():
    """Without a CUDA device cm_create must fail (CM_E_NO_DEVICE); the Python layer raises. With a GPU this is skipped."""
    from cloud_merger_b200 import CloudMerger, CloudMergerError, _lib
    lib = _lib.load()
    if lib.cm_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(CloudMergerError) as e:
        CloudMerger()
    assert e.value.code == _lib.CM_E_NO_DEVICE


def test_product_package_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under cloud_merger_b200/ may import, link or execute it."""
    pkg = os.path.join(ROOT, "cloud_merger_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "cm_oracle" not in text and "np_oracle" not in text and "from oracle" not in text, f
    for f in os.listdir(os.path.join(ROOT, "include")):
        text = open(os.path.join(ROOT, "include", f), errors="ignore").read()
        assert "cm_oracle" not in text, f


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    from cloud_merger_b200 import _lib
    monkeypatch.setattr(_lib, "_LIB", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()
