"""CPU tests of the sensor_msgs/PointCloud2 adapters of the C ABI (cm_layout_from_pointcloud2, cm_pointcloud2_describe):
pure host functions, so they run without a GPU. They replace the field lookup of the pcl_ros subscriber
(pc_preprocessing_main.cpp:520-525) and the header part of pcl::toROSMsg (:199-220)."""
import ctypes as C

import pytest

from cloud_merger_b200 import _lib
from cloud_merger_b200._lib import CM_PC2_FLOAT32, CmLayout, CmPc2Desc, CmPc2Field

F32, U16, U8, F64 = CM_PC2_FLOAT32, 4, 2, 8


def layout_of(fields, point_step, bigendian=0, dense=1):
    lib = _lib.load()
    arr = (CmPc2Field * max(len(fields), 1))()
    for i, (name, off, dt, cnt) in enumerate(fields):
        arr[i] = CmPc2Field(name.encode(), off, dt, cnt)
    out = CmLayout()
    rc = lib.cm_layout_from_pointcloud2(arr, len(fields), point_step, bigendian, dense, C.byref(out))
    return rc, out


@pytest.mark.parametrize("name,fields,step,want", [
    # velodyne_pointcloud of ROS Melodic: x y z pad intensity ring
    ("velodyne32", [("x", 0, F32, 1), ("y", 4, F32, 1), ("z", 8, F32, 1), ("intensity", 16, F32, 1), ("ring", 20, U16, 1)], 32,
     (32, 0, 4, 8, 16)),
    # newer velodyne driver: x y z intensity ring time
    ("velodyne22", [("x", 0, F32, 1), ("y", 4, F32, 1), ("z", 8, F32, 1), ("intensity", 12, F32, 1), ("ring", 16, U16, 1),
                    ("time", 18, F32, 1)], 22, (22, 0, 4, 8, 12)),
    # livox_ros_driver: x y z intensity tag line
    ("livox18", [("x", 0, F32, 1), ("y", 4, F32, 1), ("z", 8, F32, 1), ("intensity", 12, F32, 1), ("tag", 16, U8, 1),
                 ("line", 17, U8, 1)], 18, (18, 0, 4, 8, 12)),
    # field order in the message is irrelevant, count 0 reads as 1
    ("shuffled", [("intensity", 12, F32, 0), ("z", 8, F32, 1), ("ring", 16, U16, 1), ("x", 0, F32, 1), ("y", 4, F32, 1)], 20,
     (20, 0, 4, 8, 12)),
    # no intensity at all / an integer intensity: PCL leaves the field unmatched -> read as absent
    ("xyz_only", [("x", 0, F32, 1), ("y", 4, F32, 1), ("z", 8, F32, 1)], 12, (12, 0, 4, 8, -1)),
    ("u8_intensity", [("x", 0, F32, 1), ("y", 4, F32, 1), ("z", 8, F32, 1), ("intensity", 12, U8, 1)], 16, (16, 0, 4, 8, -1)),
])
def test_layout_from_pointcloud2(name, fields, step, want):
    rc, lay = layout_of(fields, step, dense=0)
    assert rc == _lib.CM_OK, name
    assert (lay.point_step, lay.off_x, lay.off_y, lay.off_z, lay.off_intensity) == want
    assert lay.is_dense == 0


def test_layout_from_pointcloud2_refuses_what_pcl_cannot_map():
    xyz = [("x", 0, F32, 1), ("y", 4, F32, 1), ("z", 8, F32, 1)]
    assert layout_of(xyz, 16, bigendian=1)[0] == _lib.CM_E_INVALID                      # big-endian data
    assert layout_of(xyz[:2], 16)[0] == _lib.CM_E_INVALID                                # no z
    assert layout_of([("x", 0, F64, 1)] + xyz[1:], 24)[0] == _lib.CM_E_INVALID           # FLOAT64 coordinate
    assert layout_of([("x", 0, F32, 3)] + xyz[1:], 16)[0] == _lib.CM_E_INVALID           # array field
    assert layout_of(xyz[:2] + [("z", 14, F32, 1)], 16)[0] == _lib.CM_E_INVALID          # field runs past point_step
    assert layout_of(xyz, 8)[0] == _lib.CM_E_INVALID                                     # point_step too small
    assert _lib.load().cm_layout_from_pointcloud2(None, 0, 16, 0, 1, None) == _lib.CM_E_INVALID


@pytest.mark.parametrize("step,off_i", [(32, 16), (16, 12)])
def test_pointcloud2_describe_matches_toROSMsg(step, off_i):
    """pcl::toROSMsg(pcl::PointCloud<pcl::PointXYZI>): height 1, width n, point_step = sizeof(PointXYZI) = 32, row_step =
    point_step * width, little-endian, fields x@0 y@4 z@8 intensity@16, all FLOAT32 count 1."""
    lib = _lib.load()
    d = CmPc2Desc()
    assert lib.cm_pointcloud2_describe(step, 12345, C.byref(d)) == _lib.CM_OK
    assert (d.height, d.width, d.point_step, d.row_step, d.is_bigendian, d.is_dense, d.n_fields) == (1, 12345, step, step * 12345, 0, 1, 4)
    got = [(d.fields[k].name.decode(), d.fields[k].offset, d.fields[k].datatype, d.fields[k].count) for k in range(4)]
    assert got == [("x", 0, F32, 1), ("y", 4, F32, 1), ("z", 8, F32, 1), ("intensity", off_i, F32, 1)]
    assert lib.cm_pointcloud2_describe(24, 1, C.byref(d)) == _lib.CM_E_INVALID
    # round trip: the described layout is accepted by the subscriber-side lookup
    rc, lay = layout_of(got, step)
    assert rc == _lib.CM_OK and (lay.point_step, lay.off_intensity) == (step, off_i)
