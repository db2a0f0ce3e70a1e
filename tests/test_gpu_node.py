"""GPU parity test (-m gpu) of the device-resident mirror of one main-loop iteration of the reference's built node
(cloud_merger_b200/node.py; pc_preprocessing_main.cpp callbackX -> proceedX -> fusePointclouds -> voxelgrid) against the
oracle's composition of the same sequence, stage by stage on the CPU. Bar: /points_no_ground and /points_ground
bit-identical (set, order, coordinates); /points_voxel same voxels, centroids bit-equal to the float oracle and within
1e-5 of the float64 one."""
import numpy as np
import pytest

from cloud_merger_b200 import ROI_PASSES, synth
from cloud_merger_b200.node import FRONT_PARTS, NodeParams, PreprocessingNode, zones_of_parts

from helpers import assert_bit_equal, assert_centroids_close

pytestmark = pytest.mark.gpu


def oracle_frame(oracle, clouds, mats, p: NodeParams):
    no_ground, ground = [], []
    for s, raw in enumerate(clouds):
        cur = oracle.transform(raw, mats[s][:3].reshape(-1))
        for (axis, lo, hi, neg) in p.roi_passes:
            cur = np.ascontiguousarray(cur[oracle.passthrough(cur, axis, lo, hi, bool(neg))])
        parts = p.parts[s % len(p.parts)]
        k = len(parts)
        z = oracle.zone_split(cur, zones_of_parts(parts, p.roi_z_max))
        for i in range(k):
            low, high = z[i][0], z[k + i][0]
            r = oracle.plane_ransac(low, p.distance_threshold, p.prob, p.max_iterations, True, 12345, p.sum_order)
            rest = np.ascontiguousarray(np.delete(low, r["inliers"], axis=0))
            keep = oracle.radius_outlier(rest, p.radius, p.min_neighbor)
            no_ground += [rest[keep], high]
            ground.append(low[r["inliers"]])
    ng = np.concatenate(no_ground)
    vg = oracle.voxelgrid(ng, [p.voxel_size] * 3, p.points_per_voxel, True, force64=True)
    return ng, np.concatenate(ground), vg


@pytest.mark.parametrize("concurrent", [True, False])
def test_node_frame_matches_oracle_composition(gpu_ok, oracle, concurrent):
    """concurrent: the per-sensor stages on one host thread + CUDA stream per sensor (the reference's AsyncSpinner
    callbacks); otherwise one sensor after the other on the default stream."""
    S, rings, az = 3, 32, 512
    rear = tuple((length, -dev - length, zg) for (length, dev, zg) in FRONT_PARTS)   # a second window set (mirrored in x)
    p = NodeParams(parts=[FRONT_PARTS, rear])
    node = PreprocessingNode(S, rings * az, p, concurrent=concurrent)
    try:
        mats = [synth.extrinsic(s, S) for s in range(S)]
        for s in range(S):
            node.set_extrinsic(s, mats[s])
        for frame in range(3):   # several frames: the handles are reused from frame to frame
            clouds = [synth.lidar_cloud(5100 + frame, s, frame, rings, az) for s in range(S)]
            got = node.frame(clouds)
            ng, g, vg = oracle_frame(oracle, clouds, mats, p)
            assert got["n_no_ground"] == len(ng) and got["n_ground"] == len(g)
            assert len(g) > 1000 and len(ng) > 1000
            assert_bit_equal(got["no_ground"], ng, "/points_no_ground")
            assert_bit_equal(got["ground"], g, "/points_ground")
            assert got["n_voxels"] == vg["n"] > 100
            assert_bit_equal(got["voxel"], vg["centroid"], "/points_voxel vs float oracle")
            assert assert_centroids_close(got["voxel"], vg["centroid_f64"], "/points_voxel") <= 1e-5
    finally:
        node.close()


# the zone tables of the four sensor groups (pcl_preprocessing/src/Parameter.h:45-81, pc_preprocessing_main.cpp:228-312, :428-446,
# :474-497), (length, deviation, z_max_ground); deviation = the reference's float sums with roi_mid = 15; None = plain part
REAR_PARTS = ((30.0, 30.0, 2.0), (26.0, 4.0, 1.5), (8.0, -4.0, 0.3), (11.0, -15.0, 0.5))
TOP_PARTS = ((40.0, 20.0, 1.0), (35.0, -15.0, None))
LIVOX_PARTS = ((26.0, 34.0, 1.5), (10.0, 24.0, 1.2), (10.0, 14.0, 0.8), (10.0, 4.0, 0.5))


def oracle_proceed(oracle, roi, parts, p: NodeParams):
    no_ground, ground, planes = [], [], []
    f32 = np.float32
    for (length, dev, zg) in parts:
        x = (0, float(f32(dev)), float(f32(f32(dev) + f32(length))), 0)
        part = np.ascontiguousarray(roi[oracle.passthrough(roi, *x[:3], False)])
        if zg is None:
            no_ground.append(part)
            continue
        low = np.ascontiguousarray(part[oracle.passthrough(part, 2, float(-f32(zg)), float(f32(zg)))])
        high = part[oracle.passthrough(part, 2, float(f32(np.float64(f32(zg)) + 0.01)), float(f32(p.roi_z_max)))]
        r = oracle.plane_ransac(low, p.distance_threshold, p.prob, p.max_iterations, True, 12345, p.sum_order)
        rest = np.ascontiguousarray(np.delete(low, r["inliers"], axis=0))
        keep = oracle.radius_outlier(rest, p.radius, p.min_neighbor)
        no_ground += [rest[keep], high]
        ground.append(low[r["inliers"]])
        planes.append(r)
    cat = lambda xs: np.concatenate(xs) if xs else np.zeros((0, 4), np.float32)
    return cat(no_ground), cat(ground), planes


@pytest.mark.parametrize("name,parts", [("front", FRONT_PARTS), ("rear", REAR_PARTS), ("top", TOP_PARTS), ("livox", LIVOX_PARTS),
                                        ("empty_zone", ((5.0, 500.0, 1.0), (30.0, 30.0, 2.5), (20.0, -300.0, None)))])
def test_proceed_zones_one_call_matches_the_reference_sequence(gpu_ok, oracle, name, parts):
    """cm_proceed_zones: what a proceedX does after getROI, in one C call on one handle, for each of the reference's four
    zone tables (incl. the plain near-range part of the top sensor) and for windows that hold no point at all."""
    from cloud_merger_b200 import CloudMerger
    p = NodeParams()
    S = 2
    mats = [synth.extrinsic(s, S) for s in range(S)]
    with CloudMerger(max_sensors=1, max_points_per_sensor=1 << 17, max_batch_points=1 << 17, max_batch_frames=8) as cm:
        for frame in range(2):
            raw = synth.lidar_cloud(7300 + frame, frame % S, frame, 64, 1024)
            cur = oracle.transform(raw, mats[frame % S][:3].reshape(-1))
            for (axis, lo, hi, neg) in p.roi_passes:
                cur = np.ascontiguousarray(cur[oracle.passthrough(cur, axis, lo, hi, bool(neg))])
            got = cm.proceed_zones(cur, parts, p.roi_z_max, p.radius, p.min_neighbor, p.distance_threshold, p.prob, p.max_iterations)
            ng, g, planes = oracle_proceed(oracle, cur, parts, p)
            assert_bit_equal(got["no_ground"], ng, "%s no_ground" % name)
            assert_bit_equal(got["ground"], g, "%s ground" % name)
            assert len(got["planes"]) == len(planes)
            for a, b in zip(got["planes"], planes):
                assert a["found"] == b["found"] and a["iterations"] == b["iterations"] and a["n_inliers"] == len(b["inliers"])
                if b["found"]:
                    assert_bit_equal(a["coeff"], b["coeff"], "%s plane" % name)
            if name != "empty_zone":
                assert len(ng) > 500
