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
