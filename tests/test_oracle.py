"""CPU tests of the oracle (not gpu): the C++ restatement (oracle/cm_oracle.cpp) against the hand-computed known-answer
vector, against the independent numpy restatement (oracle/np_oracle.py), and against the properties the PCL 1.8.1
algorithms guarantee. The reference ships no golden vectors of its own (parity unpinned)."""
import numpy as np
import pytest

from cloud_merger_b200 import synth
from oracle import np_oracle as npo

from helpers import assert_bit_equal, assert_centroids_close, cloud_dict, known_answer


def _ka_clouds(ka):
    mA = np.array(ka["extrinsics_row_major_3x4"]["A"], np.float32)
    mB = np.array(ka["extrinsics_row_major_3x4"]["B"], np.float32)
    return [cloud_dict(ka["A"], mA, ka["sensor_A_is_dense"]), cloud_dict(ka["B"], mB, ka["sensor_B_is_dense"])]


@pytest.mark.parametrize("min_points", [1, 2])
def test_known_answer_cpp(oracle, min_points):
    ka = known_answer()
    e = ka["expected"]
    r = oracle.merge_frame(_ka_clouds(ka), [tuple(p) for p in ka["crop_passes"]], ka["leaf"], min_points, True, False)
    assert r["survivor_src"].tolist() == e["survivor_src"]
    assert_bit_equal(r["survivor_xyzi"][:, :3], np.array(e["survivor_xyz"], np.float32), "survivor xyz")
    assert r["min_b"].tolist() == e["min_b"] and r["max_b"].tolist() == e["max_b"] and r["div_b"].tolist() == e["div_b"]
    assert r["point_idx"].tolist() == e["point_idx"]
    ex = e["min_points_%d" % min_points]
    assert r["idx"].tolist() == ex["idx"] and r["count"].tolist() == ex["count"]
    assert_bit_equal(r["centroid"], np.array(ex["centroid"], np.float32), "centroid")
    assert not r["pcl_overflow"]


@pytest.mark.parametrize("min_points", [1, 2])
def test_known_answer_numpy(min_points):
    ka = known_answer()
    e = ka["expected"]
    r = npo.merge_frame(_ka_clouds(ka), [tuple(p) for p in ka["crop_passes"]], ka["leaf"], min_points, True, False)
    assert r["survivor_src"].tolist() == e["survivor_src"]
    assert_bit_equal(r["survivor_xyzi"][:, :3], np.array(e["survivor_xyz"], np.float32), "survivor xyz")
    v = r["voxel"]
    assert v["min_b"].tolist() == e["min_b"] and v["div_b"].tolist() == e["div_b"]
    ex = e["min_points_%d" % min_points]
    assert v["idx"].tolist() == ex["idx"] and v["count"].tolist() == ex["count"]
    np.testing.assert_allclose(v["centroid_f64"], np.array(ex["centroid"], np.float64), rtol=1e-6)


def test_inclusive_bounds_and_nan(oracle):
    """PassThrough keeps v == lo and v == hi, drops NaN/Inf in x|y|z or the field, keeps order; negative flips."""
    p = np.array([[0, 0, -0.5, 1], [0, 0, 3.0, 2], [0, 0, np.nextafter(np.float32(3.0), np.float32(4)), 3],
                  [np.nan, 0, 1, 4], [0, np.inf, 1, 5], [0, 0, np.nan, 6], [0, 0, 1.0, 7],
                  [0, 0, np.nextafter(np.float32(-0.5), np.float32(-1)), 8]], np.float32)
    assert oracle.passthrough(p, 2, -0.5, 3.0).tolist() == [0, 1, 6]
    assert oracle.passthrough(p, 2, -0.5, 3.0, True).tolist() == [2, 7]
    assert np.nonzero(npo.passthrough_mask(p, 2, -0.5, 3.0))[0].tolist() == [0, 1, 6]
    assert np.nonzero(npo.passthrough_mask(p, 2, -0.5, 3.0, True))[0].tolist() == [2, 7]
    # intensity as the filter field
    assert oracle.passthrough(p, 3, 2, 6).tolist() == [1, 2]


def test_transform_is_unfused(oracle):
    """x' = ((m00*x + m01*y) + m02*z) + m03 with separately rounded products: a fused multiply-add gives another result
    for these operands, so this pins the no-FMA contract (reference built with -std=c++14 only, no -march)."""
    m = np.array([1.0000001, 3.0000002, 5.0000005, 0.1, 0.9999999, 1.0000002, 0.3333333, 7, 1e-3, 1e3, 0.1, -2], np.float32)
    rng = np.random.default_rng(7)
    p = rng.uniform(-100, 100, size=(4096, 4)).astype(np.float32)
    a = oracle.transform(p, m)
    b = npo.transform(p, m)
    assert_bit_equal(a, b, "cpp vs numpy transform")
    x, y, z = (p[:, k].astype(np.float64) for k in range(3))
    m64 = m.astype(np.float64)
    exact = np.float32(m64[0] * x + m64[1] * y + m64[2] * z + m64[3])  # what a fully fused evaluation approaches
    assert (a[:, 0] != exact).any(), "the unfused chain must differ from the exactly rounded sum somewhere"
    # non-dense clouds keep non-finite points untouched; dense clouds transform everything
    q = p.copy(); q[5, 0] = np.nan; q[9, 2] = np.inf
    nd = oracle.transform(q, m, is_dense=False)
    assert_bit_equal(nd[[5, 9]], q[[5, 9]], "non-dense keeps invalid points")
    assert_bit_equal(nd[:5], a[:5], "finite points unchanged by the flag")


@pytest.mark.parametrize("layout", [(16, 0, 4, 8, 12), (32, 0, 4, 8, 16), (22, 0, 4, 8, 12), (18, 0, 4, 8, 12),
                                    (48, 8, 12, 16, 36), (20, 4, 8, 12, -1), (19, 3, 7, 11, 15)])
def test_unpack_layouts(oracle, layout):
    step, ox, oy, oz, oi = layout
    rng = np.random.default_rng(3)
    p = rng.normal(size=(1000, 4)).astype(np.float32)
    data = synth.pack_cloud(p, step, ox, oy, oz, oi)
    a = oracle.unpack(data, len(p), step, ox, oy, oz, oi)
    b = npo.unpack(data, len(p), step, ox, oy, oz, oi)
    want = p.copy()
    if oi < 0:
        want[:, 3] = 0
    assert_bit_equal(a, want, "cpp unpack")
    assert_bit_equal(b, want, "numpy unpack")


@pytest.mark.parametrize("cfg,leaf,min_points", [("cfg1", 0.1, 2), ("cfg1", 0.1, 1), ("cfg2", 0.05, 2)])
def test_cpp_vs_numpy_on_config_shapes(oracle, cfg, leaf, min_points):
    """Both restatements, written independently, agree on the BASELINE shapes: survivors bit-exact, voxel membership,
    counts and order exact, centroids within the 1e-5 contract (float vs double accumulation)."""
    c = synth.CONFIGS[cfg]
    clouds, mats = synth.frame_clouds(cfg, 1000 * int(cfg[-1]), 0)
    cds = [cloud_dict(p, m[:3]) for p, m in zip(clouds, mats)]
    a = oracle.merge_frame(cds, c["passes"], [leaf] * 3, min_points, True, True)
    b = npo.merge_frame(cds, c["passes"], [leaf] * 3, min_points, True, True)
    assert a["n_survivors"] == len(b["survivor_src"]) > 1000
    assert (a["survivor_src"] == b["survivor_src"]).all()
    assert_bit_equal(a["survivor_xyzi"], b["survivor_xyzi"], "survivors")
    v = b["voxel"]
    assert (a["idx"] == v["idx"]).all() and (a["count"] == v["count"]).all()
    assert (a["point_idx"] == v["point_idx"]).all()
    assert a["min_b"].tolist() == v["min_b"].tolist() and a["div_b"].tolist() == v["div_b"].tolist()
    worst = assert_centroids_close(a["centroid"], v["centroid_f64"], "float(asc. index) vs numpy f64")
    np.testing.assert_allclose(a["centroid_f64"], v["centroid_f64"], rtol=1e-12, atol=1e-12)
    assert worst < 1e-5
    assert a["count"].min() >= min_points


def test_pcl_int32_and_64bit_formulations_agree(oracle):
    """Inside PCL's domain the int32 formulation (force64=0) and the 64-bit extension give identical voxels."""
    clouds, mats = synth.frame_clouds("cfg1", 1000, 1)
    x = np.concatenate([oracle.transform(p, m[:3].reshape(-1)) for p, m in zip(clouds, mats)])
    x = x[npo.crop(x, synth.ROI_BOX)]
    a = oracle.voxelgrid(x, [0.1] * 3, 2, True, force64=False)
    b = oracle.voxelgrid(x, [0.1] * 3, 2, True, force64=True)
    assert not a["pcl_overflow"]
    for k in ("idx", "count", "point_idx"):
        assert (a[k] == b[k]).all(), k
    assert_bit_equal(a["centroid"], b["centroid"], "centroid")
    # std::sort order vs ascending index order: both inside the tolerance band of the double sum
    assert_centroids_close(a["centroid_sort"], a["centroid_f64"], "std::sort order")
    assert_centroids_close(a["centroid"], a["centroid_f64"], "ascending order")


def test_pcl_overflow_guard(oracle):
    """dx*dy*dz > INT32_MAX: PCL 1.8.1 warns and returns the input unchanged; the 64-bit extension carries on."""
    x = synth.uniform_cloud(5, 20000, extent=(200.0, 200.0, 10.0))
    a = oracle.voxelgrid(x, [0.02] * 3, 1, True, force64=False)
    assert a["pcl_overflow"] and a["returned_input"] and a["n"] == len(x)
    assert_bit_equal(a["centroid"], x, "output = input")
    b = oracle.voxelgrid(x, [0.02] * 3, 1, True, force64=True)
    assert b["pcl_overflow"] and not b["returned_input"]
    n = npo.voxelgrid(x, [0.02] * 3, 1, True, True)
    assert (b["idx"] == n["idx"]).all() and (b["count"] == n["count"]).all()
    assert b["idx"].max() > 2**31
    c = oracle.voxelgrid(x, [0.5] * 3, 1, True, force64=False)
    assert not c["pcl_overflow"]


def test_voxel_membership_is_permutation_invariant(oracle):
    rng = np.random.default_rng(11)
    x = synth.uniform_cloud(9, 50000, extent=(20.0, 20.0, 4.0))
    perm = rng.permutation(len(x))
    a = oracle.voxelgrid(x, [0.25] * 3, 2)
    b = oracle.voxelgrid(x[perm], [0.25] * 3, 2)
    assert (a["idx"] == b["idx"]).all() and (a["count"] == b["count"]).all()
    assert (a["point_idx"][perm] == b["point_idx"]).all()
    assert_centroids_close(b["centroid"], a["centroid_f64"], "permuted")


def test_crop_commutes_with_concat(oracle):
    """transform -> crop -> concat (reference order) == transform -> concat -> crop (north_star order)."""
    clouds, mats = synth.frame_clouds("cfg1", 1000, 2, nan_frac=0.005)
    tr = [oracle.transform(p, m[:3].reshape(-1), is_dense=False) for p, m in zip(clouds, mats)]
    per = np.concatenate([t[npo.crop(t, synth.ROI_BOX)] for t in tr])
    allc = np.concatenate(tr)
    assert_bit_equal(per, allc[npo.crop(allc, synth.ROI_BOX)], "crop o concat")
    cds = [cloud_dict(p, m[:3], is_dense=0) for p, m in zip(clouds, mats)]
    r = oracle.merge_frame(cds, synth.ROI_BOX, [0.1] * 3, 2)
    assert_bit_equal(r["survivor_xyzi"], per, "merge_frame survivors")
    assert np.isfinite(r["survivor_xyzi"]).all()


def test_min_points_filter_and_empty_inputs(oracle):
    x = np.array([[0.01, 0.01, 0.01, 1], [0.02, 0.02, 0.02, 3], [5, 5, 5, 7]], np.float32)
    a = oracle.voxelgrid(x, [0.1] * 3, 2)
    assert a["n"] == 1 and a["count"].tolist() == [2]
    assert_bit_equal(a["centroid"], np.array([[np.float32(0.03) / np.float32(2), np.float32(0.03) / np.float32(2),
                                               np.float32(0.03) / np.float32(2), 2.0]], np.float32), "mean")
    assert oracle.voxelgrid(x, [0.1] * 3, 0)["n"] == 2 and oracle.voxelgrid(x, [0.1] * 3, 4)["n"] == 0
    assert oracle.voxelgrid(x, [0.1] * 3, 1, downsample_all=False)["centroid"][:, 3].tolist() == [0.0, 0.0]
    e = oracle.voxelgrid(np.zeros((0, 4), np.float32), [0.1] * 3, 1)
    assert e["n"] == 0
    allnan = np.full((4, 4), np.nan, np.float32)
    assert oracle.voxelgrid(allnan, [0.1] * 3, 1, is_dense=False)["n"] == 0
    r = oracle.merge_frame([cloud_dict(np.zeros((0, 4), np.float32), np.eye(4)[:3])], synth.ROI_BOX, [0.1] * 3, 1)
    assert r["n_survivors"] == 0 and r["n_voxels"] == 0


def test_tf_to_matrix(oracle):
    """Eigen 3.3 quaternion -> rotation as pcl_ros builds the Affine3f from the tf::Transform (host glue)."""
    q = np.array([0.0, 0.0, np.sin(np.pi / 4), np.cos(np.pi / 4)])  # yaw +90 deg
    m = oracle.tf_to_matrix(q, [0.0, 2.0, 0.0]).reshape(3, 4)
    np.testing.assert_allclose(m, [[0, -1, 0, 0], [1, 0, 0, 2], [0, 0, 1, 0]], atol=1e-7)
    rng = np.random.default_rng(5)
    for _ in range(20):
        qq = rng.normal(size=4); qq /= np.linalg.norm(qq)
        t = rng.normal(size=3)
        assert_bit_equal(oracle.tf_to_matrix(qq, t), npo.tf_to_matrix(qq, t), "tf matrix")


def test_threads_do_not_change_results(oracle):
    clouds, mats = synth.frame_clouds("cfg1", 1000, 3)
    cds = [cloud_dict(p, m[:3]) for p, m in zip(clouds, mats)]
    a = oracle.merge_frame(cds, synth.ROI_BOX, [0.1] * 3, 2, threads=1)
    b = oracle.merge_frame(cds, synth.ROI_BOX, [0.1] * 3, 2, threads=6)
    assert (a["idx"] == b["idx"]).all() and (a["survivor_src"] == b["survivor_src"]).all()
    assert_bit_equal(a["centroid"], b["centroid"], "threads")


def test_zone_split_oracle_against_numpy_masks(oracle):
    """The oracle's literal replay of getCloudPart + z windows (PassThrough after PassThrough, each on the copied cloud)
    against an independent restatement: one boolean mask per zone on the original cloud, float32 compares."""
    from helpers import reference_front_zones
    rng = np.random.default_rng(3)
    n = 20000
    cloud = np.column_stack([rng.uniform(-16, 61, n), rng.uniform(-6, 6, n), rng.uniform(-1, 3.5, n),
                             rng.uniform(0, 255, n)]).astype(np.float32)
    cloud[::97, 0] = np.float32(19.0)   # shared window end
    cloud[::101, 2] = np.nan
    cloud[::103, 1] = np.inf
    cloud[::107, 3] = np.nan            # only an intensity stage may reject these
    zones = reference_front_zones() + [[(3, 10.0, 20.0, 0)], [(1, -1.0, 1.0, 1)], []]
    got = oracle.zone_split(cloud, zones)
    fin = np.isfinite(cloud[:, :3]).all(axis=1)
    for z, chain in enumerate(zones):
        keep = fin.copy() if chain else np.ones(n, bool)  # no stage = no filter
        for (axis, lo, hi, neg) in chain:
            v = cloud[:, axis]
            lo, hi = np.float32(lo), np.float32(hi)
            with np.errstate(invalid="ignore"):
                inside = (v >= lo) & (v <= hi)
            keep &= np.isfinite(v) & (~inside if neg else inside)
        want = np.nonzero(keep)[0]
        assert (got[z][1] == want).all(), "zone %d" % z
        assert (got[z][0].view(np.uint32) == cloud[want].view(np.uint32)).all()
    # the reference's windows: x = 19 sits in both neighbours, the 0.01 gap above z_max_ground belongs to nobody
    on_edge = np.nonzero(fin & (cloud[:, 0] == np.float32(19.0)))[0]
    in_mid2 = np.union1d(got[2][1], got[3][1])
    in_mid = np.union1d(got[4][1], got[5][1])
    zwin = lambda i, zg: (cloud[i, 2] >= -zg) & (cloud[i, 2] <= zg) | (cloud[i, 2] >= np.float32(np.float64(np.float32(zg)) + 0.01)) & (cloud[i, 2] <= 3.0)
    assert set(on_edge[zwin(on_edge, np.float32(2.0))]) <= set(in_mid2) and set(on_edge[zwin(on_edge, np.float32(1.5))]) <= set(in_mid)


def test_radius_outlier_oracle_against_brute_force(oracle):
    """The grid-accelerated C++ restatement of pcl::RadiusOutlierRemoval (PCL 1.8.1 + FLANN L2_Simple) against an O(n^2)
    numpy count in the same float32 operation order: strictly inside the radius, the point itself counted, keep iff
    count > min_pts; negative inverts; non-finite points are neither kept nor counted."""
    rng = np.random.default_rng(9)
    n = 2500
    x = np.column_stack([rng.uniform(0, 5, n), rng.uniform(0, 3, n), rng.uniform(0, 1, n), rng.uniform(0, 1, n)]).astype(np.float32)
    x[::211, 1] = np.nan
    r = float(np.float32(0.15))
    r2 = np.float32(r * r)
    fin = np.isfinite(x[:, :3]).all(axis=1)
    d = x[:, None, :3] - x[None, :, :3]
    with np.errstate(invalid="ignore"):
        acc = (d[..., 0] * d[..., 0]).astype(np.float32)
        acc = (acc + (d[..., 1] * d[..., 1]).astype(np.float32)).astype(np.float32)
        acc = (acc + (d[..., 2] * d[..., 2]).astype(np.float32)).astype(np.float32)
        inside = (acc < r2) & fin[None, :] & fin[:, None]
    k = inside.sum(axis=1)
    for min_pts in (0, 1, 2, 4):
        for neg in (False, True):
            want = np.nonzero(fin & ((k <= min_pts) if neg else (k > min_pts)))[0]
            got = oracle.radius_outlier(x, r, min_pts, neg)
            assert len(got) == len(want) and (got == want).all(), (min_pts, neg)


# ---- RANSAC ground plane (pcl::SACSegmentation; pc_preprocessing_main.cpp:95-108) -----------------------------------------
def _ground_scene(seed, n, ground_frac=0.7, slope=0.02, noise=0.03):
    rng = np.random.default_rng(seed)
    g = int(n * ground_frac)
    pts = np.zeros((n, 4), np.float32)
    pts[:, 0] = rng.uniform(-30, 30, n)
    pts[:, 1] = rng.uniform(-10, 10, n)
    pts[:g, 2] = -1.8 + slope * pts[:g, 0] + rng.normal(0, noise, g)
    pts[g:, 2] = rng.uniform(-1.5, 1.0, n - g)
    pts[:, 3] = rng.uniform(0, 255, n)
    return pts[rng.permutation(n)]


def test_plane_sampler_engine_known_answers(oracle):
    """The sampler's engine is mt19937: the C++ standard's known answer (10000th output of the default seed 5489 is
    4123659995) and numpy's init_genrand seeding for the seed PCL uses."""
    assert oracle.mt19937_at(5489, 9999) == 4123659995
    raw = np.random.RandomState(12345)._bit_generator.random_raw(1000)
    for i in (0, 1, 623, 624, 999):
        assert oracle.mt19937_at(12345, i) == int(raw[i])


@pytest.mark.parametrize("order", [0, 1, 2])
def test_plane_score_against_numpy(oracle, order):
    from oracle import np_oracle
    x = _ground_scene(3, 4000)
    rng = np.random.default_rng(8)
    thr = float(np.float32(0.3))
    for _ in range(25):
        smp = rng.choice(len(x), 3, replace=False).astype(np.int32)
        good, c = np_oracle.plane_of_sample(x[smp[0]], x[smp[1]], x[smp[2]], order)
        got_c, got_k = oracle.plane_score(x, smp, thr, order)
        assert good and got_k == int(np_oracle.plane_inlier_mask(x, c, thr, order).sum())
        assert got_c.tobytes() == c.tobytes()
    # collinear sample (equal quotients on all axes) and coincident points are rejected
    line = np.array([[0, 0, 0, 0], [1, 2, 4, 0], [2, 4, 8, 0], [5, 1, 1, 0]], np.float32)
    assert oracle.plane_score(line, np.array([0, 1, 2], np.int32), thr, order)[1] == -1
    assert oracle.plane_score(line, np.array([0, 1, 3], np.int32), thr, order)[1] >= 3


@pytest.mark.parametrize("order", [0, 2])
def test_plane_ransac_against_numpy_and_recovers_the_ground(oracle, order):
    from oracle import np_oracle
    for seed, n, frac in ((1, 3000, 0.7), (2, 1500, 0.35), (3, 64, 0.8)):
        x = _ground_scene(seed, n, frac)
        thr = float(np.float32(0.3))
        want = np_oracle.plane_ransac(x, thr, float(np.float32(0.99)), 1000, 12345, order)
        got = oracle.plane_ransac(x, thr, float(np.float32(0.99)), 1000, optimize=False, seed=12345, sum_order=order)
        assert got["found"] and want["found"]
        assert (got["iterations"], got["draws"], got["best_count"]) == (want["iterations"], want["draws"], want["best_count"])
        assert (got["sample"] == want["sample"]).all() and got["coeff_ransac"].tobytes() == want["coeff"].tobytes()
        mask = np_oracle.plane_inlier_mask(x, want["coeff"], thr, order)
        assert (got["inliers"] == np.nonzero(mask)[0]).all()
        # the least-squares refit tightens the plane: it must be close to the generating plane z = -1.8 + 0.02 x
        ref = oracle.plane_ransac(x, thr, float(np.float32(0.99)), 1000, optimize=True, seed=12345, sum_order=order)
        c = ref["coeff"] * np.sign(ref["coeff"][2])
        truth = np.array([-0.02, 0.0, 1.0, 1.8]) / np.sqrt(1 + 0.02 ** 2)
        if frac >= 0.7 and n >= 1000:
            assert np.abs(c - truth).max() < 2e-2
        assert abs(float(np.linalg.norm(c[:3])) - 1.0) < 1e-4


def test_plane_ransac_degenerate_inputs(oracle):
    thr = 0.3
    for n in (0, 1, 2):
        r = oracle.plane_ransac(np.zeros((n, 4), np.float32), thr)
        assert not r["found"] and len(r["inliers"]) == 0
    # all points on one line: every draw is rejected, getSamples gives up after 1000 of them
    t = np.arange(50, dtype=np.float32)
    line = np.column_stack([t, 2 * t, 4 * t, t]).astype(np.float32)
    r = oracle.plane_ransac(line, thr)
    assert not r["found"] and r["iterations"] == 0 and r["draws"] == 1000
    # the iteration cap: a cloud without a dominant plane runs until iterations > max_iterations
    rng = np.random.default_rng(5)
    blob = rng.uniform(-20, 20, (4000, 4)).astype(np.float32)
    r = oracle.plane_ransac(blob, 0.05, 0.99, 40)
    assert r["found"] and r["iterations"] == 41


def test_node_zone_chains_are_the_reference_windows():
    """Host logic of cloud_merger_b200/node.py: the PassThrough chains it configures for FRONT_PARTS are the ten windows of
    proceedFront (tests/helpers.reference_front_zones), ground windows first, then the upper ones."""
    from cloud_merger_b200.node import FRONT_PARTS, zones_of_parts
    from helpers import reference_front_zones
    ref = reference_front_zones()            # interleaved: ground_k, upper_k
    got = zones_of_parts(FRONT_PARTS, 3.0)   # grouped: all ground windows, then all upper windows
    k = len(FRONT_PARTS)
    assert len(got) == 2 * k == len(ref)
    for i in range(k):
        assert [tuple(p) for p in got[i]] == [tuple(p) for p in ref[2 * i]]
        assert [tuple(p) for p in got[k + i]] == [tuple(p) for p in ref[2 * i + 1]]


def test_plane_refit_against_float64_eigensolver(oracle):
    """The restated float32 refit (running sums in PCL's order, covariance = E[ab] - E[a]E[b], pcl::eigen33's closed-form
    smallest eigenvector) must agree with a float64 eigen-decomposition of the same inliers' covariance to float accuracy:
    a transcription error in the cubic's roots or in the eigenvector selection would show here."""
    thr = float(np.float32(0.3))
    for seed, n, frac, slope in ((1, 20000, 0.8, 0.02), (2, 5000, 0.6, -0.05), (3, 800, 0.9, 0.0), (4, 60000, 0.5, 0.1)):
        rng = np.random.default_rng(seed)
        x = _ground_scene(seed, n, frac, slope=slope)
        x[:, :3] += rng.uniform(-20, 20, 3).astype(np.float32)      # away from the origin: E[ab] - E[a]E[b] cancels
        r0 = oracle.plane_ransac(x, thr, 0.99, 1000, optimize=False)
        r1 = oracle.plane_ransac(x, thr, 0.99, 1000, optimize=True)
        assert r0["found"] and r1["found"]
        inl = x[r0["inliers"], :3].astype(np.float64)             # the refit runs on the RANSAC model's inliers
        c = inl.mean(axis=0)
        w, v = np.linalg.eigh(np.cov((inl - c).T, bias=True))
        nrm = v[:, 0] * np.sign(v[2, 0])
        got = r1["coeff"].astype(np.float64) * np.sign(r1["coeff"][2])
        # measured: <= 8e-6 on the normal, <= 2e-4 on d (float32 sums of squares 20 m from the origin)
        assert np.abs(got[:3] - nrm).max() < 1e-4, (seed, got[:3], nrm)
        assert abs(got[3] + float(nrm @ c)) < 2e-3, (seed, got[3], -float(nrm @ c))
        assert abs(np.linalg.norm(got[:3]) - 1.0) < 1e-5


def test_plane_ransac_golden_fixture(oracle):
    """tests/golden/plane_ransac.json (written by the numpy restatement, tests/golden/make_plane_fixture.py) against the
    C++ oracle: sampler engine outputs, sample, iteration / draw counts, coefficients bit for bit, inlier set."""
    import json
    import os
    doc = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "plane_ransac.json")))
    for i, v in enumerate(doc["mt19937_seed_12345_first_outputs"]):
        assert oracle.mt19937_at(12345, i) == v
    for c in doc["cases"]:
        x = np.array(c["xyzi"], np.float32)
        e = c["expected"]
        r = oracle.plane_ransac(x, c["threshold"], c["probability"], c["max_iterations"], optimize=False, seed=c["seed"],
                                sum_order=c["sum_order"])
        assert (r["found"], r["iterations"], r["draws"], r["best_count"]) == (e["found"], e["iterations"], e["draws"], e["best_count"])
        assert r["sample"].tolist() == e["sample"]
        if e["found"]:
            assert r["coeff_ransac"].view(np.uint32).tolist() == e["coeff_ransac_bits"]
        assert r["inliers"].tolist() == e["inliers"]


def _aux_fixture():
    import json
    import os
    doc = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "zones_outlier.json")))
    x = np.array([[np.nan if v == "nan" else v for v in row] for row in doc["xyzi"]], np.float32)
    return doc, x


def test_zones_and_outlier_golden_fixture(oracle):
    """tests/golden/zones_outlier.json (plain numpy, tests/golden/make_aux_fixtures.py) against the C++ oracle."""
    doc, x = _aux_fixture()
    got = oracle.zone_split(x, [[tuple(p) for p in z] for z in doc["zones"]])
    for (gx, gi), want in zip(got, doc["zone_indices"]):
        assert gi.tolist() == want
    for m, want in doc["outlier_kept_by_min_pts"].items():
        assert oracle.radius_outlier(x, doc["radius"], int(m), False).tolist() == want


def test_pcl_golden_fixtures(oracle):
    """tests/golden/pcl_golden.npz is written by oracle/pcl_ref/dump_golden.py FROM THE REAL PCL on a machine that has it
    (oracle/pcl_ref/README.md). Once committed, it pins the oracle everywhere. Until then parity stays unpinned and this
    test skips, saying so."""
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pcl_golden.npz")
    if not os.path.exists(path):
        pytest.skip("no PCL-written golden file yet (parity unpinned): run oracle/pcl_ref/dump_golden.py where PCL is installed")
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(path), "..", "..", "oracle", "pcl_ref"))
    from oracle.pcl_ref import dump_golden
    g = np.load(path)
    clouds, mats = dump_golden.inputs()

    def same(a, b):
        a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32)
        return a.shape == b.shape and bool(((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))).all())
    merged = None
    for s, (c, m) in enumerate(zip(clouds, mats)):
        t = oracle.transform(c, m[:3], False)
        assert same(t, g["transform_%d" % s])
        cur = t
        for k, (axis, lo, hi, neg) in enumerate(synth.ROI_BOX):
            keep = oracle.passthrough(cur, axis, lo, hi, bool(neg))
            assert np.array_equal(keep, g["roi_%d_pass%d_idx" % (s, k)])
            cur = np.ascontiguousarray(cur[keep])
        merged = cur if merged is None else np.concatenate([merged, cur])
    assert same(merged, g["merged"])
    for leaf in dump_golden.LEAVES:
        for mp in (1, 2):
            o = oracle.voxelgrid(merged, [leaf] * 3, mp, True, force64=False)
            v, grid = g["voxel_%g_%d" % (leaf, mp)], g["grid_%g_%d" % (leaf, mp)]
            assert o["n"] == len(v) and o["min_b"].tolist() == grid[0:3].tolist() and o["div_b"].tolist() == grid[6:9].tolist()
            err = np.abs(v.astype(np.float64) - o["centroid_f64"]) / np.maximum(np.abs(o["centroid_f64"]), 1e-2)
            assert err.max() <= 1e-5
    assert np.array_equal(oracle.radius_outlier(merged, 0.15, 1), g["outlier_0.15_1_idx"])
    low = np.ascontiguousarray(merged[oracle.passthrough(merged, 2, -0.5, 0.5)])
    assert len(low) == int(g["plane_input_n"])
    assert any(np.array_equal(oracle.plane_ransac(low, 0.3, 0.99, 1000, True, 12345, order)["inliers"], g["plane_inliers"])
               for order in (0, 1, 2))
    assert same(oracle.tf_to_matrix(g["tf_q"], g["tf_t"]), g["tf_matrix"])
