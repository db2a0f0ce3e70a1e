"""GPU parity tests (-m gpu) of the radius outlier removal (outlierRemoval() of the reference,
pc_preprocessing_main.cpp:184-192: pcl::RadiusOutlierRemoval, radius 0.15 m, min_neighbor 1) through the C ABI against
the oracle's restatement of PCL 1.8.1 + FLANN (count of points strictly inside the radius, the point itself included,
float squared distances; keep iff count > min_pts). Bar: bit-exact survivor set, order and coordinates."""
import numpy as np
import pytest

from cloud_merger_b200 import ROI_PASSES, CloudMerger, CloudMergerError, _lib, synth

from helpers import assert_bit_equal

pytestmark = pytest.mark.gpu

RADIUS = float(np.float32(0.15))  # `const float radius = 0.15;` widened to the double setRadiusSearch takes (Parameter.h:23)


def _roi_cloud(oracle, seed, rings, az, sensor=0):
    cur = oracle.transform(synth.lidar_cloud(seed, sensor, 0, rings, az), synth.extrinsic(sensor, 4)[:3].reshape(-1))
    for (axis, lo, hi, neg) in ROI_PASSES:
        cur = np.ascontiguousarray(cur[oracle.passthrough(cur, axis, lo, hi, bool(neg))])
    return cur


@pytest.mark.parametrize("rings,az,min_pts,negative", [(16, 256, 1, False), (64, 1024, 1, False), (64, 1024, 3, False),
                                                       (32, 512, 2, True)])
def test_radius_outlier_matches_oracle(gpu_ok, oracle, rings, az, min_pts, negative):
    cloud = _roi_cloud(oracle, 4300, rings, az)
    want = oracle.radius_outlier(cloud, RADIUS, min_pts, negative)
    assert 0 < len(want) < len(cloud)
    with CloudMerger(max_sensors=1, max_points_per_sensor=len(cloud), max_batch_points=len(cloud)) as cm:
        gx, gi = cm.radius_outlier(cloud, RADIUS, min_pts, negative)
        assert len(gi) == len(want) and (gi == want).all(), "survivor set / order differs"
        assert_bit_equal(gx, cloud[want], "survivor coordinates")
        buf = cm.upload(cloud)
        cm.dev_radius_outlier(buf.ptr, len(cloud), RADIUS, min_pts, negative)
        dx, di = cm.radius_outlier_out()
        assert (di == want).all()
        assert_bit_equal(dx, cloud[want], "device form coordinates")


def test_radius_outlier_boundaries_and_invalid(gpu_ok, oracle):
    """Distances right at the radius (strictly-inside rule, float arithmetic), pairs that straddle cell borders, duplicate
    points, non-finite points (never kept, never counted), clouds of 0 / 1 / 2 points."""
    f32 = np.float32
    r = RADIUS
    rows = []
    rng = np.random.default_rng(21)
    for k in range(400):  # pairs at distance r * (1 +- a few ulp) along random directions, anywhere in the ROI
        c = np.array([rng.uniform(-15, 60), rng.uniform(-5, 5), rng.uniform(-0.5, 3.0)])
        d = rng.normal(size=3); d /= np.linalg.norm(d)
        s = r * (1.0 + (k % 9 - 4) * 2.0 ** -23)
        rows.append(c); rows.append(c + d * s)
    pts = np.array(rows, np.float64)
    cloud = np.column_stack([pts, np.arange(len(pts))]).astype(np.float32)
    extra = np.array([[1.0, 1.0, 1.0, 0.0], [1.0, 1.0, 1.0, 1.0],                 # duplicates: each other's neighbour
                      [np.nan, 0.0, 0.0, 2.0], [0.0, np.inf, 0.0, 3.0],           # non-finite
                      [40.0, 4.0, 2.0, 4.0]], np.float32)                         # isolated
    cloud = np.concatenate([cloud, extra])
    for min_pts, neg in ((1, False), (1, True), (0, False)):
        want = oracle.radius_outlier(cloud, r, min_pts, neg)
        with CloudMerger(max_sensors=1, max_points_per_sensor=len(cloud), max_batch_points=len(cloud)) as cm:
            gx, gi = cm.radius_outlier(cloud, r, min_pts, neg)
        assert len(gi) == len(want) and (gi == want).all(), "min_pts %d negative %s" % (min_pts, neg)
        assert_bit_equal(gx, cloud[want], "coordinates")
    # both sides of the strict rule must be present in the data, otherwise the test proves nothing
    k1 = set(oracle.radius_outlier(cloud[:800], r, 1, False).tolist())
    assert 0 < len(k1) < 800
    with CloudMerger(max_sensors=1, max_points_per_sensor=16, max_batch_points=16) as cm:
        for n in (0, 1, 2):
            c = np.array([[0, 0, 0, 1], [0.1, 0, 0, 2]], np.float32)[:n]
            gx, gi = cm.radius_outlier(c, r, 1)
            assert gi.tolist() == oracle.radius_outlier(c, r, 1).tolist()
        with pytest.raises(CloudMergerError):
            cm.radius_outlier(np.zeros((4, 4), np.float32), 0.0, 1)


def test_radius_outlier_properties_full_size(gpu_ok):
    """Size-independent checks on 1 Mi points: the result is a subsequence of the input (order, bit-exact coordinates);
    the filter is monotone in min_pts; with a negative filter the two outputs partition the finite points."""
    n = 1 << 20
    rng = np.random.default_rng(33)
    cloud = np.column_stack([rng.uniform(-15, 60, n), rng.uniform(-5, 5, n), rng.uniform(-0.5, 3, n),
                             rng.uniform(0, 255, n)]).astype(np.float32)
    with CloudMerger(max_sensors=1, max_points_per_sensor=n, max_batch_points=n) as cm:
        x1, i1 = cm.radius_outlier(cloud, RADIUS, 1)
        x3, i3 = cm.radius_outlier(cloud, RADIUS, 3)
        xn, in_ = cm.radius_outlier(cloud, RADIUS, 1, negative=True)
    assert (np.diff(i1.astype(np.int64)) > 0).all() and (np.diff(in_.astype(np.int64)) > 0).all()
    assert_bit_equal(x1, cloud[i1], "kept coordinates")
    assert set(i3.tolist()) <= set(i1.tolist()) and 0 < len(i3) < len(i1) < n
    assert len(i1) + len(in_) == n and len(np.intersect1d(i1, in_)) == 0


def test_radius_outlier_multi_equals_separate_runs(gpu_ok, oracle):
    """Several clouds in one call (the per-zone outlierRemoval calls of one proceedX): every cloud must come out exactly as
    its own run -- neighbours are never counted across clouds, even where two clouds overlap in space; empty and tiny
    clouds and clouds with non-finite points in between."""
    a = _roi_cloud(oracle, 4400, 32, 512)
    b = _roi_cloud(oracle, 4401, 16, 256, sensor=1)
    shifted = a.copy()
    shifted[:, 0] += np.float32(0.05)          # lies inside `a`'s space: cross-cloud neighbours would change the result
    bad = b.copy()
    bad[::53, 1] = np.nan
    clouds = [a, shifted, np.zeros((0, 4), np.float32), b[:1], bad, b[:2], a[::3]]
    total = sum(len(c) for c in clouds)
    with CloudMerger(max_sensors=1, max_points_per_sensor=total, max_batch_points=total, max_batch_frames=8) as cm:
        for min_pts, neg in ((1, False), (2, True)):
            got = cm.radius_outlier_multi(clouds, RADIUS, min_pts, neg)
            assert len(got) == len(clouds)
            for c, (gx, gi) in zip(clouds, got):
                want = oracle.radius_outlier(c, RADIUS, min_pts, neg)
                assert len(gi) == len(want) and (gi == want).all()
                assert_bit_equal(gx, c[want], "survivor coordinates")
        # device form
        begin = np.concatenate([[0], np.cumsum([len(c) for c in clouds])]).astype(np.int64)
        buf = cm.upload(np.ascontiguousarray(np.concatenate(clouds)))
        cm.dev_radius_outlier_multi(buf.ptr, begin, RADIUS, 1)
        zones = cm.zone_out()
        for k, c in enumerate(clouds):
            want = oracle.radius_outlier(c, RADIUS, 1, False)
            assert (zones[k][1].astype(np.int64) - begin[k] == want).all()
    with CloudMerger(max_sensors=1, max_points_per_sensor=total, max_batch_points=total) as cm1:   # one frame only
        with pytest.raises(CloudMergerError):
            cm1.radius_outlier_multi(clouds[:2], RADIUS, 1)


def test_radius_outlier_table_and_search_paths_agree(gpu_ok, oracle, monkeypatch):
    """The direct-address cell table (small key spaces) and the binary searches (large ones, forced here through the
    CM_ROR_NO_TABLE test hook) must give the same survivors -- on one cloud and on several clouds per call."""
    a = _roi_cloud(oracle, 4500, 64, 1024)
    b = _roi_cloud(oracle, 4501, 32, 512, sensor=1)
    want_a = oracle.radius_outlier(a, RADIUS, 1, False)
    want_b = oracle.radius_outlier(b, RADIUS, 2, False)
    cap = len(a) + 2 * len(b)
    with CloudMerger(max_sensors=1, max_points_per_sensor=cap, max_batch_points=cap, max_batch_frames=4) as cm:
        for no_table in (False, True):
            if no_table:
                monkeypatch.setenv("CM_ROR_NO_TABLE", "1")
            else:
                monkeypatch.delenv("CM_ROR_NO_TABLE", raising=False)
            gx, gi = cm.radius_outlier(a, RADIUS, 1)
            assert (gi == want_a).all()
            got = cm.radius_outlier_multi([b, a, b], RADIUS, 2)
            assert (got[0][1] == want_b).all() and (got[2][1] == want_b).all()
            assert (got[1][1] == oracle.radius_outlier(a, RADIUS, 2, False)).all()
    # a cloud spread over kilometres: the key space is far too large for the table, the searches run by themselves
    rng = np.random.default_rng(3)
    wide = np.column_stack([rng.uniform(-2000, 2000, 20000), rng.uniform(-2000, 2000, 20000), rng.uniform(-5, 5, 20000),
                            np.zeros(20000)]).astype(np.float32)
    wide[:5000, :3] = (wide[5000:10000, :3] + rng.normal(0, 0.05, (5000, 3))).astype(np.float32)   # some close pairs
    monkeypatch.delenv("CM_ROR_NO_TABLE", raising=False)
    with CloudMerger(max_sensors=1, max_points_per_sensor=len(wide), max_batch_points=len(wide)) as cm:
        gx, gi = cm.radius_outlier(wide, RADIUS, 1)
        want = oracle.radius_outlier(wide, RADIUS, 1, False)
        assert 0 < len(want) < len(wide) and (gi == want).all()
