"""GPU parity tests (-m gpu) of zone slicing -- several PassThrough chains over one cloud in one pass -- through the C ABI
against the oracle's literal replay of the reference sequence (getCloudPart x5, each followed by the two z windows of
removeGround; pc_preprocessing_main.cpp:49-59, 80-92, 228-270). Bar: bit-exact membership, order and coordinates."""
import numpy as np
import pytest

from cloud_merger_b200 import ROI_PASSES, CloudMerger, CloudMergerError, _lib, make_layout, synth

from helpers import assert_bit_equal, reference_front_zones

pytestmark = pytest.mark.gpu


def _roi_cloud(oracle, seed, rings, az, sensor=0):
    """One synthetic sensor cloud, transformed and cropped to the ROI on the CPU: what getROI hands to getCloudPart."""
    raw = synth.lidar_cloud(seed, sensor, 0, rings, az)
    m = synth.extrinsic(sensor, 4)
    cur = oracle.transform(raw, m[:3].reshape(-1))
    for (axis, lo, hi, neg) in ROI_PASSES:
        cur = np.ascontiguousarray(cur[oracle.passthrough(cur, axis, lo, hi, bool(neg))])
    return cur


def _check(got, want, what):
    assert len(got) == len(want)
    for z, ((gx, gs), (ox, os_)) in enumerate(zip(got, want)):
        assert len(gs) == len(os_), "%s zone %d: %d points vs oracle %d" % (what, z, len(gs), len(os_))
        assert (gs == os_).all(), "%s zone %d: membership / order differs" % (what, z)
        assert_bit_equal(gx, ox, "%s zone %d coordinates" % (what, z))


@pytest.mark.parametrize("rings,az", [(16, 64), (128, 1024)])
def test_reference_front_zones(gpu_ok, oracle, rings, az):
    cloud = _roi_cloud(oracle, 4100, rings, az)
    zones = reference_front_zones()
    want = oracle.zone_split(cloud, zones)
    assert sum(len(s) for _, s in want) > 0
    with CloudMerger(max_sensors=1, max_points_per_sensor=rings * az) as cm:
        cm.set_zones(zones)
        buf = cm.upload(cloud)
        cm.dev_zone_split(buf.ptr, len(cloud))
        _check(cm.zone_out(), want, "device form")
        _check(cm.zone_split(cloud), want, "host form")


def test_zone_edges_overlap_gap_and_invalid(gpu_ok, oracle):
    """Window ends are inclusive and shared (a point on x = 19 belongs to both neighbours); z in (zg, zg + 0.01) belongs
    to neither z window; non-finite coordinates never pass; a non-finite intensity only matters to an intensity stage;
    negative windows keep the outside."""
    f32 = np.float32
    zones = reference_front_zones() + [[(3, 10.0, 20.0, 0)], [(1, -1.0, 1.0, 1), (0, 0.0, 100.0, 0)], []]
    zlo = float(f32(np.float64(f32(2.0)) + 0.01))
    rows = [(19.0, 0.0, 0.0, 1.0), (30.0, 1.0, 2.0, 2.0), (4.0, 0.0, 1.5, 3.0), (-4.0, 0.0, 0.3, 4.0), (-15.0, 0.0, -0.5, 5.0),
            (60.0, 0.0, 2.5, 6.0), (25.0, 0.0, 2.005, 7.0), (25.0, 0.0, zlo, 8.0), (25.0, 0.0, np.nextafter(f32(zlo), f32(0)), 9.0),
            (np.nan, 0.0, 0.0, 10.0), (25.0, np.inf, 0.0, 15.0), (25.0, 0.0, -np.inf, 15.0), (25.0, 0.0, 1.0, np.nan),
            (25.0, 2.0, 1.0, 15.0), (25.0, -1.0, 1.0, 15.0), (-20.0, 5.0, 0.0, 12.0), (60.000004, 0.0, 0.0, 15.0)]
    cloud = np.array(rows, np.float32)
    rng = np.random.default_rng(7)
    filler = np.column_stack([rng.uniform(-16, 61, 5000), rng.uniform(-5, 5, 5000), rng.uniform(-0.5, 3.0, 5000),
                              rng.uniform(0, 255, 5000)]).astype(np.float32)
    cloud = np.concatenate([cloud, filler, cloud])
    want = oracle.zone_split(cloud, zones)
    # the shared end point: index 0 (x = 19) is in both the mid2 (zones 2, 3) and the mid (zones 4, 5) x windows
    assert 0 in want[2][1] and 0 in want[4][1]
    assert len(want[-1][1]) == len(cloud)  # a chain without stages is no filter: it keeps everything
    with CloudMerger(max_sensors=1, max_points_per_sensor=len(cloud)) as cm:
        cm.set_zones(zones)
        _check(cm.zone_split(cloud, capacity=len(zones) * len(cloud)), want, "edges")


def test_zone_split_of_last_crop_and_degenerate(gpu_ok, oracle):
    """xyzi_dev = NULL splits the merged cropped cloud of the last transform_crop run; empty inputs and empty zones work;
    overlapping zones may exceed the input size; a too small caller buffer is an error."""
    S, rings, az = 2, 32, 256
    zones = reference_front_zones()
    with CloudMerger(max_sensors=S, max_points_per_sensor=rings * az) as cm:
        items, merged = [], []
        for s in range(S):
            m = synth.extrinsic(s, S)
            cm.set_extrinsic(s, m)
            raw = synth.lidar_cloud(4200, s, 0, rings, az)
            buf = cm.upload(raw)
            items.append((buf.ptr, len(raw), make_layout(), s, 0))
            cur = oracle.transform(raw, m[:3].reshape(-1))
            for (axis, lo, hi, neg) in ROI_PASSES:
                cur = np.ascontiguousarray(cur[oracle.passthrough(cur, axis, lo, hi, bool(neg))])
            merged.append(cur)
        merged = np.concatenate(merged)
        cm.set_crop(ROI_PASSES)
        cm.set_zones(zones)
        cm.dev_transform_crop(cm.make_segments(items))
        cm.dev_zone_split()
        _check(cm.zone_out(), oracle.zone_split(merged, zones), "zones of the cropped merge")
        # nothing in, nothing out
        got = cm.zone_split(np.zeros((0, 4), np.float32))
        assert all(len(s) == 0 for _, s in got)
        # zones that match nothing
        cm.set_zones([[(0, 1000.0, 2000.0, 0)], [(2, 5.0, 4.0, 0)]])
        got = cm.zone_split(merged)
        assert [len(s) for _, s in got] == [0, 0]
        # sixteen zones that each keep everything: 16 x the input (the library grows its output arrays); a caller buffer
        # that is too small gets CM_E_CAPACITY
        cm.set_zones([[] for _ in range(16)])
        got = cm.zone_split(merged, capacity=16 * len(merged))
        assert all((s == np.arange(len(merged))).all() for _, s in got)
        with pytest.raises(CloudMergerError) as e:
            cm.zone_split(merged, capacity=len(merged))
        assert e.value.code == _lib.CM_E_CAPACITY
        with pytest.raises((CloudMergerError, ValueError)):
            cm.set_zones([[(0, 0.0, 1.0, 0)] * 5])  # more stages than a zone holds


def test_zone_partition_properties_full_size(gpu_ok):
    """Size-independent checks at 2 Mi points: windows that tile the x axis without shared end points partition the cloud
    (every finite in-range point lands in exactly one zone), order is preserved, coordinates are copied bit for bit."""
    n = 1 << 21
    rng = np.random.default_rng(11)
    cloud = np.column_stack([rng.uniform(-15, 60, n), rng.uniform(-5, 5, n), rng.uniform(-0.5, 3, n),
                             rng.uniform(0, 255, n)]).astype(np.float32)
    edges = np.array([-15, -4, 4, 19, 30, 60.5], np.float32)
    zones = [[(0, float(edges[i]), float(np.nextafter(edges[i + 1], np.float32(-1e9))), 0)] for i in range(5)]
    with CloudMerger(max_sensors=1, max_points_per_sensor=n) as cm:
        cm.set_zones(zones)
        buf = cm.upload(cloud)
        cm.dev_zone_split(buf.ptr, n)
        got = cm.zone_out()
    seen = np.zeros(n, np.int32)
    for z, (gx, gs) in enumerate(got):
        assert (np.diff(gs.astype(np.int64)) > 0).all(), "zone %d: input order not preserved" % z
        assert_bit_equal(gx, cloud[gs], "zone %d coordinates" % z)
        assert ((cloud[gs, 0] >= edges[z]) & (cloud[gs, 0] < edges[z + 1])).all()
        seen[gs] += 1
    assert (seen == 1).all(), "not a partition"


def test_zones_and_outlier_golden_fixture(gpu_ok):
    """The committed fixture tests/golden/zones_outlier.json (plain numpy windows / O(n^2) count) through the C ABI."""
    import json
    import os
    doc = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "zones_outlier.json")))
    x = np.array([[np.nan if v == "nan" else v for v in row] for row in doc["xyzi"]], np.float32)
    with CloudMerger(max_sensors=1, max_points_per_sensor=4096, max_batch_points=4096) as cm:
        cm.set_zones([[tuple(p) for p in z] for z in doc["zones"]])
        got = cm.zone_split(x, capacity=len(doc["zones"]) * len(x))
        for (gx, gi), want in zip(got, doc["zone_indices"]):
            assert gi.tolist() == want
            assert gx.tobytes() == x[want].tobytes()
        for m, want in doc["outlier_kept_by_min_pts"].items():
            kx, ki = cm.radius_outlier(x, doc["radius"], int(m))
            assert ki.tolist() == want
            assert kx.tobytes() == x[want].tobytes()
