"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle on the same seeded inputs.

Bar (north_star): bit-exact surviving-point indices and coordinates, voxel membership, counts and output order;
centroids within 1e-5 relative of the double-precision oracle (and, because the radix sort is stable, bit-equal to the
oracle's float accumulation in ascending point order).
"""
import numpy as np
import pytest

from cloud_merger_b200 import CloudMerger, CloudMergerError, _lib, make_layout, synth

from helpers import assert_bit_equal, assert_centroids_close, cloud_dict, known_answer

pytestmark = pytest.mark.gpu

LAYOUTS = {
    "packed16": (16, 0, 4, 8, 12),
    "pcl32": (32, 0, 4, 8, 16),
    "aligned48": (48, 8, 12, 16, 36),
    "noint20": (20, 4, 8, 12, -1),
    "velodyne22": (22, 0, 4, 8, 12),
    "livox18": (18, 0, 4, 8, 12),
    "odd19": (19, 3, 7, 11, 15),
    "bytes100": (101, 1, 5, 9, 13),
}


def _layout(t, dense=1):
    return make_layout(t[0], t[1], t[2], t[3], t[4], dense)


def _upload_segments(cm, frames, layout_of):
    """frames: list (per frame) of list (per sensor) of (xyzi, is_dense). Returns (segments, oracle cloud dicts per frame)."""
    items, per_frame = [], []
    for f, sensors in enumerate(frames):
        cds = []
        for s, (xyzi, dense) in enumerate(sensors):
            lt = layout_of(f, s)
            data = synth.pack_cloud(xyzi, *lt)
            buf = cm.upload(data if len(data) else np.zeros(16, np.uint8))
            items.append((buf.ptr, len(xyzi), _layout(lt, dense), s, f))
            cds.append(dict(data=data, n_points=len(xyzi), point_step=lt[0], off_x=lt[1], off_y=lt[2], off_z=lt[3],
                            off_i=lt[4], is_dense=int(dense), m=cm.get_extrinsic(s)))
        per_frame.append(cds)
    return cm.make_segments(items), per_frame


def _check_survivors(out, frames_info, oracle_frames):
    for f, (fi, o) in enumerate(zip(frames_info, oracle_frames)):
        sx = out["survivor_xyzi"][fi.survivor_begin:fi.survivor_end]
        ss = out["survivor_src"][fi.survivor_begin:fi.survivor_end]
        assert len(ss) == o["n_survivors"], "frame %d: %d survivors vs oracle %d" % (f, len(ss), o["n_survivors"])
        assert (ss == o["survivor_src"]).all(), "frame %d: surviving-point indices differ" % f
        assert_bit_equal(sx, o["survivor_xyzi"], "frame %d survivor coordinates" % f)


def _check_voxels(out, frames_info, oracle_frames, check_membership=True):
    worst = 0.0
    idx_bits = out["key_idx_bits"]
    mask = (1 << idx_bits) - 1 if idx_bits < 64 else (1 << 64) - 1
    for f, (fi, o) in enumerate(zip(frames_info, oracle_frames)):
        v0, v1 = fi.voxel_begin, fi.voxel_end
        assert v1 - v0 == o["n_voxels"], "frame %d: %d voxels vs oracle %d" % (f, v1 - v0, o["n_voxels"])
        if o["n_survivors"]:
            assert list(fi.min_b) == o["min_b"].tolist() and list(fi.div_b) == o["div_b"].tolist(), "frame %d grid" % f
            assert fi.pcl_overflow == o["pcl_overflow"]
        assert (out["voxel_idx"][v0:v1].astype(np.int64) == o["idx"]).all(), "frame %d: voxel ids / order differ" % f
        assert (out["voxel_count"][v0:v1] == o["count"]).all(), "frame %d: voxel counts differ" % f
        assert_bit_equal(out["voxel_xyzi"][v0:v1], o["centroid"], "frame %d centroid vs float oracle (asc. index)" % f)
        worst = max(worst, assert_centroids_close(out["voxel_xyzi"][v0:v1], o["centroid_f64"], "frame %d centroid" % f))
        if check_membership:
            s0, s1 = fi.survivor_begin, fi.survivor_end
            keys = out["sorted_key"][s0:s1]
            pts = out["sorted_point"][s0:s1].astype(np.int64)
            assert (np.diff(keys.astype(np.float64)) >= 0).all() and (keys[1:] >= keys[:-1]).all(), "keys not sorted"
            assert ((keys >> np.uint64(idx_bits)) == f).all() if idx_bits < 64 else True
            member = np.full(s1 - s0, -2, np.int64)
            member[pts - s0] = (keys & np.uint64(mask)).astype(np.int64)
            assert (member == o["point_idx"]).all(), "frame %d: voxel membership differs" % f
            # stable sort: inside a voxel the points are in ascending index order
            same = keys[1:] == keys[:-1]
            assert (pts[1:][same] > pts[:-1][same]).all(), "sort is not stable"
    return worst


def _oracle_frames(oracle, per_frame, passes, leaf, min_points, downsample_all=True):
    return [oracle.merge_frame(cds, passes, leaf, min_points, downsample_all, True) for cds in per_frame]


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("min_points", [1, 2])
def test_known_answer_host_path(gpu_ok, min_points):
    """The hand-computed vector of tests/golden/known_answer.json through submit_cloud / merge_frame."""
    ka = known_answer()
    e = ka["expected"]
    with CloudMerger(max_sensors=2, max_points_per_sensor=16) as cm:
        cm.set_extrinsic(0, np.array(ka["extrinsics_row_major_3x4"]["A"], np.float32))
        cm.set_extrinsic(1, np.array(ka["extrinsics_row_major_3x4"]["B"], np.float32))
        cm.set_crop([tuple(p) for p in ka["crop_passes"]])
        cm.set_voxel(ka["leaf"], min_points, True)
        cm.submit_cloud(0, ka["A"], 4, make_layout(is_dense=1), stamp=5)
        cm.submit_cloud(1, ka["B"], 4, make_layout(is_dense=0), stamp=9)
        r = cm.merge_frame(capacity=8)
    assert r.survivor_src.tolist() == e["survivor_src"]
    assert_bit_equal(r.survivor_xyzi[:, :3], np.array(e["survivor_xyz"], np.float32), "survivors")
    assert list(r.info.min_b) == e["min_b"] and list(r.info.max_b) == e["max_b"] and list(r.info.div_b) == e["div_b"]
    ex = e["min_points_%d" % min_points]
    assert r.voxel_idx.tolist() == ex["idx"] and r.voxel_count.tolist() == ex["count"]
    assert_bit_equal(r.voxel_xyzi, np.array(ex["centroid"], np.float32), "centroids")
    assert r.used_mask == 3 and r.stamp == 9 and not r.info.pcl_overflow


@pytest.mark.parametrize("name", list(LAYOUTS))
@pytest.mark.parametrize("dense", [1, 0])
def test_transform_crop_layouts(gpu_ok, oracle, name, dense):
    """K1 alone on every record layout (16-byte fast path, PCL 32-byte, 4-byte aligned, TMA-staged unaligned, byte
    fallback), ragged sizes around the 2048-point tile, with and without non-finite points."""
    lt = LAYOUTS[name]
    sizes = [0, 1, 31, 2047, 2048, 2049, 5000, 12345]
    rng = np.random.default_rng(42)
    sensors = []
    for s, n in enumerate(sizes):
        p = synth.lidar_cloud(77, s, 0, 1, max(n, 1))[:n]
        if not dense and n:
            bad = rng.random(n) < 0.02
            p[bad, rng.integers(0, 3)] = np.nan
            if n > 10:
                p[3, 1] = np.inf
        sensors.append((p, dense))
    with CloudMerger(max_sensors=len(sizes), max_batch_points=sum(sizes), max_batch_frames=1) as cm:
        for s in range(len(sizes)):
            cm.set_extrinsic(s, synth.extrinsic(s, len(sizes)))
        cm.set_crop(synth.ROI_BOX)
        segs, per_frame = _upload_segments(cm, [sensors], lambda f, s: lt)
        cm.dev_transform_crop(segs)
        out = cm.fetch_batch_outputs()
        o = _oracle_frames(oracle, per_frame, synth.ROI_BOX, [0.1] * 3, 1)
        _check_survivors(out, out["frames"], o)
        assert out["stats"].survivors == o[0]["n_survivors"] > 100
        assert out["stats"].points_in == sum(sizes)


def test_transform_crop_mixed_layouts_and_pass_kinds(gpu_ok, oracle):
    """Sensors with different layouts in one frame; negative pass and an intensity pass; limits sitting exactly on data."""
    names = ["packed16", "velodyne22", "pcl32", "livox18", "odd19", "aligned48"]
    sensors = [(synth.lidar_cloud(5, s, 1, 16, 512, nan_frac=0.01), 0) for s in range(len(names))]
    some = np.sort(sensors[0][0][:, 2])
    passes = [(2, float(some[100]), float(some[-100]), 0), (0, -2.0, 2.0, 1), (3, 10.0, 200.0, 0), (1, -40.0, 40.0, 0)]
    with CloudMerger(max_sensors=len(names), max_batch_points=len(names) * 8192) as cm:
        for s in range(len(names)):
            cm.set_extrinsic(s, synth.extrinsic(s, len(names)))
        cm.set_crop(passes)
        segs, per_frame = _upload_segments(cm, [sensors], lambda f, s: LAYOUTS[names[s]])
        cm.dev_transform_crop(segs)
        out = cm.fetch_batch_outputs()
        o = _oracle_frames(oracle, per_frame, passes, [0.1] * 3, 1)
        _check_survivors(out, out["frames"], o)
        assert 0 < out["stats"].survivors < out["stats"].points_in


def test_no_crop_keeps_everything_and_voxel_skips_invalid(gpu_ok, oracle):
    """Zero passes = plain cloud_fusion concat (CloudFusionNode.h:59-72): every point survives, non-finite ones included
    (left untransformed for a non-dense cloud); VoxelGrid then skips them like PCL does for !is_dense."""
    sensors = [(synth.lidar_cloud(6, s, 0, 8, 700, nan_frac=0.03), 0) for s in range(3)]
    with CloudMerger(max_sensors=3, max_batch_points=3 * 5600) as cm:
        for s in range(3):
            cm.set_extrinsic(s, synth.extrinsic(s, 3))
        cm.set_crop([])
        cm.set_voxel(0.5, 1, True)
        segs, per_frame = _upload_segments(cm, [sensors], lambda f, s: LAYOUTS["packed16"])
        cm.run_batch(segs)
        out = cm.fetch_batch_outputs()
        o = _oracle_frames(oracle, per_frame, [], [0.5] * 3, 1)
        assert out["stats"].survivors == 3 * 5600
        _check_survivors(out, out["frames"], o)
        assert np.isnan(out["survivor_xyzi"]).any()
        _check_voxels(out, out["frames"], o, check_membership=False)
        # membership: invalid points carry the sentinel frame and are sorted to the end
        n_bad = int((~np.isfinite(out["survivor_xyzi"][:, :3]).all(axis=1)).sum())
        tail = out["sorted_point"][len(out["sorted_point"]) - n_bad:]
        assert (~np.isfinite(out["survivor_xyzi"][tail, :3]).all(axis=1)).all()


@pytest.mark.parametrize("leaf,min_points", [(1.0, 1), (0.5, 2), (0.2, 3), (0.1, 2), (0.05, 1), (0.02, 1), (0.01, 2)])
def test_voxelgrid_only_leaf_sweep(gpu_ok, oracle, leaf, min_points):
    """VoxelGrid alone (cm_dev_voxelgrid) across occupancy regimes; the fine leaves exceed PCL's INT32 cell limit and run
    on 64-bit keys (oracle force64)."""
    x = synth.uniform_cloud(21, 300000, extent=(200.0, 200.0, 10.0))
    x[:50000] = synth.uniform_cloud(22, 50000, extent=(6.0, 6.0, 2.0))  # a dense clump: long runs
    with CloudMerger(max_batch_points=len(x)) as cm:
        cm.set_voxel(leaf, min_points, True)
        buf = cm.upload(x)
        cm.dev_voxelgrid(buf.ptr, len(x))
        out = cm.fetch_batch_outputs()
    o = oracle.voxelgrid(x, [leaf] * 3, min_points, True, force64=True)
    o.update(n_voxels=o["n"], n_survivors=len(x))
    assert out["stats"].pcl_overflow == int(o["pcl_overflow"])
    assert out["key_bytes"] == (4 if out["stats"].key_bits <= 32 else 8)
    worst = _check_voxels(out, out["frames"], [o])
    assert worst <= 1e-5


def test_voxelgrid_degenerate_inputs(gpu_ok, oracle):
    with CloudMerger(max_batch_points=4096) as cm:
        cm.set_voxel(0.1, 1, True)
        # a single point, all points identical, all points invalid, empty
        for x in (np.array([[1, 2, 3, 4]], np.float32), np.tile(np.array([[0.5, -0.5, 0.25, 9]], np.float32), (3000, 1)),
                  np.full((100, 4), np.nan, np.float32), np.zeros((0, 4), np.float32)):
            buf = cm.upload(x if len(x) else np.zeros((1, 4), np.float32))
            cm.dev_voxelgrid(buf.ptr, len(x), is_dense=False)
            out = cm.fetch_batch_outputs()
            o = oracle.voxelgrid(x, [0.1] * 3, 1, True, force64=True, is_dense=False)
            assert out["stats"].voxels_out == o["n"]
            if o["n"]:
                assert (out["voxel_count"] == o["count"]).all()
                assert_bit_equal(out["voxel_xyzi"], o["centroid"], "centroid")


def test_downsample_all_false_and_pcl_record_layout(gpu_ok, oracle):
    x = synth.uniform_cloud(3, 20000, extent=(10.0, 10.0, 2.0))
    with CloudMerger(max_batch_points=len(x), out_point_step=32) as cm:
        cm.set_voxel(0.25, 2, False)
        buf = cm.upload(x)
        cm.dev_voxelgrid(buf.ptr, len(x))
        out = cm.fetch_batch_outputs()
    o = oracle.voxelgrid(x, [0.25] * 3, 2, False, force64=True)
    rec = out["voxel_records"]
    assert rec.shape[1] == 8 and (rec[:, 3] == 1.0).all() and (rec[:, 5:] == 0).all() and (rec[:, 4] == 0).all()
    assert_bit_equal(out["voxel_xyzi"], o["centroid"], "centroid (xyz only)")


def test_batch_of_frames_matches_per_frame_oracle(gpu_ok, oracle):
    """F frames x S sensors in one device-resident batch: frame-major keys, per-frame grids, per-frame slices."""
    F, S = 5, 3
    frames = [[(synth.lidar_cloud(31, s, f, 16, 300 + 37 * f + 11 * s), 1) for s in range(S)] for f in range(F)]
    frames[2][1] = (np.zeros((0, 4), np.float32), 1)          # an empty cloud inside a frame
    frames[3] = [(np.full((10, 4), 1e6, np.float32), 1)] * S  # a frame whose points are all cropped away
    total = sum(len(p) for fr in frames for p, _ in fr)
    with CloudMerger(max_sensors=S, max_batch_points=total, max_batch_frames=F) as cm:
        for s in range(S):
            cm.set_extrinsic(s, synth.extrinsic(s, S))
        cm.set_crop(synth.ROI_BOX)
        cm.set_voxel(0.1, 2, True)
        segs, per_frame = _upload_segments(cm, frames, lambda f, s: LAYOUTS["packed16" if (f + s) % 2 else "pcl32"])
        cm.run_batch(segs)
        out = cm.fetch_batch_outputs()
        o = _oracle_frames(oracle, per_frame, synth.ROI_BOX, [0.1] * 3, 2)
        assert out["stats"].frames == F and out["key_bytes"] == 4
        _check_survivors(out, out["frames"], o)
        _check_voxels(out, out["frames"], o)
        assert out["frames"][3].survivor_begin == out["frames"][3].survivor_end
        assert cm.launch_count() >= 5


@pytest.mark.parametrize("cfg,nan_frac", [("cfg1", 0.0), ("cfg2", 0.0), ("cfg2", 0.005)])
def test_config_shapes_host_path(gpu_ok, oracle, cfg, nan_frac):
    """BASELINE configs 1 and 2 at full size through the host path (submit per sensor, merge), mixed record layouts."""
    c = synth.CONFIGS[cfg]
    clouds, mats = synth.frame_clouds(cfg, 1000 * int(cfg[-1]), 0, nan_frac)
    S, n = c["sensors"], len(clouds[0])
    lts = [LAYOUTS["packed16"], LAYOUTS["pcl32"], LAYOUTS["velodyne22"], LAYOUTS["livox18"]]
    dense = 0 if nan_frac else 1
    with CloudMerger(max_sensors=S, max_points_per_sensor=n, max_point_step=32) as cm:
        cds = []
        for s in range(S):
            cm.set_extrinsic(s, mats[s])
            lt = lts[s % len(lts)]
            data = synth.pack_cloud(clouds[s], *lt)
            cm.submit_cloud(s, data, n, _layout(lt, dense), stamp=100 + s)
            cds.append(cloud_dict(clouds[s], mats[s][:3], dense, *lt))
        cm.set_crop(c["passes"])
        cm.set_voxel(c["leaf"], c["min_points"], True)
        r = cm.merge_frame(capacity=S * n)
        st = cm.stats()
    o = oracle.merge_frame(cds, c["passes"], [c["leaf"]] * 3, c["min_points"], True, True)
    assert len(r.survivor_src) == o["n_survivors"] and (r.survivor_src == o["survivor_src"]).all()
    assert_bit_equal(r.survivor_xyzi, o["survivor_xyzi"], "survivors")
    assert len(r.voxel_idx) == o["n_voxels"] > 1000
    assert (r.voxel_idx.astype(np.int64) == o["idx"]).all() and (r.voxel_count == o["count"]).all()
    assert_bit_equal(r.voxel_xyzi, o["centroid"], "centroid vs float oracle")
    assert_centroids_close(r.voxel_xyzi, o["centroid_f64"], "centroid vs f64 oracle")
    assert r.stamp == 100 + S - 1 and r.used_mask == (1 << S) - 1
    assert st.survivors == o["n_survivors"] and st.voxels_out == o["n_voxels"]


def test_pipelined_frames_and_optional_sensor(gpu_ok, oracle):
    """Three frames in flight (merge_frame_async / wait_frame); a sensor missing from one frame is skipped like the
    reference's optional top sensor (pc_preprocessing_main.cpp:134-136)."""
    S, n = 3, 4000
    with CloudMerger(max_sensors=S, max_points_per_sensor=n, frames_in_flight=3) as cm:
        mats = [synth.extrinsic(s, S) for s in range(S)]
        for s in range(S):
            cm.set_extrinsic(s, mats[s])
        cm.set_crop(synth.ROI_BOX)
        cm.set_voxel(0.2, 1, True)
        tickets, expect = [], []
        for f in range(3):
            cds = []
            for s in range(S):
                if f == 1 and s == 2:
                    continue
                p = synth.lidar_cloud(55, s, f, 8, n // 8)
                cm.submit_cloud(s, p, len(p), make_layout(), stamp=f)
                cds.append(cloud_dict(p, mats[s][:3]))
            tickets.append(cm.merge_frame_async())
            expect.append(oracle.merge_frame(cds, synth.ROI_BOX, [0.2] * 3, 1, True, True))
        with pytest.raises(CloudMergerError) as e:   # every slot is now in flight
            cm.submit_cloud(0, np.zeros((1, 4), np.float32), 1, make_layout())
        assert e.value.code == _lib.CM_E_CAPACITY
        for f in (0, 1, 2):
            r = cm.wait_frame(tickets[f], capacity=S * n)
            o = expect[f]
            assert r.used_mask == (0b011 if f == 1 else 0b111)
            assert (r.survivor_src == o["survivor_src"]).all()
            assert (r.voxel_idx.astype(np.int64) == o["idx"]).all() and (r.voxel_count == o["count"]).all()
            assert_bit_equal(r.voxel_xyzi, o["centroid"], "frame %d" % f)


def test_concurrent_sensor_callbacks_and_frame_graphs(gpu_ok, oracle):
    """The reference runs its six sensor callbacks on ros::AsyncSpinner(6) and fuses on the main thread
    (pc_preprocessing_main.cpp:513, :574-578): cm_submit_cloud from one thread per sensor, concurrently, then the merge.
    Twelve frames of the same shape also take the host path through its plain / captured / replayed-graph stages; every
    frame must match the oracle."""
    import threading
    S, rings, az = 6, 16, 128
    n = rings * az
    passes, leaf, mp = synth.ROI_BOX, 0.1, 2
    with CloudMerger(max_sensors=S, max_points_per_sensor=n, frames_in_flight=2) as cm:
        mats = [synth.extrinsic(s, S) for s in range(S)]
        for s in range(S):
            cm.set_extrinsic(s, mats[s])
        cm.set_crop(passes)
        cm.set_voxel(leaf, mp, True)
        for f in range(12):
            clouds = [synth.lidar_cloud(77, s, f, rings, az) for s in range(S)]
            errs = []

            def callback(s):
                try:
                    cm.submit_cloud(s, clouds[s], n, make_layout(), stamp=100 + f)
                except Exception as e:  # noqa: BLE001
                    errs.append(e)
            th = [threading.Thread(target=callback, args=(s,)) for s in range(S)]
            for t in th:
                t.start()
            for t in th:
                t.join()
            assert not errs, errs
            r = cm.merge_frame(capacity=S * n)
            o = oracle.merge_frame([cloud_dict(c, m[:3]) for c, m in zip(clouds, mats)], passes, [leaf] * 3, mp, True, True)
            assert r.used_mask == (1 << S) - 1 and r.stamp == 100 + f
            assert (r.survivor_src == o["survivor_src"]).all(), "frame %d" % f
            assert_bit_equal(r.survivor_xyzi, o["survivor_xyzi"], "frame %d survivors" % f)
            assert (r.voxel_idx.astype(np.int64) == o["idx"]).all() and (r.voxel_count == o["count"]).all(), "frame %d" % f
            assert_bit_equal(r.voxel_xyzi, o["centroid"], "frame %d centroids" % f)


def test_submit_clouds_pinned_matches_single_submits(gpu_ok, oracle):
    """cm_submit_clouds_pinned == one cm_submit_cloud_pinned per cloud, whether or not the clouds are adjacent in host
    memory (adjacent clouds of consecutive sensors travel as one copy), with ragged sizes and a skipped sensor id."""
    from cloud_merger_b200 import host_alloc
    S, cap = 4, 16 * 512
    sizes = [cap, cap - 37 * 16 // 16, 1234, cap]
    passes, leaf, mp = [(2, -0.5, 3.0, 0), (0, -15.0, 60.0, 0)], 0.1, 2
    arena, addr = host_alloc(S * cap * 16 + 4096)
    clouds, addrs, off = [], [], 0
    for s in range(S):
        c = synth.lidar_cloud(555, s, 0, 16, 512)[:sizes[s]]
        if s == 3:
            off += 256  # a gap: the last cloud is not adjacent to its predecessor
        arena[off:off + c.nbytes] = c.view(np.uint8).reshape(-1)
        clouds.append(c); addrs.append(addr + off)
        off += c.nbytes
    with CloudMerger(max_sensors=6, max_points_per_sensor=cap, frames_in_flight=2) as cm:
        ids = [0, 1, 2, 4]  # sensor 3 is never submitted: 0-1-2 form one run, 4 stands alone (and is not adjacent anyway)
        for s, sid in enumerate(ids):
            cm.set_extrinsic(sid, synth.extrinsic(s, S))
        cm.set_crop(passes); cm.set_voxel(leaf, mp, True)
        cm.submit_clouds_pinned(ids, addrs, sizes, [make_layout()] * S, stamps=[7, 8, 9, 10])
        got = cm.merge_frame(capacity=S * cap)
        for s, sid in enumerate(ids):
            cm.submit_cloud(sid, clouds[s], sizes[s], make_layout(), stamp=7 + s)
        ref = cm.merge_frame(capacity=S * cap)
        mats = [cm.get_extrinsic(sid) for sid in ids]
    assert got.stamp == ref.stamp == 10
    assert (got.survivor_src == ref.survivor_src).all() and len(got.survivor_src) > 1000
    assert_bit_equal(got.survivor_xyzi, ref.survivor_xyzi, "survivors")
    assert (got.voxel_idx == ref.voxel_idx).all() and (got.voxel_count == ref.voxel_count).all()
    assert_bit_equal(got.voxel_xyzi, ref.voxel_xyzi, "centroids")
    o = oracle.merge_frame([cloud_dict(c, m) for c, m in zip(clouds, mats)], passes, [leaf] * 3, mp, True, True)
    assert (got.survivor_src == o["survivor_src"]).all()
    assert (got.voxel_idx.astype(np.int64) == o["idx"]).all()


def test_pcl_overflow_modes(gpu_ok, oracle):
    """Leaf too small for the extent: PCL 1.8.1 returns the input unchanged (mode 1); the default carries on in 64 bits."""
    p = synth.uniform_cloud(8, 5000, extent=(300.0, 300.0, 20.0))
    for pcl_like in (False, True):
        with CloudMerger(max_sensors=1, max_points_per_sensor=len(p)) as cm:
            cm.set_crop([])
            cm.set_voxel(0.01, 1, True)
            cm.set_overflow_mode(pcl_like)
            cm.submit_cloud(0, p, len(p), make_layout())
            r = cm.merge_frame(capacity=len(p))
        assert r.info.pcl_overflow
        if pcl_like:
            o = oracle.voxelgrid(p, [0.01] * 3, 1, True, force64=False)
            assert o["returned_input"] and len(r.voxel_xyzi) == len(p)
            assert_bit_equal(r.voxel_xyzi, p, "output = input")
        else:
            o = oracle.voxelgrid(p, [0.01] * 3, 1, True, force64=True)
            assert (r.voxel_idx.astype(np.int64) == o["idx"]).all()
            assert_bit_equal(r.voxel_xyzi, o["centroid"], "64-bit path")


def test_error_behaviour(gpu_ok):
    with CloudMerger(max_sensors=2, max_points_per_sensor=100, max_point_step=16) as cm:
        with pytest.raises(CloudMergerError) as e:
            cm.submit_cloud(0, np.zeros((101, 4), np.float32), 101, make_layout())
        assert e.value.code == _lib.CM_E_CAPACITY
        with pytest.raises(CloudMergerError) as e:
            cm.submit_cloud(0, np.zeros((10, 8), np.float32), 10, make_layout(32, 0, 4, 8, 16))
        assert e.value.code == _lib.CM_E_CAPACITY
        with pytest.raises(CloudMergerError) as e:
            cm.submit_cloud(0, np.zeros((10, 4), np.float32), 10, make_layout(16, 0, 4, 14, 12))
        assert e.value.code == _lib.CM_E_INVALID
        with pytest.raises(CloudMergerError) as e:
            cm.submit_cloud(5, np.zeros((10, 4), np.float32), 10, make_layout())
        assert e.value.code == _lib.CM_E_INVALID
        with pytest.raises(CloudMergerError) as e:
            cm.merge_frame(capacity=10)
        assert e.value.code == _lib.CM_E_NOT_READY
        cm.set_crop([])
        cm.set_voxel(0.1, 1, True)
        cm.submit_cloud(0, np.arange(40, dtype=np.float32).reshape(10, 4), 10, make_layout())
        with pytest.raises(CloudMergerError) as e:   # caller buffer too small: reports instead of overrunning
            cm.merge_frame(capacity=3)
        assert e.value.code == _lib.CM_E_CAPACITY


def test_full_size_batch_properties(gpu_ok, oracle):
    """cfg 2 at bench size (16 frames x 4 x 128k points) checked through size-independent properties, plus an exact
    comparison of two of its frames with the oracle."""
    c = synth.CONFIGS["cfg2"]
    F, S = 16, c["sensors"]
    n = c["rings"] * c["azimuth"]
    with CloudMerger(max_sensors=S, max_batch_points=F * S * n, max_batch_frames=F) as cm:
        mats = [synth.extrinsic(s, S) for s in range(S)]
        for s in range(S):
            cm.set_extrinsic(s, mats[s])
        cm.set_crop(c["passes"])
        cm.set_voxel(c["leaf"], c["min_points"], True)
        frames = [[(synth.lidar_cloud(2000, s, f, c["rings"], c["azimuth"]), 1) for s in range(S)] for f in range(F)]
        segs, per_frame = _upload_segments(cm, frames, lambda f, s: LAYOUTS["packed16"])
        cm.run_batch(segs)
        out = cm.fetch_batch_outputs()
        st = out["stats"]
        assert st.points_in == F * S * n and st.frames == F and st.device_error == 0
        keys = out["sorted_key"]
        assert (keys[1:] >= keys[:-1]).all(), "keys must be globally sorted (frame-major)"
        assert np.array_equal(np.sort(out["sorted_point"]), np.arange(st.survivors, dtype=np.uint32)), "a permutation"
        # every voxel's count is the length of its run of equal keys; dropped runs are the short ones
        uniq, cnt = np.unique(keys, return_counts=True)
        kept = cnt >= c["min_points"]
        assert kept.sum() == st.voxels_out and (out["voxel_count"] == cnt[kept]).all()
        # centroids lie inside their voxel (up to rounding at the faces)
        fi = out["frames"]
        for f in (0, F - 1):
            v0, v1 = fi[f].voxel_begin, fi[f].voxel_end
            idx = out["voxel_idx"][v0:v1].astype(np.int64)
            d0, d1 = fi[f].div_b[0], fi[f].div_b[1]
            ijk = np.stack([idx % d0, (idx // d0) % d1, idx // (d0 * d1)], axis=1) + np.array(fi[f].min_b)
            cell = np.floor(out["voxel_xyzi"][v0:v1, :3].astype(np.float64) / c["leaf"] + 1e-4)
            assert np.abs(cell - ijk).max() <= 1
        o = [oracle.merge_frame(per_frame[f], c["passes"], [c["leaf"]] * 3, c["min_points"], True, True) for f in (0, F - 1)]
        sel = [fi[0], fi[F - 1]]
        sub = dict(out)
        # compare frames 0 and F-1 exactly
        for k, (f, fo) in enumerate(zip((0, F - 1), o)):
            s0, s1, v0, v1 = sel[k].survivor_begin, sel[k].survivor_end, sel[k].voxel_begin, sel[k].voxel_end
            assert (out["survivor_src"][s0:s1] == fo["survivor_src"]).all()
            assert_bit_equal(out["survivor_xyzi"][s0:s1], fo["survivor_xyzi"], "frame %d survivors" % f)
            assert (out["voxel_idx"][v0:v1].astype(np.int64) == fo["idx"]).all()
            assert (out["voxel_count"][v0:v1] == fo["count"]).all()
            assert_bit_equal(out["voxel_xyzi"][v0:v1], fo["centroid"], "frame %d centroids" % f)
        del sub
