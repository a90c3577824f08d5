"""GPU parity tests proper: the CUDA path, called through the C ABI (include/vmap/vmap.h), against
the CPU oracle on identical inputs.  Bar: bit-exact -- per-scan free / occupied key SETS equal,
per-leaf float32 log-odds equal as bit patterns, query statuses equal, projected points equal as
bit patterns.  No tolerance anywhere (the path is integer / exactly-rounded IEEE arithmetic)."""
import os

import numpy as np
import pytest

import oracle
import volumetric_mapping_b200 as vm
from volumetric_mapping_b200 import scenes

pytestmark = pytest.mark.gpu

I4 = np.eye(4)
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# key-set construction variants: bitmask marks in the dense window (default), bitmask marks in the
# hash table's mask columns (window disabled), and emit + radix sort + unique
MODES = ["dense", "table", "sort"]
MODE_IDS = MODES
MODE_KW = {"dense": dict(dedupe_mode=vm.DEDUPE_BITMASK),
           "table": dict(dedupe_mode=vm.DEDUPE_BITMASK, dense_window_bytes=1),
           "sort": dict(dedupe_mode=vm.DEDUPE_SORT)}


def packed(k3):
    return np.sort(oracle.pack_keys(k3))


def leaves_packed(w):
    k, v = w.export_leaves()
    p = oracle.pack_keys(k)
    order = np.argsort(p, kind="stable")
    return p[order], v.view(np.uint32)[order]


def assert_same_keys(gw, ow):
    gf, go = gw.last_scan_keys()
    of, oo = ow.last_scan_keys()
    assert np.array_equal(packed(go), packed(oo)), "occupied key sets differ"
    assert np.array_equal(packed(gf), packed(of)), "free key sets differ"


def assert_same_leaves(gw, ow):
    gk, gv = leaves_packed(gw)
    ok, ov = leaves_packed(ow)
    assert gk.shape == ok.shape and np.array_equal(gk, ok), "leaf key sets differ"
    assert np.array_equal(gv, ov), "leaf log-odds bit patterns differ"


def assert_same_stats(gw, ow, keys=("points_in", "points_valid", "points_in_range", "rays_cast",
                                    "traversals", "occupied_emitted", "unique_free",
                                    "unique_occupied", "longest_ray")):
    g, o = gw.last_scan_stats(), ow.last_scan_stats()
    for k in keys:
        assert g[k] == o[k], (k, g[k], o[k])


def pair(mode="dense", **kw):
    gw = vm.OctomapWorld(**MODE_KW[mode], **kw)
    gw.set_debug_capture(True)
    return gw, oracle.OracleWorld(**kw)


def test_logodds_constants_bit_exact():
    gw, ow = pair()
    assert np.array_equal(gw.logodds_constants().view(np.uint32), ow.logodds_constants().view(np.uint32))
    assert gw.logodds_constants().view(np.uint32).tolist() == [0x3f1e795b, 0xbecf991f, 0xbfff07f4,
                                                               0x405e7867, 0x3f58e883]


# ---------------------------------------------------------------------------------------
# K1: projection
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["reproject_small_a", "reproject_small_b", "reproject_nomissing"])
def test_reproject_matches_cv2_golden(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    gw = vm.OctomapWorld()
    out = gw.reprojectImageTo3D(g["disparity"], g["Q"])
    assert np.array_equal(out.view(np.uint32), g["xyz"].view(np.uint32))


def test_reproject_cfg2_full_size_vs_oracle():
    c = scenes.cfg2_disparity()
    gw = vm.OctomapWorld()
    out = gw.reprojectImageTo3D(c["disparity"], c["Q"])
    exp = oracle.reproject(c["disparity"], c["Q"])
    assert np.array_equal(out.view(np.uint32), exp.view(np.uint32))


# ---------------------------------------------------------------------------------------
# hand-derived cases (same as tests/test_oracle_insertion.py)
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", MODES, ids=MODE_IDS)
def test_single_point_known_answer(mode):
    gw = vm.OctomapWorld(**MODE_KW[mode], resolution=0.1, sensor_max_range=5.0)
    gw.set_debug_capture(True)
    gw.insertPointcloud(I4, [[0.55, 0.05, 0.05]])
    f, o = gw.last_scan_keys()
    assert set(map(tuple, o.tolist())) == {(32773, 32768, 32768)}
    assert set(map(tuple, f.tolist())) == {(32768 + i, 32768, 32768) for i in range(5)}
    st = gw.last_scan_stats()
    assert st["traversals"] == 5 and st["rays_cast"] == 1 and st["points_in_range"] == 1
    k, v = gw.export_leaves()
    d = {tuple(kk): int(vv) for kk, vv in zip(k.tolist(), v.view(np.uint32).tolist())}
    assert d[(32773, 32768, 32768)] == 0x3f1e795b
    assert all(d[(32768 + i, 32768, 32768)] == 0xbecf991f for i in range(5))
    assert len(d) == 6


@pytest.mark.parametrize("mode", MODES, ids=MODE_IDS)
def test_out_of_range_clipped_free_only(mode):
    gw = vm.OctomapWorld(**MODE_KW[mode], resolution=0.1, sensor_max_range=1.0)
    gw.set_debug_capture(True)
    gw.insertPointcloud(I4, [[2.05, 0.0, 0.0]])
    f, o = gw.last_scan_keys()
    assert len(o) == 0
    assert set(map(tuple, f.tolist())) == {(32768 + i, 32768, 32768) for i in range(9)}


@pytest.mark.parametrize("mode", MODES, ids=MODE_IDS)
def test_skip_rule_orderings(mode):
    far = [3.05, 0.05, 0.05]
    far2 = [3.09, 0.09, 0.09]
    a = [0.51, 0.01, 0.05]
    b = [0.59, 0.09, 0.05]
    for pts, rng_ in (([far2, far], 3.06), ([far, far2], 3.06), ([a, b], 5.0), ([b, a], 5.0)):
        gw, ow = pair(mode, resolution=0.1, sensor_max_range=rng_)
        gw.insertPointcloud(I4, pts)
        ow.insertPointcloud(I4, pts)
        assert_same_stats(gw, ow)
        assert_same_keys(gw, ow)


@pytest.mark.parametrize("mode", MODES, ids=MODE_IDS)
def test_empty_and_degenerate_clouds(mode):
    gw, ow = pair(mode, resolution=0.1)
    out = gw.insertPointcloud(I4, np.zeros((0, 3), np.float32))
    assert out.shape == (0, 3) and gw.leaf_count() == 0
    # same-voxel ray -> no free keys, endpoint occupied; out-of-bounds point -> nothing at all
    pts = np.array([[0.01, 0.01, 0.01], [4000.0, 0.0, 0.0], [0.0, -5000.0, 1.0]], np.float32)
    p = dict(resolution=0.1, sensor_max_range=-1.0)
    gw, ow = pair(mode, **p)
    gw.insertPointcloud(I4, pts)
    ow.insertPointcloud(I4, pts)
    assert_same_stats(gw, ow)
    assert_same_keys(gw, ow)
    assert_same_leaves(gw, ow)


@pytest.mark.parametrize("mode", MODES, ids=MODE_IDS)
def test_nan_filter_only_when_not_dense(mode):
    pts = np.array([[0.55, 0.05, 0.05], [np.nan, 0, 0], [0.05, 0.55, 0.05], [np.inf, 0, 0],
                    [0.3, np.nan, 0.2], [-0.7, 0.2, 0.1]], dtype=np.float32)
    for dense in (False, True):
        gw, ow = pair(mode, resolution=0.1)
        g_out = gw.insertPointcloud(I4, pts, is_dense=dense)
        o_out = ow.insertPointcloud(I4, pts, is_dense=dense)
        assert g_out.shape == o_out.shape
        assert np.array_equal(g_out.view(np.uint32), o_out.view(np.uint32))
        assert_same_keys(gw, ow)
        assert_same_leaves(gw, ow)
        assert_same_stats(gw, ow, keys=("points_valid", "rays_cast", "traversals"))


def test_world_frame_cloud_written_back_bit_exact():
    rng = np.random.default_rng(9)
    T = scenes.quat_to_matrix(scenes.random_unit_quat(rng), rng.uniform(-3, 3, 3))
    pts = rng.uniform(-4, 4, (5000, 3)).astype(np.float32)
    gw, ow = pair(resolution=0.1)
    g = gw.insertPointcloud(T, pts)
    o = ow.insertPointcloud(T, pts)
    assert np.array_equal(g.view(np.uint32), o.view(np.uint32))


# ---------------------------------------------------------------------------------------
# randomised differential tests
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", MODES, ids=MODE_IDS)
@pytest.mark.parametrize("seed,res,max_range", [(0, 0.15, 5.0), (1, 0.1, 2.5), (2, 0.2, -1.0),
                                                (3, 0.05, 4.0)])
def test_random_clouds_multi_scan(mode, seed, res, max_range):
    rng = np.random.default_rng(seed)
    gw, ow = pair(mode, resolution=res, sensor_max_range=max_range)
    for scan in range(6):
        T = scenes.quat_to_matrix(scenes.random_unit_quat(rng), rng.uniform(-2, 2, 3))
        n = 3000
        d = rng.normal(size=(n, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        pts = (d * rng.uniform(0.2, 6.0, (n, 1))).astype(np.float32)
        pts[::17] = pts[1::17][: len(pts[::17])] * np.float32(1.001)   # near-duplicate endpoints
        gw.insertPointcloud(T, pts)
        ow.insertPointcloud(T, pts)
        assert_same_stats(gw, ow)
        assert_same_keys(gw, ow)
    assert_same_leaves(gw, ow)


@pytest.mark.parametrize("mode", MODES, ids=MODE_IDS)
def test_clamping_ladder_many_scans_same_cloud(mode):
    # 12 identical scans drive leaves into both clamps; then flip occupied->free ordering
    rng = np.random.default_rng(5)
    gw, ow = pair(mode, resolution=0.2, sensor_max_range=6.0)
    d = rng.normal(size=(500, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    pts = (d * rng.uniform(1.0, 5.0, (500, 1))).astype(np.float32)
    for _ in range(12):
        gw.insertPointcloud(I4, pts)
        ow.insertPointcloud(I4, pts)
    for _ in range(3):
        gw.insertPointcloud(I4, pts * np.float32(1.3))
        ow.insertPointcloud(I4, pts * np.float32(1.3))
    assert_same_leaves(gw, ow)
    _, v = gw.export_leaves()
    bits = set(v.view(np.uint32).tolist())
    assert 0xbfff07f4 in bits and 0x405e7867 in bits   # both clamps reached


@pytest.mark.parametrize("mode", MODES, ids=MODE_IDS)
def test_free_space_filter(mode):
    rng = np.random.default_rng(4)
    kw = dict(resolution=0.1, sensor_max_range=4.0, max_free_space=1.5, min_height_free_space=0.3)
    gw, ow = pair(mode, **kw)
    T = scenes.quat_to_matrix([1, 0, 0, 0], [0.0, 0.0, 1.0])
    d = rng.normal(size=(2000, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    pts = (d * rng.uniform(0.5, 6.0, (2000, 1))).astype(np.float32)
    gw.insertPointcloud(T, pts)
    ow.insertPointcloud(T, pts)
    assert_same_keys(gw, ow)
    assert_same_leaves(gw, ow)
    assert_same_stats(gw, ow, keys=("rays_cast", "traversals", "unique_free", "unique_occupied"))


@pytest.mark.parametrize("mode", MODES, ids=MODE_IDS)
def test_cfg1_pointcloud(mode):
    s = scenes.cfg1_pointcloud()
    gw, ow = pair(mode, **s["params"])
    gw.insertPointcloud(s["T"], s["points"])
    ow.insertPointcloud(s["T"], s["points"])
    st = gw.last_scan_stats()
    assert (st["points_in_range"], st["rays_cast"], st["traversals"], st["unique_free"],
            st["unique_occupied"]) == (5915, 9638, 362225, 93443, 5553)
    assert_same_stats(gw, ow)
    assert_same_keys(gw, ow)
    assert_same_leaves(gw, ow)


@pytest.mark.parametrize("mode", MODES, ids=MODE_IDS)
def test_cfg2_disparity_full_size(mode):
    c = scenes.cfg2_disparity()
    gw, ow = pair(mode, **c["params"])
    gw.insertDisparityImage(c["q"], c["t"], c["disparity"], c["Q"], c["full_image_size"][0])
    ow.insertDisparityImage(c["q"], c["t"], c["disparity"], c["Q"], c["full_image_size"][0])
    assert_same_stats(gw, ow)
    assert_same_keys(gw, ow)
    assert_same_leaves(gw, ow)


def test_disparity_downsampled_q_fixup_and_projected_entry():
    c = scenes.cfg2_disparity(rows=120, cols=160, seed=3)
    gw, ow = pair(**c["params"])
    gw.insertDisparityImage(c["q"], c["t"], c["disparity"], c["Q"], 640.0)   # factor 4
    ow.insertDisparityImage(c["q"], c["t"], c["disparity"], c["Q"], 640.0)
    assert_same_keys(gw, ow)
    assert_same_leaves(gw, ow)
    # insertProjectedDisparityIntoMapImpl on an explicit CV_32FC3 image
    xyz = oracle.reproject(c["disparity"], oracle.fixup_q(c["Q"], 640.0, 160))
    gw2, ow2 = pair(**c["params"])
    gw2.insertProjectedDisparity(c["q"], c["t"], xyz)
    ow2.insertProjectedDisparity(c["q"], c["t"], xyz)
    assert_same_stats(gw2, ow2)
    assert_same_keys(gw2, ow2)
    assert_same_leaves(gw2, ow)


@pytest.mark.parametrize("mode", MODES, ids=MODE_IDS)
def test_cfg3_hdl64_scan_full_size(mode):
    s = scenes.cfg3_scan()
    gw, ow = pair(mode, **s["params"])
    ow.set_faithful_inner_update(False)        # leaf values are unaffected (oracle.h)
    gw.insertPointcloud(s["T"], s["points"])
    ow.insertPointcloud(s["T"], s["points"])
    assert_same_stats(gw, ow)
    assert_same_keys(gw, ow)
    assert_same_leaves(gw, ow)


def test_hdl64_trajectory_five_scans_leaves():
    gw = vm.OctomapWorld(resolution=0.2, sensor_max_range=50.0)
    ow = oracle.OracleWorld(resolution=0.2, sensor_max_range=50.0)
    ow.set_faithful_inner_update(False)
    for i in range(5):
        s = scenes.trajectory_scan(4 * i, resolution=0.2)
        gw.insertPointcloudFast(s["T"], s["points"])
        ow.insertPointcloud(s["T"], s["points"])
        assert_same_stats(gw, ow, keys=("rays_cast", "traversals", "unique_free", "unique_occupied"))
    assert_same_leaves(gw, ow)


@pytest.mark.parametrize("mode", ["dense", "table"])
def test_map_growth_from_tiny_initial_capacity(mode):
    # forces both the table overflow path and the pool growth path
    s = scenes.cfg1_pointcloud()
    gw = vm.OctomapWorld(initial_blocks=1, **MODE_KW[mode], **s["params"])
    ow = oracle.OracleWorld(**s["params"])
    retries = 0
    for k in range(3):
        T = scenes.quat_to_matrix([1, 0, 0, 0], [3.0 * k, 0, 0])
        gw.insertPointcloud(T, s["points"])
        ow.insertPointcloud(T, s["points"])
        retries += gw.last_scan_stats()["retries"]
    assert gw.last_scan_stats()["blocks_total"] > 64 and retries >= 2
    assert_same_leaves(gw, ow)


def test_max_blocks_cap_is_an_error_not_corruption():
    s = scenes.cfg1_pointcloud()
    gw = vm.OctomapWorld(initial_blocks=64, max_blocks=64, **s["params"])
    with pytest.raises(vm.VmapError) as e:
        gw.insertPointcloud(s["T"], s["points"])
        gw.flush()                  # (pipelined insertion: the failure is reported when the scan is retired)
    assert e.value.status == 3      # VMAP_ERR_CAPACITY
    # the refused scan left nothing behind: the map is still empty and a scan that fits is exact
    assert gw.leaf_count() == 0
    ow = oracle.OracleWorld(**s["params"])
    small = s["points"][:40] * np.float32(0.2)
    gw.insertPointcloud(s["T"], small)
    ow.insertPointcloud(s["T"], small)
    assert_same_leaves(gw, ow)


# ---------------------------------------------------------------------------------------
# updateOccupancy on explicit keys, import / export, reset, parameters
# ---------------------------------------------------------------------------------------
def test_apply_keys_occupied_wins_and_duplicates():
    rng = np.random.default_rng(11)
    free = rng.integers(32700, 32840, (20000, 3)).astype(np.uint16)
    occ = rng.integers(32700, 32840, (3000, 3)).astype(np.uint16)
    occ[:500] = free[:500]          # overlap: occupied wins
    gw, ow = pair(resolution=0.1)
    for _ in range(3):
        gw.updateOccupancy(free, occ)
        ow.updateOccupancy(free, occ)
        assert_same_keys(gw, ow)
    assert_same_leaves(gw, ow)


def test_export_import_round_trip_and_reset():
    s = scenes.cfg1_pointcloud()
    gw = vm.OctomapWorld(**s["params"])
    gw.insertPointcloud(s["T"], s["points"])
    k, v = gw.export_leaves()
    gw2 = vm.OctomapWorld(initial_blocks=16, **s["params"])
    gw2.import_leaves(k, v)
    k2, v2 = leaves_packed(gw2)
    k1, v1 = leaves_packed(gw)
    assert np.array_equal(k1, k2) and np.array_equal(v1, v2)
    # continue inserting into the imported copy: stays identical to the original
    T = scenes.quat_to_matrix([1, 0, 0, 0], [0.4, 0.1, 0.0])
    gw.insertPointcloud(T, s["points"])
    gw2.insertPointcloud(T, s["points"])
    a, b = leaves_packed(gw), leaves_packed(gw2)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    gw.resetMap()
    assert gw.leaf_count() == 0
    assert gw.getCellStatusPoint([[0.1, 0.1, 0.1]]).tolist() == [1]   # unknown -> occupied


def test_set_params_resolution_change_resets_map():
    gw, ow = pair(resolution=0.1)
    pts = np.array([[0.55, 0.05, 0.05]], np.float32)
    gw.insertPointcloud(I4, pts)
    p = gw.getOctomapParameters()
    p.probability_hit = 0.8
    gw.setOctomapParameters(p)
    assert gw.leaf_count() == 6                 # non-resolution change keeps the map
    p.resolution = 0.2
    gw.setOctomapParameters(p)
    assert gw.leaf_count() == 0                 # octomap_world.cc:85-89
    op = ow.getOctomapParameters()
    op.probability_hit = 0.8
    op.resolution = 0.2
    ow.setOctomapParameters(op)
    gw.insertPointcloud(I4, pts)
    ow.insertPointcloud(I4, pts)
    assert_same_leaves(gw, ow)


# ---------------------------------------------------------------------------------------
# K5: queries
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("unk", [1, 0])
def test_queries_vs_oracle(unk):
    rng = np.random.default_rng(8)
    kw = dict(resolution=0.2, sensor_max_range=6.0, treat_unknown_as_occupied=unk)
    gw, ow = pair(**kw)
    for _ in range(3):
        d = rng.normal(size=(4000, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        pts = (d * rng.uniform(2.0, 5.0, (4000, 1))).astype(np.float32)
        gw.insertPointcloud(I4, pts)
        ow.insertPointcloud(I4, pts)
    qp = rng.uniform(-6, 6, (50000, 3))
    qp[:10] = [[1e9, 0, 0]] * 10                       # out of bounds -> NULL node
    assert np.array_equal(gw.getCellStatusPoint(qp), ow.getCellStatusPoint(qp))
    seg = np.concatenate([rng.uniform(-2, 2, (50000, 3)), rng.uniform(-6, 6, (50000, 3))], axis=1)
    seg[:5, 3:] = seg[:5, :3]                          # same-voxel segments -> free
    seg[5:10, 3] = 1e7                                 # out of bounds end -> free (ray empty)
    g, o = gw.getLineStatus(seg), ow.getLineStatus(seg)
    assert np.array_equal(g, o)
    assert len(set(g.tolist())) >= 2
    assert gw.getLineStatus(np.zeros((0, 6))).shape == (0,)


def test_cfg5_style_queries_on_hdl64_map():
    gw = vm.OctomapWorld(resolution=0.1, sensor_max_range=50.0)
    ow = oracle.OracleWorld(resolution=0.1, sensor_max_range=50.0)
    ow.set_faithful_inner_update(False)
    for i in range(2):
        s = scenes.trajectory_scan(10 * i, resolution=0.1)
        gw.insertPointcloudFast(s["T"], s["points"])
        ow.insertPointcloud(s["T"], s["points"])
    q = scenes.cfg5_queries(n_segments=200000, n_points=200000)
    assert np.array_equal(gw.getCellStatusPoint(q["points"]), ow.getCellStatusPoint(q["points"]))
    g, o = gw.getLineStatus(q["segments"]), ow.getLineStatus(q["segments"])
    assert np.array_equal(g, o)
    assert set(g.tolist()) == {0, 1}        # treat_unknown_as_occupied folds unknown into 1


def test_bench_query_workloads_vs_oracle_subsample():
    """The two segment workloads bench.py times (uniform in the mapped box; "hard": starting in known-free space
    along the sensor path), 100k segments each, on a map of trajectory scans -- statuses against the oracle, and
    the probe counter behind the query roofline against the oracle's key counts."""
    import torch
    n_scans = 3
    gw = vm.OctomapWorld(resolution=0.1, sensor_max_range=50.0)
    ow = oracle.OracleWorld(resolution=0.1, sensor_max_range=50.0)
    ow.set_faithful_inner_update(False)
    for i in range(n_scans):
        s = scenes.trajectory_scan(i, resolution=0.1)
        gw.insertPointcloudFast(s["T"], s["points"])
        ow.insertPointcloud(s["T"], s["points"])
    m = 100000
    q = scenes.cfg5_queries(n_segments=m, n_points=m, bounds=((-45.0, 45.0 + 0.5 * n_scans), (-45.0, 45.0), (-0.5, 4.0)))
    rng = np.random.default_rng(7)
    a = np.stack([rng.uniform(0.0, 0.5 * n_scans, m), rng.uniform(-8.0, 8.0, m), rng.uniform(0.3, 1.7, m)], axis=1)
    d = rng.normal(size=(m, 3))
    d[:, 2] *= 0.15
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    hard = np.concatenate([a, a + d * rng.uniform(2.0, 30.0, (m, 1))], axis=1)
    for name, segs in (("uniform", q["segments"]), ("hard", hard)):
        g, o = gw.getLineStatus(segs), ow.getLineStatus(segs)
        assert np.array_equal(g, o), name
        seg = torch.from_numpy(np.ascontiguousarray(segs)).cuda()
        st = torch.zeros(m, dtype=torch.uint8, device="cuda")
        probes = gw.getLineStatusProbesDevice(seg.data_ptr(), m, st.data_ptr())
        assert np.array_equal(st.cpu().numpy(), o), name
        assert probes >= int((o == 0).sum())          # a free segment probes at least its one voxel
    # the hard subset really walks: a good share of its segments cross free space and end free
    assert (ow.getLineStatus(hard) == 0).mean() > 0.2
    assert np.array_equal(gw.getCellStatusPoint(q["points"]), ow.getCellStatusPoint(q["points"]))


def test_walker_variants_build_the_same_map(monkeypatch):
    """Every segment-walker variant (block addressing, dilated 32- / 64-voxel words, queued marks) must leave the
    same map bit for bit; the default (queued) one is also what every other test in this file runs."""
    scans = [scenes.trajectory_scan(5 * i, resolution=0.2) for i in range(2)]
    ref = None
    for addr in ("0", "1", "7", "9", "8"):
        monkeypatch.setenv("VMAP_WALK_ADDR", addr)
        w = vm.OctomapWorld(resolution=0.2, sensor_max_range=50.0)
        stats = []
        for s in scans:
            w.insertPointcloud(s["T"], s["points"])
            st = w.last_scan_stats()
            stats.append((st["traversals"], st["unique_free"], st["unique_occupied"], st["longest_ray"]))
        dg = (w.leaf_digest(), stats)
        w.close()
        if ref is None:
            ref = dg
        assert dg == ref, addr
    monkeypatch.delenv("VMAP_WALK_ADDR")


# ---------------------------------------------------------------------------------------
# BASELINE.json full sizes: cfg4 against the oracle (one scan), and size-independent
# properties where the oracle would take too long
# ---------------------------------------------------------------------------------------
def test_cfg4_scan_full_size_0p05m():
    s = scenes.cfg4_scan(37)                       # 0.05 m voxels, ~56 M traversals
    gw, ow = pair("dense", **s["params"])
    ow.set_faithful_inner_update(False)
    gw.insertPointcloud(s["T"], s["points"])
    ow.insertPointcloud(s["T"], s["points"])
    assert gw.last_scan_stats()["traversals"] > 50_000_000
    assert_same_stats(gw, ow)
    assert_same_keys(gw, ow)
    assert_same_leaves(gw, ow)


def test_full_size_properties_idempotence_at_clamps_and_round_trip():
    s = scenes.cfg3_scan()
    gw = vm.OctomapWorld(**s["params"])
    gw.set_debug_capture(True)
    for _ in range(7):                             # 5 misses reach the lower clamp, 6 hits the upper
        gw.insertPointcloudFast(s["T"], s["points"])
    f, o = gw.last_scan_keys()
    pf, po = oracle.pack_keys(f), oracle.pack_keys(o)
    assert np.intersect1d(pf, po).size == 0        # free and occupied sets are disjoint
    assert np.unique(pf).size == pf.size and np.unique(po).size == po.size
    k7, v7 = leaves_packed(gw)
    assert set(np.unique(v7).tolist()) == {0xbfff07f4, 0x405e7867}   # every leaf sits at a clamp
    gw.insertPointcloudFast(s["T"], s["points"])    # ... so one more identical scan changes nothing
    k8, v8 = leaves_packed(gw)
    assert np.array_equal(k7, k8) and np.array_equal(v7, v8)
    assert k7.size == pf.size + po.size             # the map is exactly the scan's key sets
    # export -> import -> export is the identity at full size (9.8 M leaves)
    k, v = gw.export_leaves()
    g2 = vm.OctomapWorld(**s["params"])
    g2.import_leaves(k, v)
    k2, v2 = leaves_packed(g2)
    assert np.array_equal(k2, k8) and np.array_equal(v2, v8)


def test_checkpoint_and_list_growth_fine_resolution():
    # 20k rays of up to ~1700 DDA steps at 1 cm: far more segment checkpoints and touched blocks
    # than the initial buffers hold -> exercises the grow-and-rerun paths of the segment walker
    rng = np.random.default_rng(17)
    kw = dict(resolution=0.01, sensor_max_range=10.0)
    gw, ow = pair("dense", **kw)
    ow.set_faithful_inner_update(False)
    d = rng.normal(size=(20000, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    pts = (d * rng.uniform(1.0, 12.0, (20000, 1))).astype(np.float32)
    T = scenes.quat_to_matrix(scenes.random_unit_quat(rng), [0.3, -0.2, 0.1])
    gw.insertPointcloud(T, pts)
    ow.insertPointcloud(T, pts)
    assert gw.last_scan_stats()["retries"] >= 1
    assert_same_stats(gw, ow)
    assert_same_keys(gw, ow)
    assert_same_leaves(gw, ow)
