"""Pipelined scan insertion (include/vmap/vmap.h, "the scan pipeline"): up to two scans in flight when a
call returns, on three scan contexts (two with VMAP_PIPE_CTX=2), map updates ordered by events, refusals on
the device replayed by the host.
Bar: whatever the depth and however often the map has to grow under the pipeline, the leaves equal
the CPU oracle's bit for bit (octomap_world.cc:110-141, 238-264 called scan after scan)."""
import numpy as np
import pytest

import oracle
import volumetric_mapping_b200 as vm
from volumetric_mapping_b200 import scenes, sharded

pytestmark = pytest.mark.gpu


def leaves(w):
    k, v = w.export_leaves()
    p = oracle.pack_keys(k)
    o = np.argsort(p)
    return p[o], v.view(np.uint32)[o]


def hdl_scans(n, res):
    return [scenes.trajectory_scan(3 * i, resolution=res) for i in range(n)]


@pytest.mark.parametrize("depth,contexts", [(0, 3), (1, 3), (2, 3), (2, 2)])
def test_pipelined_build_equals_oracle_with_growth_under_the_pipeline(depth, contexts, monkeypatch):
    """initial_blocks = 64: the first scans do not fit, their updates are refused on the device
    (table, then pool), the scans behind them refuse themselves, and the host replays them all in order."""
    params = dict(resolution=0.2, sensor_max_range=50.0)
    scans = hdl_scans(7, 0.2)
    monkeypatch.setenv("VMAP_PIPE_CTX", str(contexts))
    gw = vm.OctomapWorld(initial_blocks=64, **params)
    gw.set_pipeline_depth(depth)
    ow = oracle.OracleWorld(**params)
    ow.set_faithful_inner_update(False)
    for s in scans:
        gw.insertPointcloudFast(s["T"], s["points"])      # no reading call in between: scans overlap
        ow.insertPointcloud(s["T"], s["points"])
    t = gw.totals()
    assert t["scans"] == len(scans)
    if depth:
        assert t["replayed"] >= 1
    gk, gv = leaves(gw)
    ok, ov = leaves(ow)
    assert np.array_equal(gk, ok) and np.array_equal(gv, ov)
    n, d = gw.leaf_digest()
    assert (n, d) == sharded.leaf_digest(*gw.export_leaves())


def test_totals_are_the_sum_of_the_per_scan_counters():
    params = dict(resolution=0.2, sensor_max_range=50.0)
    scans = hdl_scans(4, 0.2)
    a = vm.OctomapWorld(**params)
    a.set_pipeline_depth(0)
    per = []
    for s in scans:
        a.insertPointcloudFast(s["T"], s["points"])
        per.append(a.last_scan_stats())
    b = vm.OctomapWorld(**params)
    for s in scans:
        b.insertPointcloudFast(s["T"], s["points"])
    tb = b.totals()
    for key in ("points_in", "points_valid", "points_in_range", "rays_cast", "traversals", "occupied_emitted",
                "unique_free", "unique_occupied", "blocks_touched"):
        assert tb[key] == sum(p[key] for p in per), key
    assert tb["longest_ray"] == max(p["longest_ray"] for p in per)
    assert a.leaf_digest() == b.leaf_digest()
    # the stats of the scan inserted last are available after the pipeline has drained
    last = b.last_scan_stats()
    for key in ("traversals", "unique_free", "unique_occupied", "rays_cast"):
        assert last[key] == per[-1][key], key


@pytest.mark.parametrize("stride,is_dense", [(12, True), (16, True), (32, True), (16, False)])
def test_pipelined_write_back_of_the_world_frame_cloud(stride, is_dense):
    """insertPointcloudIntoMapImpl rewrites the caller's cloud (octomap_world.cc:115-119); with a
    padded point type only x, y, z may change, and with is_dense == false NaN points are dropped."""
    import ctypes as C
    rng = np.random.default_rng(5)
    params = dict(resolution=0.15, sensor_max_range=5.0)
    gw, ow = vm.OctomapWorld(**params), oracle.OracleWorld(**params)
    L = gw._L
    for k in range(3):
        n = 5000
        d = rng.normal(size=(n, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        pts = (d * rng.uniform(0.5, 7.0, (n, 1))).astype(np.float32)
        if not is_dense:
            pts[rng.integers(0, n, 200)] = np.nan
        T = scenes.quat_to_matrix(scenes.random_unit_quat(rng), rng.uniform(-1, 1, 3))
        w4 = stride // 4
        buf = rng.normal(size=(n, w4)).astype(np.float32)      # payload in the extra floats
        buf[:, :3] = pts
        before = buf.copy()
        Tm = np.ascontiguousarray(T, dtype=np.float64).reshape(16)
        n_out = C.c_size_t(0)
        st = L.vmap_insert_pointcloud(gw.handle, buf.ctypes.data, n, stride, int(is_dense),
                                      Tm.ctypes.data_as(C.POINTER(C.c_double)), buf.ctypes.data, C.byref(n_out))
        assert st == 0
        want = ow.insertPointcloud(T, pts, is_dense=is_dense)
        assert n_out.value == want.shape[0]
        assert np.array_equal(buf[: n_out.value, :3].view(np.uint32), want.view(np.uint32))
        if is_dense and w4 > 3:
            assert np.array_equal(buf[:, 3:].view(np.uint32), before[:, 3:].view(np.uint32))
    gk, gv = leaves(gw)
    ok, ov = leaves(ow)
    assert np.array_equal(gk, ok) and np.array_equal(gv, ov)


def test_deferred_capacity_error_drops_the_scan_and_nothing_else():
    s = scenes.cfg1_pointcloud()
    gw = vm.OctomapWorld(initial_blocks=64, max_blocks=64, **s["params"])
    small = s["points"][:40] * np.float32(0.2)
    ow = oracle.OracleWorld(**s["params"])
    gw.insertPointcloudFast(s["T"], small)
    ow.insertPointcloud(s["T"], small)
    gw.insertPointcloudFast(s["T"], s["points"])           # does not fit: refused, reported later
    gw.insertPointcloudFast(s["T"], small * np.float32(1.5))   # queued behind it: must still be applied
    ow.insertPointcloud(s["T"], small * np.float32(1.5))
    with pytest.raises(vm.VmapError) as e:
        gw.flush()
    assert e.value.status == 3
    gk, gv = leaves(gw)
    ok, ov = leaves(ow)
    assert np.array_equal(gk, ok) and np.array_equal(gv, ov)


def test_region_timer_and_mixed_calls_drain_the_pipeline():
    params = dict(resolution=0.2, sensor_max_range=50.0)
    scans = hdl_scans(5, 0.2)
    gw, ow = vm.OctomapWorld(**params), oracle.OracleWorld(**params)
    ow.set_faithful_inner_update(False)
    q = scenes.cfg5_queries(n_segments=20000, n_points=20000)
    gw.region_begin()
    for i, s in enumerate(scans):
        gw.insertPointcloudFast(s["T"], s["points"])
        ow.insertPointcloud(s["T"], s["points"])
        if i == 2:     # a reading call in the middle sees exactly the scans inserted so far
            assert np.array_equal(gw.getLineStatus(q["segments"]), ow.getLineStatus(q["segments"]))
    ms = gw.region_end()
    assert ms > 0.0
    assert np.array_equal(gw.getCellStatusPoint(q["points"]), ow.getCellStatusPoint(q["points"]))
    gk, gv = leaves(gw)
    ok, ov = leaves(ow)
    assert np.array_equal(gk, ok) and np.array_equal(gv, ov)
