"""Property tests (hypothesis) of the oracle's scan insertion -- SURVEY section 4(d): the free and
occupied key sets of a scan are disjoint, every free key lies on some cast ray, clamped leaves are
fixed points of further identical scans, and the closed form of the skip rule
(cast(i) <=> i <= first[K_i]) agrees with the sequential loop."""
import numpy as np
from hypothesis import given, settings, strategies as st

import oracle
import pyref

I4 = np.eye(4)


def cloud(seed, n, rmax):
    rng = np.random.default_rng(seed)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    pts = (d * rng.uniform(0.1, rmax, (n, 1))).astype(np.float32)
    m = min(len(pts[::7]), len(pts[1::7]))
    pts[::7][:m] = pts[1::7][:m] * np.float32(1.0005)               # shared end voxels
    return pts


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 10_000), n=st.integers(1, 300),
       res=st.sampled_from([0.05, 0.1, 0.15, 0.3]), max_range=st.sampled_from([-1.0, 1.5, 4.0]))
def test_key_sets_are_disjoint_and_lie_on_cast_rays(seed, n, res, max_range):
    pts = cloud(seed, n, 5.0)
    w = oracle.OracleWorld(resolution=res, sensor_max_range=max_range)
    w.insertPointcloud(I4, pts)
    f, o = w.last_scan_keys()
    fs, os_ = set(map(tuple, f.tolist())), set(map(tuple, o.tolist()))
    assert not (fs & os_)
    on_rays = set()
    for p in pts:                                   # union over ALL rays is a superset of the cast ones
        r = float(np.linalg.norm(p.astype(np.float64)))
        end = p
        if max_range >= 0 and np.float32(np.sqrt(np.float32(p @ p))) > max_range:
            end = (p / np.float32(np.sqrt(np.float32(p @ p)))) * np.float32(max_range)
        keys = w.computeRayKeys((0.0, 0.0, 0.0), end)
        if keys is not None:
            on_rays |= set(map(tuple, keys.tolist()))
    assert fs <= on_rays
    st_ = w.last_scan_stats()
    assert st_["unique_free"] == len(fs) and st_["unique_occupied"] == len(os_)
    assert st_["rays_cast"] <= n and st_["traversals"] >= len(fs)


@settings(max_examples=15, deadline=None)
@given(seed=st.integers(0, 10_000), n=st.integers(1, 200))
def test_skip_rule_closed_form(seed, n):
    res, max_range = 0.2, 3.0
    pts = cloud(seed, n, 4.5)
    w = oracle.OracleWorld(resolution=res, sensor_max_range=max_range)
    w.insertPointcloud(I4, pts)
    # closed form: first[K] = min index with (in range and in bounds and key K); cast(i) <=> i <= first[K_i]
    keys = [tuple(w.coordToKey(float(c)) for c in p) for p in pts]
    in_range = [float(np.sqrt(np.float64(np.float32(np.float32(p[0] * p[0]) + np.float32(p[1] * p[1]) + np.float32(p[2] * p[2]))))) <= max_range
                for p in pts]
    first = {}
    for i, (k, r) in enumerate(zip(keys, in_range)):
        if r and k not in first:
            first[k] = i
    cast = sum(1 for i, k in enumerate(keys) if i <= first.get(k, 1 << 60))
    assert cast == w.last_scan_stats()["rays_cast"]


@settings(max_examples=8, deadline=None)
@given(seed=st.integers(0, 10_000))
def test_clamped_map_is_a_fixed_point(seed):
    pts = cloud(seed, 150, 3.0)
    w = oracle.OracleWorld(resolution=0.2, sensor_max_range=4.0)
    for _ in range(7):
        w.insertPointcloud(I4, pts)
    k1, v1 = w.export_leaves()
    w.insertPointcloud(I4, pts)
    k2, v2 = w.export_leaves()
    a, b = np.argsort(oracle.pack_keys(k1)), np.argsort(oracle.pack_keys(k2))
    assert np.array_equal(oracle.pack_keys(k1)[a], oracle.pack_keys(k2)[b])
    assert np.array_equal(v1.view(np.uint32)[a], v2.view(np.uint32)[b])
    assert set(np.unique(v1.view(np.uint32)).tolist()) <= {0xbfff07f4, 0x405e7867}


def test_python_restatement_agrees_on_hypothesis_style_cloud():
    pts = cloud(123, 120, 4.0)
    w = oracle.OracleWorld(resolution=0.15, sensor_max_range=3.0)
    pw = pyref.PyWorld(resolution=0.15, sensor_max_range=3.0)
    w.insertPointcloud(I4, pts)
    pw.insert_pointcloud(I4, pts)
    f, o = w.last_scan_keys()
    assert set(map(tuple, f.tolist())) == pw.last_free and set(map(tuple, o.tolist())) == pw.last_occ
