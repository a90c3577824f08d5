"""The product's host-side tree fold and stream writers / readers (vmap_export_host.inl), run on
the CPU through the pure-host entry points vmap_host_tree_stream / vmap_host_parse_stream: maps are
built by the oracle, and the product's bytes must equal the oracle's recursive pointer-tree
writers'.  (The oracle is only the checker here: it provides the leaves and the expected bytes.)"""
import ctypes as C

import numpy as np
import pytest

import oracle
from volumetric_mapping_b200 import _capi, scenes

I4 = np.eye(4)


def product_stream(params, k, v, kind):
    L = _capi.lib()
    k = np.ascontiguousarray(k, np.uint16); v = np.ascontiguousarray(v, np.float32)
    n, nodes = C.c_size_t(0), C.c_size_t(0)
    assert L.vmap_host_tree_stream(C.byref(params), k.ctypes.data, v.ctypes.data, len(v), kind, None, 0,
                                   C.byref(n), C.byref(nodes)) == 0
    buf = (C.c_uint8 * max(1, n.value))()
    assert L.vmap_host_tree_stream(C.byref(params), k.ctypes.data, v.ctypes.data, len(v), kind, buf, n.value,
                                   C.byref(n), C.byref(nodes)) == 0
    return bytes(buf[: n.value]), nodes.value


def product_parse(params, data, kind):
    L = _capi.lib()
    buf = (C.c_uint8 * max(1, len(data))).from_buffer_copy(data or b"\0")
    n, res = C.c_size_t(0), C.c_double(0)
    st = L.vmap_host_parse_stream(C.byref(params), kind, buf, len(data), None, None, 0, C.byref(n), C.byref(res))
    assert st in (0, 5)
    k = np.zeros((n.value, 3), np.uint16); v = np.zeros(n.value, np.float32)
    if n.value:
        assert L.vmap_host_parse_stream(C.byref(params), kind, buf, len(data), k.ctypes.data, v.ctypes.data,
                                        n.value, C.byref(n), C.byref(res)) == 0
    return k, v, res.value


def sorted_leaves(k, v):
    p = oracle.pack_keys(k)
    o = np.argsort(p)
    return p[o], np.asarray(v, np.float32).view(np.uint32)[o]


@pytest.mark.parametrize("res,seed,n_scans", [(0.15, 1, 1), (0.1, 2, 4), (0.3, 3, 8)])
def test_product_writers_match_oracle_writers(res, seed, n_scans):
    rng = np.random.default_rng(seed)
    ow = oracle.OracleWorld(resolution=res, sensor_max_range=5.0)
    for k in range(n_scans):
        d = rng.normal(size=(1500, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        pts = (d * rng.uniform(0.5, 6.0, (1500, 1))).astype(np.float32)
        ow.insertPointcloud(scenes.quat_to_matrix([1, 0, 0, 0], [0.1 * k, 0, 0]), pts)
    params = _capi.default_params(resolution=res, sensor_max_range=5.0)
    k, v = ow.export_leaves()
    # full tree (inner = max of children, equal-leaf pruning): node count and both .ot forms
    ot_data, nodes = product_stream(params, k, v, 1)
    ow.prune()
    assert nodes == ow.tree_node_count()
    assert ot_data == ow.write_ot(with_header=False)
    assert product_stream(params, k, v, 2)[0] == ow.write_ot(with_header=True)
    # readers: the product parses its own and the oracle's streams back to the same leaves
    k2, v2, r2 = product_parse(params, ow.write_ot(with_header=True), 2)
    assert r2 == pytest.approx(res)
    a, b = sorted_leaves(k, v), sorted_leaves(k2, v2)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    # .bt: max-likelihood + prune (the oracle thresholds its live tree, like OcTree::writeBinary)
    bt, bt_nodes = product_stream(params, k, v, 0)
    o_bt = ow.write_bt(live=True)
    assert bt == o_bt and bt_nodes == ow.tree_node_count()
    k3, v3, _ = product_parse(params, o_bt, 0)
    ko, vo = ow.export_leaves()                      # thresholded leaves
    a, b = sorted_leaves(ko, vo), sorted_leaves(k3, v3)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_empty_and_malformed_streams():
    params = _capi.default_params(resolution=0.1)
    data, nodes = product_stream(params, np.zeros((0, 3), np.uint16), np.zeros(0, np.float32), 0)
    assert nodes == 0 and data.endswith(b"size 0\nres 0.1\ndata\n")
    k, v, _ = product_parse(params, data, 0)
    assert len(v) == 0
    L = _capi.lib()
    bad = b"not an octomap stream"
    buf = (C.c_uint8 * len(bad)).from_buffer_copy(bad)
    n, res = C.c_size_t(0), C.c_double(0)
    assert L.vmap_host_parse_stream(C.byref(params), 0, buf, len(bad), None, None, 0, C.byref(n), C.byref(res)) == 1
    # truncated data section
    ow = oracle.OracleWorld(resolution=0.1)
    ow.insertPointcloud(I4, [[0.55, 0.05, 0.05], [0.05, 0.75, 0.05]])
    good = ow.write_bt(live=False)
    cut = good[:-3]
    buf = (C.c_uint8 * len(cut)).from_buffer_copy(cut)
    assert L.vmap_host_parse_stream(C.byref(params), 0, buf, len(cut), None, None, 0, C.byref(n), C.byref(res)) == 1


def test_cpp_exceptions_do_not_cross_the_c_abi():
    """An impossible allocation inside an entry point comes back as an error code and a message."""
    L = _capi.lib()
    params = _capi.default_params(resolution=0.1)
    k = np.zeros((1, 3), np.uint16); v = np.zeros(1, np.float32)
    n = C.c_size_t(0)
    st = L.vmap_host_tree_stream(C.byref(params), k.ctypes.data, v.ctypes.data, C.c_size_t(1 << 60), 0, None, 0,
                                 C.byref(n), None)
    assert st in (1, 3)
    L.vmap_last_error.restype = C.c_char_p
    msg = L.vmap_last_error(None)
    assert msg and (b"allocation" in msg or b"exception" in msg)
