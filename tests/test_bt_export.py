"""Export to octomap's tree form (north_star: "Export to octomap::OcTree, with pruning and inner-
node max, on demand on the host"; octomap_world.cc:846-856).  octomap is not available here, so
the ".bt" layout is restated from upstream [UPSTREAM-recalled]; it is pinned by a hand-derived
stream (CPU, oracle) and by two independent writers agreeing byte for byte (GPU: bottom-up fold of
Morton-sorted leaves; oracle: recursive walk of the pointer octree)."""
import os
import tempfile

import numpy as np
import pytest

import oracle
from volumetric_mapping_b200 import scenes

HEADER = (b"# Octomap OcTree binary file\n# (feel free to add / change comments, but leave the first "
          b"line as it is!)\n#\nid OcTree\nsize %d\nres %s\ndata\n")


def test_oracle_bt_hand_derived_stream():
    # occupied voxel (32768,32768,32768) hit twice, free neighbour (32769,32768,32768):
    # root -> child 7 -> 14 x child 0 -> {child 0 occupied, child 1 free}; 18 nodes
    w = oracle.OracleWorld(resolution=0.1)
    occ = np.array([[32768, 32768, 32768]], np.uint16)
    free = np.array([[32769, 32768, 32768]], np.uint16)
    w.updateOccupancy(free, occ)
    w.updateOccupancy(np.zeros((0, 3), np.uint16), occ)
    assert w.tree_node_count() == 18
    data = bytes([0x00, 0xC0]) + bytes([0x03, 0x00]) * 14 + bytes([0x06, 0x00])
    assert w.write_bt(live=False) == HEADER % (18, b"0.1") + data


def test_oracle_prune_after_max_likelihood_collapses_uniform_octants():
    # 8 sibling voxels, 7 hit once (0.619, below the 0.847 threshold) and 1 missed once: different
    # values, so no pruning -- until toMaxLikelihood maps all eight to the clamping minimum
    w = oracle.OracleWorld(resolution=0.1)
    sib = np.array([[32768 + (i & 1), 32768 + ((i >> 1) & 1), 32768 + (i >> 2)] for i in range(8)], np.uint16)
    w.updateOccupancy(sib[7:], sib[:7])
    assert w.tree_node_count() == 16 + 8
    live = w.write_bt(live=True)
    assert w.tree_node_count() == 16            # the eight leaves collapsed into their parent
    data = bytes([0x00, 0xC0]) + bytes([0x03, 0x00]) * 13 + bytes([0x01, 0x00])   # free leaf at depth 15
    assert live == HEADER % (16, b"0.1") + data


def test_oracle_ot_hand_derived_stream():
    # same 18-node tree as above: every node = float value + child byte, depth first
    import struct
    w = oracle.OracleWorld(resolution=0.1)
    occ = np.array([[32768, 32768, 32768]], np.uint16)
    free = np.array([[32769, 32768, 32768]], np.uint16)
    w.updateOccupancy(free, occ)
    w.updateOccupancy(np.zeros((0, 3), np.uint16), occ)
    hit2 = struct.pack("<I", 0x3f9e795b)           # two hits (SURVEY 3.3 ladder)
    miss = struct.pack("<I", 0xbecf991f)
    inner = hit2                                   # inner node = max of its children
    data = inner + bytes([0x80]) + (inner + bytes([0x01])) * 14 + inner + bytes([0x03]) + hit2 + b"\0" + miss + b"\0"
    assert w.write_ot(with_header=False) == data
    hdr = (b"# Octomap OcTree file\n# (feel free to add / change comments, but leave the first line as "
           b"it is!)\n#\nid OcTree\nsize 18\nres 0.1\ndata\n")
    assert w.write_ot(with_header=True) == hdr + data


@pytest.mark.gpu
def test_gpu_ot_stream_equals_oracle_stream_and_round_trips():
    import volumetric_mapping_b200 as vm
    rng = np.random.default_rng(13)
    gw = vm.OctomapWorld(resolution=0.1, sensor_max_range=6.0)
    ow = oracle.OracleWorld(resolution=0.1, sensor_max_range=6.0)
    for k in range(3):
        d = rng.normal(size=(5000, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        pts = (d * rng.uniform(1.0, 7.0, (5000, 1))).astype(np.float32)
        T = scenes.quat_to_matrix([1, 0, 0, 0], [0.2 * k, 0, 0])
        gw.insertPointcloud(T, pts)
        ow.insertPointcloud(T, pts)
    for hdr in (False, True):
        assert gw.getOctomapFullMsgData(with_header=hdr) == ow.write_ot(with_header=hdr)
    g2 = vm.OctomapWorld(resolution=0.25)
    g2.setOctomapFromFullMsgData(gw.getOctomapFullMsgData(with_header=True), with_header=True)
    k1, v1 = gw.export_leaves()
    k2, v2 = g2.export_leaves()
    a, b = np.argsort(oracle.pack_keys(k1)), np.argsort(oracle.pack_keys(k2))
    assert np.array_equal(oracle.pack_keys(k1)[a], oracle.pack_keys(k2)[b])
    assert np.array_equal(v1.view(np.uint32)[a], v2.view(np.uint32)[b])


@pytest.mark.gpu
@pytest.mark.parametrize("res,n_scans", [(0.15, 1), (0.1, 3)])
def test_gpu_bt_stream_equals_oracle_stream(res, n_scans):
    import volumetric_mapping_b200 as vm
    rng = np.random.default_rng(12)
    gw = vm.OctomapWorld(resolution=res, sensor_max_range=6.0)
    ow = oracle.OracleWorld(resolution=res, sensor_max_range=6.0)
    for k in range(n_scans):
        d = rng.normal(size=(6000, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        pts = (d * rng.uniform(1.0, 7.0, (6000, 1))).astype(np.float32)
        T = scenes.quat_to_matrix([1, 0, 0, 0], [0.2 * k, 0, 0])
        gw.insertPointcloud(T, pts)
        ow.insertPointcloud(T, pts)
    # node count of the (non-thresholded) pruned tree
    ow.prune()
    assert gw.tree_size(max_likelihood=False) == ow.tree_node_count()
    # const writer: the oracle writes its tree as it is (log-odds values, pruned by equality); the
    # GPU side thresholds a copy first, so compare after the live write instead
    g_const = gw.writeOctomapToBinaryConst()
    o_live = ow.write_bt(live=True)
    assert g_const == o_live
    assert gw.tree_size(max_likelihood=True) == ow.tree_node_count()
    # writeOctomapToFile thresholds the LIVE map on both sides
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "map.bt")
        gw.writeOctomapToFile(path)
        assert open(path, "rb").read() == o_live
        gk, gv = gw.export_leaves()
        ok, ov = ow.export_leaves()
        go, oo = np.argsort(oracle.pack_keys(gk)), np.argsort(oracle.pack_keys(ok))
        assert np.array_equal(oracle.pack_keys(gk)[go], oracle.pack_keys(ok)[oo])
        assert np.array_equal(gv.view(np.uint32)[go], ov.view(np.uint32)[oo])
        # loadOctomapFromFile into a fresh world at another resolution: map and resolution replaced
        g2 = vm.OctomapWorld(resolution=0.3)
        g2.loadOctomapFromFile(path)
        assert g2.getOctomapParameters().resolution == pytest.approx(res)
        k2, v2 = g2.export_leaves()
        o2 = np.argsort(oracle.pack_keys(k2))
        assert np.array_equal(oracle.pack_keys(k2)[o2], oracle.pack_keys(gk)[go])
        assert np.array_equal(v2.view(np.uint32)[o2], gv.view(np.uint32)[go])
        assert g2.writeOctomapToBinaryConst() == o_live
        g3 = vm.OctomapWorld(resolution=res)
        g3.setOctomapFromBinaryMsg(o_live)
        assert g3.writeOctomapToBinaryConst() == o_live
