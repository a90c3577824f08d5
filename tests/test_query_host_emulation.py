"""getCellStatusBoundingBox / checkPathForCollisionsWithRobot: the product's per-lane
query body (vmap_query.cuh), built for the host and run on maps in the product's own block-hash
layout, against the oracle's restatement on its pruned pointer tree (leaf_bbx_iterator,
getUnknownLeafCenters).  One emulated lane and 32 emulated lanes must both reproduce the oracle."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle
from volumetric_mapping_b200 import scenes

HERE = os.path.dirname(os.path.abspath(__file__))
NATIVE = os.path.join(HERE, "native")
CSRC = os.path.join(HERE, "..", "volumetric_mapping_b200", "csrc")
SRC = os.path.join(NATIVE, "query_check.cc")
LIB = os.path.join(NATIVE, "libquery_check.so")


@pytest.fixture(scope="module")
def lib():
    deps = [SRC, os.path.join(NATIVE, "cuda_host_shim.h")] + [os.path.join(CSRC, f) for f in
                                                               ("vmap_device.cuh", "vmap_query.cuh")]
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared",
                               "-DVMAP_HOST_EMULATION", "-include", os.path.join(NATIVE, "cuda_host_shim.h"),
                               "-x", "c++", SRC, "-o", LIB])
    L = C.CDLL(LIB)
    vp = C.c_void_p
    L.qc_map_create.argtypes = [vp, vp, C.c_size_t]
    L.qc_map_create.restype = vp
    L.qc_map_destroy.argtypes = [vp]
    L.qc_bbox_status.argtypes = [vp, C.c_double, C.c_float, C.c_int, vp, C.c_size_t, vp, C.c_uint32, vp]
    L.qc_point_status.argtypes = [vp, C.c_double, C.c_float, C.c_int, vp, C.c_size_t, C.c_int, vp, vp]
    L.qc_line_status.argtypes = [vp, C.c_double, C.c_float, C.c_int, vp, C.c_size_t, vp]
    L.qc_visibility.argtypes = [vp, C.c_double, C.c_float, C.c_int, vp, C.c_size_t, C.c_int, vp]
    return L


class HostMap:
    def __init__(self, L, ow):
        self.L, self.ow = L, ow
        k, v = ow.export_leaves()
        self.k = np.ascontiguousarray(k, np.uint16)
        self.v = np.ascontiguousarray(v, np.float32)
        self.h = L.qc_map_create(self.k.ctypes.data, self.v.ctypes.data, len(self.v))
        assert self.h
        p = ow.getOctomapParameters()
        self.res, self.unk = p.resolution, int(p.treat_unknown_as_occupied)
        self.occ = float(ow.logodds_constants()[4])

    def bbox(self, xyz, size, lanes):
        xyz = np.ascontiguousarray(xyz, np.float64).reshape(-1, 3)
        size = np.ascontiguousarray(size, np.float64)
        st = np.zeros(len(xyz), np.uint8)
        self.L.qc_bbox_status(self.h, self.res, self.occ, self.unk, xyz.ctypes.data, len(xyz), size.ctypes.data,
                              lanes, st.ctypes.data)
        return st

    def points(self, xyz, true_status=0):
        xyz = np.ascontiguousarray(xyz, np.float64).reshape(-1, 3)
        st = np.zeros(len(xyz), np.uint8)
        lo = np.zeros(len(xyz), np.float32)
        self.L.qc_point_status(self.h, self.res, self.occ, self.unk, xyz.ctypes.data, len(xyz), true_status,
                               st.ctypes.data, lo.ctypes.data)
        return st, lo

    def lines(self, ab):
        ab = np.ascontiguousarray(ab, np.float64).reshape(-1, 6)
        st = np.zeros(len(ab), np.uint8)
        self.L.qc_line_status(self.h, self.res, self.occ, self.unk, ab.ctypes.data, len(ab), st.ctypes.data)
        return st

    def visibility(self, ab, stop_at_unknown):
        ab = np.ascontiguousarray(ab, np.float64).reshape(-1, 6)
        st = np.zeros(len(ab), np.uint8)
        self.L.qc_visibility(self.h, self.res, self.occ, self.unk, ab.ctypes.data, len(ab), int(stop_at_unknown),
                             st.ctypes.data)
        return st

    def close(self):
        self.L.qc_map_destroy(self.h)


def build_world(res, unknown_as_occupied, n_scans=3, seed=0):
    p = oracle.default_params(resolution=res, sensor_max_range=12.0, treat_unknown_as_occupied=int(unknown_as_occupied))
    ow = oracle.OracleWorld(params=p)
    for i in range(n_scans):
        s = scenes.hdl64_scan(position=(0.4 * i, 0.1 * i, 1.2), yaw_rad=0.3 * i, box=(-6.0, 7.0, -5.0, 6.0),
                              seed=seed + i, n_elev=24, n_azim=240)
        for _ in range(2):          # two hits make a voxel occupied at the default thresholds
            ow.insertPointcloud(s["T"], s["points"])
    return ow


@pytest.mark.parametrize("res,unk", [(0.1, True), (0.1, False), (0.15, False), (0.25, True)])
def test_bbox_status_matches_oracle(lib, res, unk):
    ow = build_world(res, unk)
    hm = HostMap(lib, ow)
    rng = np.random.default_rng(7)
    pts = rng.uniform([-7, -6, -0.5], [8, 7, 3.0], size=(1500, 3))
    seen = set()
    for size in ([0.05, 0.05, 0.05], [0.3, 0.3, 0.3], [0.5, 0.4, 0.25], [1.0, 1.0, 1.0], [0.0, 0.6, 0.6], [2.2, 0.3, 0.9]):
        want = ow.getCellStatusBoundingBox(pts, size)
        for lanes in (1, 32):
            got = hm.bbox(pts, size, lanes)
            bad = np.flatnonzero(got != want)
            assert bad.size == 0, (size, lanes, bad[:5], pts[bad[:5]], got[bad[:5]], want[bad[:5]])
        seen.update(np.unique(want).tolist())
        # a box never reports "more free" than its centre
        centre = ow.getCellStatusPoint(pts)
        assert np.all((centre == 0) | (want == centre))
    assert seen >= ({0, 1} if unk else {0, 1, 2})
    hm.close()


def test_voxel_aligned_boxes_and_boundaries(lib):
    """Centres on voxel centres / corners, sizes that are multiples of the resolution: exercises the
    float accumulation of getUnknownLeafCenters and the <= / >= faces of the key range."""
    res = 0.1
    ow = build_world(res, False, n_scans=2, seed=5)
    hm = HostMap(lib, ow)
    g = np.arange(-20, 21)
    cx, cy = np.meshgrid(g, g)
    for off in (0.0, 0.05):
        pts = np.stack([cx.ravel() * 0.1 + off, cy.ravel() * 0.1 + off, np.full(cx.size, 1.05)], 1)
        for size in ([0.2, 0.2, 0.2], [0.3, 0.3, 0.1], [0.4, 0.4, 0.4]):
            want = ow.getCellStatusBoundingBox(pts, size)
            got = hm.bbox(pts, size, 32)
            bad = np.flatnonzero(got != want)
            assert bad.size == 0, (off, size, bad[:5], pts[bad[:5]], got[bad[:5]], want[bad[:5]])
    hm.close()


def test_pruned_leaf_below_the_box_is_visited_like_upstream(lib):
    """leaf_bbx_iterator's overlap test is conservative by one key on the lower side for leaves
    above the finest depth: an occupied PRUNED leaf that ends exactly where the box begins is
    visited, and passes the reference's own inside test when the faces coincide.  The product
    reconstructs the canonical pruned leaf of every occupied voxel, so it reports the same."""
    res = 0.25
    ow = oracle.OracleWorld(params=oracle.default_params(resolution=res, treat_unknown_as_occupied=0))
    B = 32768
    occ = np.array([[B + 2 + i, B + 2 + j, B + 2 + k] for i in range(2) for j in range(2) for k in range(2)], np.uint16)
    free = np.array([[B + i, B + j, B + k] for i in range(4, 8) for j in range(8) for k in range(8)], np.uint16)
    for _ in range(3):
        ow.updateOccupancy(free, occ)          # the eight equal occupied voxels collapse into one depth-15 leaf
    hm = HostMap(lib, ow)
    q = np.array([[1.25, 0.625, 0.625], [1.25, 0.875, 0.875], [1.125, 0.625, 0.625], [1.375, 0.625, 0.625]])
    for size, want in (([0.5, 0.5, 0.5], [1, 1, 1, 0]), ([0.25, 0.25, 0.25], [0, 0, 1, 0])):
        ref = ow.getCellStatusBoundingBox(q, size)
        assert ref.tolist() == want                       # the quirk is really there in the restated reference
        assert np.array_equal(hm.bbox(q, size, 32), ref)
        assert np.array_equal(hm.bbox(q, size, 1), ref)
    hm.close()


@pytest.mark.parametrize("res,lo", [(0.25, 16), (0.1, 64)])
def test_large_pruned_leaves_spanning_several_blocks(lib, res, lo):
    """A 16^3-voxel cube of identical occupied voxels collapses into one depth-12 leaf (two 8^3
    blocks per side): the leaf reconstruction has to look beyond the voxel's own block.  At 0.1 m
    the cube's upper x / y faces sit at 8.0 m -- exactly representable, so boxes that start there
    meet the leaf's face computed at the LEAF's depth, where the last ulp depends on the depth."""
    ow = oracle.OracleWorld(params=oracle.default_params(resolution=res, treat_unknown_as_occupied=0))
    B = 32768
    r16 = np.arange(16)
    occ = np.stack(np.meshgrid(B + lo + r16, B + lo + r16, B + r16, indexing="ij"), -1).reshape(-1, 3).astype(np.uint16)
    rr = np.arange(lo - 8, lo + 32)
    shell = np.stack(np.meshgrid(B + rr, B + rr, B + np.arange(-8, 24), indexing="ij"), -1).reshape(-1, 3)
    inside = np.all((shell >= [B + lo, B + lo, B]) & (shell < [B + lo + 16, B + lo + 16, B + 16]), axis=1)
    free = shell[~inside].astype(np.uint16)
    for _ in range(3):
        ow.updateOccupancy(free, occ)
    assert ow.tree_node_count() < 3 * len(free)          # the occupied cube is one node, not 4096
    hm = HostMap(lib, ow)
    # boxes whose min faces touch the cube's max faces, boxes that overlap it, boxes one voxel away
    g = np.arange(lo + 12, lo + 24)
    gx, gy, gz = np.meshgrid(g, g, np.arange(12, 22), indexing="ij")
    for off in (0.0, 0.5):
        pts = np.stack([(gx.ravel() + off) * res, (gy.ravel() + off) * res, (gz.ravel() + off) * res], 1)
        for size in ([2 * res] * 3, [res] * 3, [4 * res, 2 * res, 6 * res]):
            want = ow.getCellStatusBoundingBox(pts, size)
            got = hm.bbox(pts, size, 32)
            bad = np.flatnonzero(got != want)
            assert bad.size == 0, (off, size, bad[:5], pts[bad[:5]], got[bad[:5]], want[bad[:5]])
            assert {0, 1} <= set(np.unique(want).tolist())
    hm.close()


@pytest.mark.parametrize("res", [0.25, 0.125])
def test_face_aligned_boxes_at_binary_resolutions(lib, res):
    """Box faces exactly on voxel faces (exactly representable coordinates), solid occupied boxes
    that prune into large leaves.  The oracle tree is pruned first: lazy box edits leave equal
    siblings un-merged upstream, and the lower-side quirk depends on that history."""
    ow = build_world(res, False, n_scans=2, seed=1)
    ow.setOccupied(np.array([[2.0, 2.0, 1.0], [-3.0, 1.0, 1.0]]), [2.0, 2.0, 2.0])
    ow.prune()
    hm = HostMap(lib, ow)
    g = np.arange(-24, 25)
    cx, cy = np.meshgrid(g, g)
    for z in (1.0, res / 2 + 1.0):
        for off in (0.0, res / 2):
            pts = np.stack([cx.ravel() * res + off, cy.ravel() * res + off, np.full(cx.size, z)], 1)
            for size in ([2 * res] * 3, [res] * 3, [4 * res, 2 * res, 2 * res]):
                want = ow.getCellStatusBoundingBox(pts, size)
                got = hm.bbox(pts, size, 32)
                bad = np.flatnonzero(got != want)
                assert bad.size == 0, (z, off, size, bad[:5], pts[bad[:5]], got[bad[:5]], want[bad[:5]])
    hm.close()


def test_outside_the_map_and_empty_map(lib):
    res = 0.1
    ow = build_world(res, False, n_scans=1)
    hm = HostMap(lib, ow)
    far = np.array([[4000.0, 0, 0], [0, -3300.0, 1.0], [3276.75, 0.0, 0.0], [1e9, 1e9, 1e9], [np.nan, 0, 0]])
    for size in ([0.3, 0.3, 0.3], [1.0, 1.0, 1.0]):
        assert np.array_equal(hm.bbox(far, size, 32), ow.getCellStatusBoundingBox(far, size))
    hm.close()
    empty = oracle.OracleWorld(resolution=0.1)
    hm = HostMap(lib, empty)
    q = np.zeros((3, 3))
    assert np.array_equal(hm.bbox(q, [0.5, 0.5, 0.5], 32), empty.getCellStatusBoundingBox(q, [0.5, 0.5, 0.5]))
    hm.close()


def test_path_collision_index(lib):
    ow = build_world(0.1, False)
    hm = HostMap(lib, ow)
    # a straight path from the sensor towards a wall: the first colliding pose is where the robot
    # box first touches something that is not free
    path = np.stack([np.linspace(0.0, 9.0, 91), np.full(91, 0.3), np.full(91, 1.2)], 1)
    robot = [0.5, 0.5, 0.5]
    hit, idx = ow.checkPathForCollisionsWithRobot(path, robot)
    st = hm.bbox(path, robot, 32)
    first = np.flatnonzero(st == 1)          # unknown is not a collision when it is not treated as occupied
    assert hit == (first.size > 0)
    assert hit and idx == int(first[0]) and 0 < idx < 90
    hm.close()


@pytest.mark.parametrize("res,unk", [(0.1, True), (0.15, False)])
def test_point_line_and_visibility_queries_match_oracle(lib, res, unk):
    """The per-thread bodies of k_query_points / k_query_lines / k_query_visibility (the kernels
    the GPU suite checks end to end) on the CPU."""
    ow = build_world(res, unk)
    hm = HostMap(lib, ow)
    rng = np.random.default_rng(11)
    pts = rng.uniform([-8, -7, -1.0], [9, 8, 3.5], size=(20000, 3))
    pts[:5] = [[4000.0, 0, 0], [0, 0, -3300.0], [np.nan, 0, 0], [0.05, 0.05, 1.25], [1e12, -1e12, 0]]
    st, lo = hm.points(pts)
    assert np.array_equal(st, ow.getCellStatusPoint(pts))
    want_st, want_p = ow.getCellProbabilityPoint(pts)
    st_true, lo = hm.points(pts, true_status=1)
    assert np.array_equal(st_true, want_st)
    known = want_st != 2
    p = 1.0 - 1.0 / (1.0 + np.exp(lo[known].astype(np.float64)))      # OcTreeNode::getOccupancy()
    assert np.array_equal(p, want_p[known]) and np.all(np.isnan(lo[~known]))
    a = rng.uniform([-7, -6, 0.0], [8, 7, 3.0], size=(6000, 3))
    d = rng.normal(size=(6000, 3))
    b = a + d / np.linalg.norm(d, axis=1, keepdims=True) * rng.uniform(0.05, 9.0, (6000, 1))
    ab = np.hstack([a, b])
    ab[0] = [0.05, 0.05, 1.2, 0.05, 0.05, 1.2]               # same voxel: empty ray -> free
    ab[1] = [0.05, 0.05, 1.2, 5000.0, 0.0, 1.2]              # end outside the map: no keys -> free
    want = ow.getLineStatus(ab)
    got = hm.lines(ab)
    assert np.array_equal(got, want)
    assert set(np.unique(want).tolist()) >= ({0, 1} if unk else {0, 1, 2})
    for stop in (0, 1):
        assert np.array_equal(hm.visibility(ab, stop), ow.getVisibility(ab, stop))
    hm.close()
