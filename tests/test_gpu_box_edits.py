"""setFree / setOccupied / setLogOddsBoundingBox (octomap_world.cc:497-523, 781-818, 1175-1224):
the box loops run on the host in the reference's arithmetic, the voxels are written on the device.
GPU map vs oracle tree after edits interleaved with scans, all three BoundHandling modes."""
import numpy as np
import pytest

import oracle
import volumetric_mapping_b200 as vm

pytestmark = pytest.mark.gpu
I4 = np.eye(4)


def leaves(w):
    k, v = w.export_leaves()
    p = oracle.pack_keys(k)
    o = np.argsort(p)
    return p[o], v.view(np.uint32)[o]


@pytest.mark.parametrize("mode", [0, 1, 2], ids=["default", "include_partial", "ignore_partial"])
def test_box_edits_match_oracle(mode):
    rng = np.random.default_rng(40 + mode)
    kw = dict(resolution=0.15, sensor_max_range=5.0)
    gw, ow = vm.OctomapWorld(**kw), oracle.OracleWorld(**kw)
    d = rng.normal(size=(3000, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    pts = (d * rng.uniform(1.0, 6.0, (3000, 1))).astype(np.float32)
    for w in (gw, ow):
        w.insertPointcloud(I4, pts)
        w.setFree([[0.3, -0.2, 0.1], [2.05, 0.45, -0.3]], [1.2, 0.9, 0.6], mode)
        w.setOccupied([[-1.0, 1.0, 0.525]], [0.45, 0.3, 0.15], mode)
        w.setLogOddsBoundingBox([[4000.0, 0.0, 0.0]], [0.5, 0.5, 0.5], 0.25, mode)   # out of bounds: ignored
        w.setLogOddsBoundingBox([[0.0, 3.0, 0.0]], [0.3, 0.3, 0.3], -0.75, mode)     # value inside the clamps: kept
        w.insertPointcloud(I4, pts * np.float32(0.9))
    gk, gv = leaves(gw)
    ok, ov = leaves(ow)
    assert np.array_equal(gk, ok) and np.array_equal(gv, ov)
    assert 0xbf400000 in set(gv.tolist()) or mode == 2      # -0.75 survived where no ray touched it
    q = rng.uniform(-3, 3, (5000, 3))
    assert np.array_equal(gw.getCellStatusPoint(q), ow.getCellStatusPoint(q))


def test_box_edit_values_are_clamped():
    """OccupancyOcTreeBase::setNodeValue clamps to [clamping_thres_min, clamping_thres_max] before it
    writes: a box set to +-9 holds exactly the clamp constants (known answer, and the oracle agrees)."""
    kw = dict(resolution=0.2, sensor_max_range=5.0)
    gw, ow = vm.OctomapWorld(**kw), oracle.OracleWorld(**kw)
    for w in (gw, ow):
        w.setLogOddsBoundingBox([[1.0, 1.0, 1.0]], [0.4, 0.4, 0.4], 9.0, 0)
        w.setLogOddsBoundingBox([[-1.0, -1.0, 1.0]], [0.4, 0.4, 0.4], -9.0, 0)
    cmin, cmax = gw.logodds_constants()[2:4]
    gk, gv = leaves(gw)
    ok, ov = leaves(ow)
    assert np.array_equal(gk, ok) and np.array_equal(gv, ov)
    assert set(gv.tolist()) == {int(np.float32(cmin).view(np.uint32)), int(np.float32(cmax).view(np.uint32))}
