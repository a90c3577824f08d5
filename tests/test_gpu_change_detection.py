"""Change detection (OctomapParameters::change_detection_enabled, getChangedPoints,
octomap_world.cc:99,1413-1440): the set of changed keys and the occupancy reported for each must
equal the oracle's (octomap's KeyBoolMap semantics: created leaves, flips, flip-backs erased)."""
import numpy as np
import pytest

import oracle
import volumetric_mapping_b200 as vm

pytestmark = pytest.mark.gpu
I4 = np.eye(4)


def as_dict(k, o):
    return {tuple(kk): bool(oo) for kk, oo in zip(k.tolist(), o.tolist())}


@pytest.mark.parametrize("mode_kw", [dict(), dict(dense_window_bytes=1), dict(dedupe_mode=vm.DEDUPE_SORT)],
                         ids=["dense", "table", "sort"])
def test_changed_points_match_oracle(mode_kw):
    rng = np.random.default_rng(51)
    kw = dict(resolution=0.2, sensor_max_range=5.0, change_detection_enabled=1)
    gw, ow = vm.OctomapWorld(initial_blocks=64, **mode_kw, **kw), oracle.OracleWorld(**kw)
    d = rng.normal(size=(2500, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    base = (d * rng.uniform(1.0, 6.0, (2500, 1))).astype(np.float32)
    seen_nonempty = 0
    for scan in range(9):
        # alternate two clouds so voxels flip occupied -> free -> occupied between reads
        pts = base if scan % 3 != 2 else (base * np.float32(1.15))
        gw.insertPointcloud(I4, pts)
        ow.insertPointcloud(I4, pts)
        if scan % 2 == 1 or scan == 8:
            g, o = as_dict(*gw.getChangedPoints()), as_dict(*ow.getChangedPoints())
            assert g == o
            seen_nonempty += bool(g)
            assert as_dict(*gw.getChangedPoints()) == {} and as_dict(*ow.getChangedPoints()) == {}
    assert seen_nonempty >= 3


def test_changed_points_need_the_parameter():
    gw = vm.OctomapWorld(resolution=0.2)
    gw.insertPointcloud(I4, [[1.0, 0.1, 0.1]])
    with pytest.raises(vm.VmapError):
        gw.getChangedPoints()
    p = gw.getOctomapParameters()
    p.change_detection_enabled = 1
    gw.setOctomapParameters(p)                       # enableChangeDetection(true) on the live map
    gw.insertPointcloud(I4, [[1.0, 0.1, 0.1], [0.1, 1.0, 0.1]])
    k, o = gw.getChangedPoints()
    # second hit makes the first end voxel occupied (flip -> recorded), the second ray is new
    d = as_dict(k, o)
    assert d[(32773, 32768, 32768)] is True and len(d) > 1
