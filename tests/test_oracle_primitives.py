"""Known-answer tests that pin the CPU oracle (the reference ships none).

Every expectation here is derivable by hand from SURVEY.md section 3.3 or was probed with an
independent tool (numpy float32 arithmetic, pure-Python DDA in tests/pyref.py).
"""
import math

import numpy as np
import pytest

import oracle
import pyref


@pytest.fixture(scope="module")
def w():
    return oracle.OracleWorld(resolution=0.1)


def test_logodds_constants_bit_patterns():
    w = oracle.OracleWorld()
    got = [hex(v) for v in w.logodds_constants().view(np.uint32)]
    # hit(0.65) miss(0.4) cmin(0.12) cmax(0.97) occ(0.7) -- SURVEY 3.3
    assert got == ["0x3f1e795b", "0xbecf991f", "0xbfff07f4", "0x405e7867", "0x3f58e883"]


def test_coord_to_key_and_back(w):
    assert w.coordToKey(0.0) == 32768
    assert w.coordToKey(0.05) == 32768
    assert w.coordToKey(-0.05) == 32767
    assert w.coordToKey(0.1) == 32768 + 1 or w.coordToKey(0.1) == 32768  # 10*0.1 rounding
    assert w.coordToKey(0.1) == pyref.coord_to_key(0.1, 0.1)
    assert w.keyToCoord(32768) == pytest.approx(0.05)
    assert w.keyToCoord(0) == (-32768 + 0.5) * 0.1
    # checked variant: bounds are +-32768 voxels
    assert w.coordToKeyChecked(3276.75) == (True, 65535)
    assert w.coordToKeyChecked(3276.85)[0] is False
    assert w.coordToKeyChecked(-3276.8)[0] is True
    assert w.coordToKeyChecked(-3276.81)[0] is False
    # unchecked wraps modulo 2^16
    assert w.coordToKey(3276.85) == (65536 + 0) & 0xFFFF
    # NaN / inf / huge follow x86 cvttsd2si: INT_MIN + 32768 -> low 16 bits = 0x8000
    for bad in (float("nan"), float("inf"), -float("inf"), 1e300):
        assert w.coordToKey(bad) == 0x8000
        assert w.coordToKeyChecked(bad)[0] is False


def test_ray_axis_aligned(w):
    keys = w.computeRayKeys((0.05, 0.05, 0.05), (0.55, 0.05, 0.05))
    expect = [(32768 + i, 32768, 32768) for i in range(5)]     # origin + intermediates, no end
    assert [tuple(k) for k in keys] == expect


def test_ray_negative_direction(w):
    keys = w.computeRayKeys((0.05, 0.05, 0.05), (0.05, -0.35, 0.05))
    assert [tuple(k) for k in keys] == [(32768, 32768 - i, 32768) for i in range(4)]


def test_ray_same_voxel_is_empty(w):
    keys = w.computeRayKeys((0.01, 0.01, 0.01), (0.09, 0.09, 0.09))
    assert keys is not None and len(keys) == 0


def test_ray_out_of_bounds_returns_false(w):
    assert w.computeRayKeys((0, 0, 0), (4000.0, 0, 0)) is None
    assert w.computeRayKeys((0, 0, -4000.0), (1, 0, 0)) is None
    assert w.computeRayKeys((0, 0, 0), (float("nan"), 0, 0)) is None


def test_ray_diagonal_tie_goes_to_later_axis(w):
    # exact diagonal from a voxel centre: tMax equal on all axes -> dim 2, then 1, then 0
    keys = w.computeRayKeys((0.05, 0.05, 0.05), (0.25, 0.25, 0.25))
    ks = [tuple(k) for k in keys]
    assert ks[0] == (32768, 32768, 32768)
    assert ks[1] == (32768, 32768, 32769)
    assert ks[2] == (32768, 32769, 32769)
    assert ks[3] == (32769, 32769, 32769)
    assert ks == [tuple(k) for k in pyref.compute_ray_keys((0.05, 0.05, 0.05), (0.25, 0.25, 0.25), 0.1)]


def test_ray_matches_python_restatement_random():
    rng = np.random.default_rng(5)
    for res in (0.05, 0.1, 0.15):
        w = oracle.OracleWorld(resolution=res)
        for _ in range(150):
            o = rng.uniform(-3, 3, 3).astype(np.float32)
            e = (o + rng.normal(size=3) * rng.uniform(0.01, 6.0)).astype(np.float32)
            got = w.computeRayKeys(o, e)
            exp = pyref.compute_ray_keys(o, e, res)
            assert (got is None) == (exp is None)
            if got is not None:
                assert [tuple(k) for k in got] == exp


def test_ray_structure_properties():
    rng = np.random.default_rng(6)
    w = oracle.OracleWorld(resolution=0.1)
    for _ in range(200):
        o = rng.uniform(-5, 5, 3).astype(np.float32)
        e = rng.uniform(-5, 5, 3).astype(np.float32)
        keys = w.computeRayKeys(o, e)
        ko = tuple(w.coordToKey(float(c)) for c in o)
        ke = tuple(w.coordToKey(float(c)) for c in e)
        ks = [tuple(k) for k in keys]
        if ko == ke:
            assert ks == []
            continue
        assert ks[0] == ko                       # origin included
        assert ke not in ks                      # end voxel never included
        for a, b in zip(ks[:-1], ks[1:]):        # 6-connected walk
            assert sum(abs(int(x) - int(y)) for x, y in zip(a, b)) == 1
        assert len(set(ks)) == len(ks)


def test_hit_miss_ladders_bit_exact():
    # SURVEY 3.3 ladders, float32 bit patterns, via updateOccupancy on a single key
    w = oracle.OracleWorld()
    key = np.array([[32770, 32771, 32772]], dtype=np.uint16)
    none = np.zeros((0, 3), dtype=np.uint16)
    hits = []
    for _ in range(7):
        w.updateOccupancy(none, key)
        k, v = w.export_leaves()
        assert k.shape[0] == 1 and tuple(k[0]) == (32770, 32771, 32772)
        hits.append(hex(v.view(np.uint32)[0]))
    assert hits[:6] == ["0x3f1e795b", "0x3f9e795b", "0x3fedb608", "0x401e795b", "0x404617b2",
                        "0x405e7867"]
    assert hits[6] == "0x405e7867"
    w.resetMap()
    miss = []
    for _ in range(6):
        w.updateOccupancy(key, none)
        miss.append(hex(w.export_leaves()[1].view(np.uint32)[0]))
    assert miss == ["0xbecf991f", "0xbf4f991f", "0xbf9bb2d7", "0xbfcf991f", "0xbfff07f4",
                    "0xbfff07f4"]


def test_update_is_not_commutative_across_scans():
    key = np.array([[32768, 32768, 32768]], dtype=np.uint16)
    none = np.zeros((0, 3), dtype=np.uint16)

    def run(seq):
        w = oracle.OracleWorld()
        for c in seq:
            if c == "H":
                w.updateOccupancy(none, key)
            else:
                w.updateOccupancy(key, none)
        return hex(w.export_leaves()[1].view(np.uint32)[0])

    assert run("HHHHHHM") == "0x40448543"
    assert run("MHHHHHH") == "0x4053c2e5"


def test_occupied_wins_over_free_in_same_scan():
    w = oracle.OracleWorld()
    key = np.array([[100, 200, 300]], dtype=np.uint16)
    w.updateOccupancy(key, key)
    k, v = w.export_leaves()
    assert hex(v.view(np.uint32)[0]) == "0x3f1e795b"
    f, o = w.last_scan_keys()
    assert len(f) == 0 and len(o) == 1


def test_one_hit_is_free_two_hits_occupied_unknown_policy():
    p = oracle.default_params(treat_unknown_as_occupied=0)
    w = oracle.OracleWorld(p)
    key = np.array([[32768, 32768, 32768]], dtype=np.uint16)
    none = np.zeros((0, 3), dtype=np.uint16)
    pt = np.array([[0.01, 0.01, 0.01]])
    assert w.getCellStatusPoint(pt)[0] == 2
    w.updateOccupancy(none, key)
    assert w.getCellStatusPoint(pt)[0] == 0       # 0.619 < 0.847
    w.updateOccupancy(none, key)
    assert w.getCellStatusPoint(pt)[0] == 1
    w2 = oracle.OracleWorld()
    assert w2.getCellStatusPoint(pt)[0] == 1       # unknown -> occupied by default
    assert w2.getCellStatusPoint(np.array([[1e9, 0, 0]]))[0] == 1   # out of bounds -> NULL


def test_pruning_keeps_leaf_values_and_shrinks_tree():
    # 8 siblings with equal values collapse into their parent (lossless for leaf export)
    w = oracle.OracleWorld()
    sib = np.array([[32768 + (i & 1), 32768 + ((i >> 1) & 1), 32768 + ((i >> 2) & 1)]
                    for i in range(8)], dtype=np.uint16)
    none = np.zeros((0, 3), dtype=np.uint16)
    w.updateOccupancy(none, sib[:7])
    n7 = w.tree_node_count()
    w.updateOccupancy(none, sib[7:])
    assert w.tree_node_count() == n7 - 7          # +1 created, then 8 children pruned away
    k, v = w.export_leaves()
    assert k.shape[0] == 8 and len(set(v.view(np.uint32).tolist())) == 1
    # updating one child of a pruned node expands it again
    w.updateOccupancy(sib[:1], none)
    k, v = w.export_leaves()
    assert k.shape[0] == 8 and len(set(v.view(np.uint32).tolist())) == 2


def test_quat_to_matrix_matches_scene_helper():
    from volumetric_mapping_b200 import scenes
    rng = np.random.default_rng(3)
    for _ in range(20):
        q = scenes.random_unit_quat(rng)
        t = rng.uniform(-10, 10, 3)
        a = oracle.quat_to_matrix(q, t)
        b = scenes.quat_to_matrix(q, t)
        assert np.array_equal(a.view(np.uint64), b.view(np.uint64))
        assert np.allclose(a[:3, :3] @ a[:3, :3].T, np.eye(3), atol=1e-12)
