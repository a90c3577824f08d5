"""The front-end arithmetic of the insertion kernels (vmap_front.cuh), built for the host and run on
the CPU: cv::reprojectImageTo3D against the committed cv2 4.13 golden vectors and the oracle, the
PCL matrix transform and the kindr / Eigen quaternion transform against operation-by-operation
numpy restatements, castRay's range test / clip / endpoint key against numpy float32 arithmetic."""
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
SRC = os.path.join(NATIVE, "front_check.cc")
LIB = os.path.join(NATIVE, "libfront_check.so")
GOLDEN = os.path.join(HERE, "golden")
F, D = np.float32, np.float64


@pytest.fixture(scope="module")
def lib():
    deps = [SRC, os.path.join(NATIVE, "cuda_host_shim.h")] + [os.path.join(CSRC, f) for f in
                                                               ("vmap_device.cuh", "vmap_front.cuh")]
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared",
                               "-DVMAP_HOST_EMULATION", "-include", os.path.join(NATIVE, "cuda_host_shim.h"),
                               "-x", "c++", SRC, "-o", LIB])
    L = C.CDLL(LIB)
    vp = C.c_void_p
    L.fc_reproject.argtypes = [vp, C.c_int, C.c_int, vp, vp]
    L.fc_matrix_transform.argtypes = [vp, vp, C.c_size_t, vp]
    L.fc_quat_transform.argtypes = [vp, vp, vp, C.c_size_t, vp, vp]
    L.fc_classify.argtypes = [C.c_double, C.c_double, vp, vp, C.c_size_t, vp, vp, vp, vp]
    return L


def reproject(L, disp, Q):
    disp = np.ascontiguousarray(disp, F)
    Q = np.ascontiguousarray(Q, D).reshape(16)
    out = np.zeros(disp.shape + (3,), F)
    L.fc_reproject(disp.ctypes.data, disp.shape[0], disp.shape[1], Q.ctypes.data, out.ctypes.data)
    return out


def bits(a):
    return np.ascontiguousarray(a, F).view(np.uint32)


@pytest.mark.parametrize("name", ["reproject_small_a", "reproject_small_b", "reproject_nomissing"])
def test_reproject_pixel_matches_cv2_golden(lib, name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    assert np.array_equal(bits(reproject(lib, g["disparity"], g["Q"])), bits(g["xyz"]))


def test_reproject_pixel_matches_oracle_at_cfg2_size(lib):
    c = scenes.cfg2_disparity()
    Q = oracle.fixup_q(c["Q"], c["full_image_width"], c["disparity"].shape[1]) if "full_image_width" in c else c["Q"]
    assert np.array_equal(bits(reproject(lib, c["disparity"], Q)), bits(oracle.reproject(c["disparity"], Q)))


def test_matrix_transform_is_pcl_arithmetic(lib):
    rng = np.random.default_rng(0)
    T = scenes.quat_to_matrix(scenes.random_unit_quat(rng), rng.normal(size=3) * 20.0)
    p = (rng.normal(size=(50000, 3)) * 30.0).astype(F)
    p[:3] = [[0, 0, 0], [1e-30, -1e30, 3.0], [np.inf, 1.0, -1.0]]
    out = np.zeros_like(p)
    Tm = np.ascontiguousarray(T, D)
    lib.fc_matrix_transform(Tm.ctypes.data, p.ctypes.data, len(p), out.ctypes.data)
    x, y, z = (p[:, i].astype(D) for i in range(3))
    with np.errstate(all="ignore"):
        want = np.stack([(((T[r, 0] * x + T[r, 1] * y) + T[r, 2] * z) + T[r, 3]).astype(F) for r in range(3)], 1)
    assert np.array_equal(bits(out), bits(want))


def test_quat_transform_is_eigen_arithmetic_and_validity(lib):
    rng = np.random.default_rng(1)
    q = scenes.random_unit_quat(rng)
    t = rng.normal(size=3) * 5.0
    p = (rng.normal(size=(50000, 3)) * 10.0).astype(F)
    p[:6, 2] = [10000.0, np.inf, -np.inf, -0.5, 0.0, -0.0]
    out = np.zeros_like(p)
    valid = np.zeros(len(p), np.uint8)
    qa, ta = np.ascontiguousarray(q, D), np.ascontiguousarray(t, D)
    lib.fc_quat_transform(qa.ctypes.data, ta.ctypes.data, p.ctypes.data, len(p), out.ctypes.data, valid.ctypes.data)
    # Eigen: uv = 2 * vec(q) x v;  v' = v + w * uv + vec(q) x uv;  + t
    w, x, y, z = q
    v0, v1, v2 = (p[:, i].astype(D) for i in range(3))
    with np.errstate(all="ignore"):
        u0, u1, u2 = y * v2 - z * v1, z * v0 - x * v2, x * v1 - y * v0
        u0, u1, u2 = u0 + u0, u1 + u1, u2 + u2
        c0, c1, c2 = y * u2 - z * u1, z * u0 - x * u2, x * u1 - y * u0
        want = np.stack([((v0 + w * u0) + c0) + t[0], ((v1 + w * u1) + c1) + t[1], ((v2 + w * u2) + c2) + t[2]], 1).astype(F)
    assert np.array_equal(bits(out), bits(want))
    pz = p[:, 2]
    assert np.array_equal(valid.astype(bool), (pz != F(10000.0)) & ~np.isinf(pz) & ~(pz < 0))
    assert valid[:6].tolist() == [0, 0, 0, 0, 1, 1]


@pytest.mark.parametrize("res,max_range", [(0.1, 5.0), (0.15, -1.0), (0.05, 50.0)])
def test_classify_math_is_castray_arithmetic(lib, res, max_range):
    rng = np.random.default_rng(2)
    o = np.array([0.37, -1.2, 1.8], F)
    p = (o + rng.normal(size=(60000, 3)) * rng.uniform(0.01, 3.0 * abs(max_range), (60000, 1))).astype(F)
    p[0] = o                                   # zero-length ray
    p[1] = [4000.0, 0.0, 0.0]                  # endpoint outside the map
    p[2] = o + F(max_range if max_range > 0 else 1.0) * np.array([1, 0, 0], F)
    key = np.zeros(len(p), np.uint64); flags = np.zeros(len(p), np.uint8)
    end = np.zeros_like(p); ln = np.zeros(len(p), np.uint32)
    lib.fc_classify(res, max_range, o.ctypes.data, p.ctypes.data, len(p), key.ctypes.data, flags.ctypes.data,
                    end.ctypes.data, ln.ctypes.data)
    rf = 1.0 / res
    with np.errstate(all="ignore"):
        d = (p - o).astype(F)
        nsq = ((d[:, 0] * d[:, 0]).astype(F) + (d[:, 1] * d[:, 1]).astype(F)).astype(F)
        nsq = (nsq + (d[:, 2] * d[:, 2]).astype(F)).astype(F)
        norm = np.sqrt(nsq.astype(D))
        in_range = np.full(len(p), True) if max_range < 0 else norm <= max_range
        dn = np.where((norm > 0)[:, None], (d / norm.astype(F)[:, None]).astype(F), d)
        clipped = (o + (dn * F(max_range)).astype(F)).astype(F)
        want_end = np.where(in_range[:, None], p, clipped)
        sk = np.floor(rf * p.astype(D)).astype(np.int64) + 32768
        bound = np.all((sk >= 0) & (sk < 65536), axis=1)
        want_key = (sk[:, 0] & 0xFFFF) | ((sk[:, 1] & 0xFFFF) << 16) | ((sk[:, 2] & 0xFFFF) << 32)
        ek = np.floor(rf * want_end.astype(D)).astype(np.int64) + 32768
        ok_ = np.floor(rf * o.astype(D)).astype(np.int64) + 32768
        walk = np.all((ek >= 0) & (ek < 65536), axis=1)
        want_len = np.where(walk, np.abs(ek - ok_).sum(1), 0)
    assert np.array_equal((flags & 2) != 0, in_range)
    assert np.array_equal((flags & 1) != 0, bound)
    assert np.array_equal(bits(end), bits(want_end))
    assert np.array_equal(key.astype(np.int64), want_key)
    assert np.array_equal(ln.astype(np.int64), want_len)
    if max_range > 0:
        assert 0.05 < in_range.mean() < 0.95


def test_update_leaf_ladders_match_oracle(lib):
    """updateNode / updateNodeLogOdds on one voxel (early abort at the clamps, new node = 0,
    clamp order): the product's update_leaf against the oracle's tree, step by step."""
    lib.fc_update_sequence.argtypes = [C.c_void_p, C.c_size_t, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p]
    lib.fc_update_sequence.restype = C.c_uint32
    rng = np.random.default_rng(3)
    for params in ({}, {"probability_hit": 0.9, "probability_miss": 0.2, "threshold_min": 0.3, "threshold_max": 0.8}):
        for trial in range(6):
            ow = oracle.OracleWorld(resolution=0.1, **params)
            hit, miss, cmin, cmax, _ = (np.float32(v) for v in ow.logodds_constants())
            seq = (rng.random(60) < (0.2 + 0.12 * trial)).astype(np.uint8)
            trace = np.zeros(len(seq), np.uint32)
            lib.fc_update_sequence(seq.ctypes.data, len(seq), hit, miss, cmin, cmax, trace.ctypes.data)
            key = np.array([[32768 + 3, 32768 - 4, 32768 + 5]], np.uint16)
            none = np.zeros((0, 3), np.uint16)
            for i, h in enumerate(seq):
                ow.updateOccupancy(none if h else key, key if h else none)
                _, v = ow.export_leaves()
                assert len(v) == 1 and np.float32(v[0]).view(np.uint32) == trace[i], (params, trial, i)
