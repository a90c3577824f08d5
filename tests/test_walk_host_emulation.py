"""The product's per-thread walker code, built for the host (tests/native/walk_check.cc +
cuda_host_shim.h) and run one "thread" at a time on the CPU:

* walk_ray (vmap_device.cuh, the DDA every query / emit kernel uses) against the oracle's
  computeRayKeys, key by key;
* ray_checkpoints + walk_slot (vmap_walk.cuh: exact ray segmentation, the walker that ships) and
  + walk_slot_dilated (the experimental addressing / near-field variants) against walk_ray's marks
  for whole scans -- identical free-mask words,
  touched bitmap, traversal count and longest ray, with and without the free-space filter,
  including rays that end inside the origin voxel's block, axis-parallel rays, ties and rays
  longer than KeyRay's capacity.

The GPU suite checks the same kernels end to end; this file keeps the walker logic covered on
machines without a GPU."""
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
SRC = os.path.join(NATIVE, "walk_check.cc")
LIB = os.path.join(NATIVE, "libwalk_check.so")


@pytest.fixture(scope="module")
def lib():
    deps = [SRC, os.path.join(NATIVE, "cuda_host_shim.h")] + [os.path.join(CSRC, f) for f in
                                                               ("vmap_device.cuh", "vmap_seq.cuh", "vmap_walk.cuh")]
    # VMAP_TEST_SEG_STEPS=N: the same checks for another segment length (tools/ab_segsteps.py variants)
    seg = os.environ.get("VMAP_TEST_SEG_STEPS")
    out = LIB if not seg else os.path.join(NATIVE, "libwalk_check_seg%d.so" % int(seg))
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared",
                               "-DVMAP_HOST_EMULATION"] + (["-DVMAP_SEG_STEPS=%d" % int(seg)] if seg else []) +
                              ["-include", os.path.join(NATIVE, "cuda_host_shim.h"), "-x", "c++", SRC, "-o", out])
    L = C.CDLL(out)
    vp = C.c_void_p
    L.wc_ray_keys.argtypes = [C.c_double, vp, vp, vp, C.c_uint32]
    L.wc_ray_keys.restype = C.c_uint32
    L.wc_walk_plain.argtypes = [C.c_double, C.c_double, C.c_double, vp, vp, C.c_uint32, vp, vp, vp, vp, vp]
    L.wc_walk_segmented.argtypes = [C.c_double, C.c_double, C.c_double, vp, vp, C.c_uint32, vp, vp, vp, vp,
                                    C.c_uint32, C.c_int, vp]
    return L


def window_for(res, origin, ends):
    """A power-of-two window (in 8^3 blocks) around every voxel the rays can visit."""
    pts = np.vstack([np.asarray(origin, np.float64)[None], np.asarray(ends, np.float64)])
    k = np.floor(pts / res).astype(np.int64) + 32768
    # (like the product's window, at least five blocks on every side of the sensor's block)
    lo = np.minimum((k.min(0) - 2) >> 3, (k[0] >> 3) - 5)
    hi = np.maximum((k.max(0) + 2) >> 3, (k[0] >> 3) + 5)
    ext = hi - lo + 1
    dims = np.array([1 << int(np.ceil(np.log2(ext[0]))), 1 << int(np.ceil(np.log2(ext[1]))), ext[2]], np.uint32)
    return lo.astype(np.int32), dims


def run_both(L, res, origin, ends, max_free=0.0, min_height=0.0, near_limit=48):
    origin = np.ascontiguousarray(origin, np.float32)
    ends = np.ascontiguousarray(ends, np.float32)
    ob, dims = window_for(res, origin, ends)
    n_blocks = int(dims[0]) * int(dims[1]) * int(dims[2])
    out = []
    modes = ("plain", "seg3", "seg0", "seg1", "seg4", "seg5", "seg6") if max_free else ("plain", "seg3", "seg0", "seg1", "seg4", "seg5", "seg6", "seg2")
    # seg3: walk_slot (block addressing); seg0 / seg1: walk_slot_dilated<., mode>; seg4: the same with counted
    # loops; seg5: counted loops + 64-voxel mask words; seg6: the same with queued marks (walk_slot_queued, what
    # ships); seg2: near-field combining schedule
    for which in modes:
        fm = np.zeros(n_blocks * 16, np.uint32)
        tb = np.zeros((n_blocks + 31) // 32, np.uint32)
        stats = np.zeros(5, np.uint64)
        if which == "plain":
            L.wc_walk_plain(res, max_free, min_height, origin.ctypes.data, ends.ctypes.data, len(ends),
                            ob.ctypes.data, dims.ctypes.data, fm.ctypes.data, tb.ctypes.data, stats.ctypes.data)
        else:
            assert L.wc_walk_segmented(res, max_free, min_height, origin.ctypes.data, ends.ctypes.data, len(ends),
                                       ob.ctypes.data, dims.ctypes.data, fm.ctypes.data, tb.ctypes.data, near_limit,
                                       int(which[-1]), stats.ctypes.data) == 0
        out.append((fm, tb, stats))
    fa, ta, sa = out[0]
    for fb, tb_, sb in out[1:]:
        assert sa[2] == 0 and sb[2] == 0, "window overflow"
        assert int(sa[0]) == int(sb[0]), ("traversals", sa, sb)
        assert int(sa[1]) == int(sb[1]), ("longest ray", sa, sb)
        bad = np.flatnonzero(fa != fb)
        assert bad.size == 0, (bad[:5], fa[bad[:5]], fb[bad[:5]])
        assert np.array_equal(ta, tb_)
    return sa, out[-1][2] if len(out) == 8 else out[1][2]


def lidar_rays(n_elev, n_azim, seed, yaw):
    """Sensor origin and world-frame ray ends of a small spinning-lidar scan of the bench scene."""
    s = scenes.hdl64_scan(position=(1.3, -2.2, 1.8), yaw_rad=yaw, seed=seed, n_elev=n_elev, n_azim=n_azim)
    T = s["T"]
    world = (s["points"].astype(np.float64) @ T[:3, :3].T + T[:3, 3]).astype(np.float32)
    return T[:3, 3].astype(np.float32), world


def test_plain_dda_matches_the_oracle(lib):
    rng = np.random.default_rng(5)
    for res in (0.1, 0.15, 0.05):
        ow = oracle.OracleWorld(resolution=res)
        for _ in range(150):
            o = rng.uniform(-3, 3, 3).astype(np.float32)
            e = (o + rng.normal(size=3) * rng.uniform(0.01, 25.0)).astype(np.float32)
            if rng.random() < 0.2:
                e[rng.integers(3)] = o[rng.integers(3)]          # axis-parallel / degenerate components
            want = np.asarray(ow.computeRayKeys(o, e), np.uint16).reshape(-1, 3)
            got = np.zeros((max(1, len(want) + 8), 3), np.uint16)
            n = lib.wc_ray_keys(res, o.ctypes.data, e.ctypes.data, got.ctypes.data, len(got))
            assert n == len(want)
            assert np.array_equal(got[:n], want)


@pytest.mark.parametrize("res", [0.1, 0.05, 0.2])
def test_segmented_dilated_walk_equals_plain_walk_on_a_lidar_scan(lib, res):
    origin, world = lidar_rays(32, 360, seed=3, yaw=0.3)
    sa, sb = run_both(lib, res, origin, world)
    assert int(sb[3]) > len(world)            # rays really were cut into several segments
    assert int(sb[4]) >= 10                   # and whole first-segment chunks went through the near grid


def test_with_free_space_filter(lib):
    origin, world = lidar_rays(16, 180, seed=4, yaw=-1.1)
    run_both(lib, 0.1, origin, world, max_free=4.0, min_height=0.5)


def test_adversarial_rays(lib):
    rng = np.random.default_rng(11)
    res = 0.1
    o = np.array([0.05, 0.05, 0.05], np.float32)
    ends = []
    # exact diagonals (every step is a three-way tie), face diagonals, axis-parallel rays
    for L_ in (0.3, 3.1, 25.0):
        for d in ([1, 1, 1], [-1, 1, 1], [1, -1, 0], [0, 1, 1], [1, 0, 0], [0, -1, 0], [0, 0, 1], [-1, -1, -1]):
            ends.append(o + np.float32(L_) * np.asarray(d, np.float32))
    # ends on voxel borders, tiny rays inside one block, rays of one and two voxels
    ends += [o + np.array([0.05, 0.0, 0.0], np.float32), o + np.array([0.1, 0.1, 0.0], np.float32),
             o + np.array([0.15, -0.05, 0.25], np.float32)]
    d = rng.normal(size=(400, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    ends += list(o + (d * rng.uniform(0.05, 45.0, (400, 1))).astype(np.float32))
    run_both(lib, res, o, np.asarray(ends, np.float32), near_limit=0)
    run_both(lib, res, o, np.asarray(ends, np.float32), near_limit=0xFFFFFFFF)
    # a sensor off the voxel centre and at negative coordinates
    o2 = np.array([-7.33, 2.918, -0.4401], np.float32)
    run_both(lib, res, o2, o2 + np.clip(rng.normal(size=(500, 3)) * 15.0, -40.0, 40.0).astype(np.float32))


def test_very_long_rays_at_fine_resolution(lib):
    # 0.004 m voxels, rays of up to ~60000 keys (~900 segments each, many binade crossings of
    # tMax).  A ray that reaches KeyRay's capacity (100000 keys) cannot occur in a window that
    # fits the 32-bit dilated address -- the GPU suite covers that case on the default walker.
    res = 0.004
    o = np.array([-120.0, 0.002, 0.003], np.float32)
    ends = np.array([[120.0, 0.3, 0.01], [110.0, -0.2, 0.02], [-117.0, 0.2, 0.01], [-121.0, 0.1, 0.0]], np.float32)
    sa, sb = run_both(lib, res, o, ends)
    assert int(sa[1]) > 59000 and int(sb[3]) > 1500

