"""The closed-form jump-ahead of the DDA's `tMax += tDelta` recurrence (vmap_seq.cuh) against the
plain sequential loop: bit-identical element and identical count, on random DDA-like inputs, on
constructed round-half-even ties and across binade boundaries."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "native", "seq_check.cc")
LIB = os.path.join(HERE, "native", "libseq_check.so")


@pytest.fixture(scope="module")
def lib():
    hdr = os.path.join(HERE, "..", "volumetric_mapping_b200", "csrc", "vmap_seq.cuh")
    if (not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(hdr))):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared",
                               "-x", "c++", SRC, "-o", LIB])
    L = C.CDLL(LIB)
    dp, up = C.POINTER(C.c_double), C.POINTER(C.c_uint64)
    L.seq_check_many.argtypes = [dp, dp, dp, up, C.c_uint64, up]
    L.seq_check_many.restype = C.c_uint64
    L.seq_advance_closed.argtypes = [dp, C.c_double, C.c_double, C.c_uint64]
    L.seq_advance_closed.restype = C.c_uint64
    L.seq_advance_loop.argtypes = [dp, C.c_double, C.c_double, C.c_uint64]
    L.seq_advance_loop.restype = C.c_uint64
    return L


def run(L, t0, dt, tau, cap):
    t0 = np.ascontiguousarray(t0, np.float64); dt = np.ascontiguousarray(dt, np.float64)
    tau = np.ascontiguousarray(tau, np.float64); cap = np.ascontiguousarray(cap, np.uint64)
    first = C.c_uint64(0)
    dp, up = C.POINTER(C.c_double), C.POINTER(C.c_uint64)
    bad = L.seq_check_many(t0.ctypes.data_as(dp), dt.ctypes.data_as(dp), tau.ctypes.data_as(dp),
                           cap.ctypes.data_as(up), len(t0), C.byref(first))
    assert bad == 0, (bad, first.value, t0[first.value].hex(), dt[first.value].hex(),
                      tau[first.value].hex(), int(cap[first.value]))


def test_dda_like_random_cases(lib):
    rng = np.random.default_rng(0)
    n = 300000
    res = rng.choice([0.05, 0.1, 0.15, 0.2, 0.0371], n)
    d = np.abs(rng.normal(size=n)).astype(np.float32).astype(np.float64)
    d = np.clip(d / 3.0, 1e-4, 1.0)
    dt = res / d
    t0 = rng.uniform(0, 1, n) * dt * rng.choice([1.0, 1e-3, 1e-9], n)
    tau = rng.uniform(0, 120.0, n)
    cap = np.full(n, 1 << 40, np.uint64)
    run(lib, t0, dt, tau, cap)
    # with a step cap
    run(lib, t0, dt, np.full(n, np.inf), rng.integers(0, 3000, n).astype(np.uint64))
    # tiny negative / zero starts (origin on a voxel face)
    run(lib, -np.abs(t0) * 1e-12, dt, tau, cap)
    run(lib, np.zeros(n), dt, tau, cap)


def test_constructed_ties_and_binade_crossings(lib):
    rng = np.random.default_rng(1)
    t0s, dts, taus = [], [], []
    for e in range(-6, 8):
        u = 2.0 ** (e - 52)
        for _ in range(400):
            q0 = int(rng.integers(1 << 40, 1 << 50))
            frac = rng.choice([0.5, 0.25, 0.75, 0.0, 0.5 - 2.0 ** -10, 0.5 + 2.0 ** -10])
            dt = (q0 + frac) * u                      # exactly representable: q0 < 2^50
            m = int(rng.integers(1 << 52, 1 << 53))
            t0 = m * u                                # anywhere in the binade, odd or even
            t0s += [t0, t0 * 0.5, t0 * 0.25]
            dts += [dt, dt, dt]
            hi = 2.0 ** (e + 3)
            taus += [rng.uniform(t0, hi), hi, rng.uniform(t0, hi)]
    n = len(t0s)
    run(lib, t0s, dts, taus, np.full(n, 1 << 40, np.uint64))
    run(lib, t0s, dts, np.full(n, np.inf), rng.integers(1, 200, n).astype(np.uint64))


def test_tau_on_grid_points_and_edges(lib):
    # tau exactly equal to an element, just below and just above it
    rng = np.random.default_rng(2)
    t0s, dts, taus = [], [], []
    for _ in range(3000):
        dt = float(rng.uniform(0.05, 3.0))
        t = float(rng.uniform(0, dt))
        k = int(rng.integers(1, 2000))
        x = np.float64(t)
        for _ in range(k):
            x = np.float64(x + np.float64(dt))
        for tau in (x, np.nextafter(x, 0.0), np.nextafter(x, np.inf)):
            t0s.append(t); dts.append(dt); taus.append(float(tau))
    run(lib, t0s, dts, taus, np.full(len(t0s), 1 << 40, np.uint64))
    # degenerate: tau <= t0, tau NaN, huge dt
    run(lib, [1.0, 1.0, 0.3, 5.0], [0.1, 0.1, 1e300, 1.7976931348623157e308],
        [0.5, np.nan, 1e308, np.inf], np.array([10, 10, 10, 3], np.uint64))


def test_long_runs_hit_every_binade(lib):
    rng = np.random.default_rng(3)
    n = 2000
    dt = rng.uniform(0.05, 0.3, n)
    t0 = rng.uniform(0, 0.05, n)
    run(lib, t0, dt, np.full(n, 5000.0), np.full(n, 1 << 40, np.uint64))
    t = C.c_double(0.0123)
    k = lib.seq_advance_closed(C.byref(t), 0.1, 1e6, 1 << 40)
    t2 = C.c_double(0.0123)
    assert k == lib.seq_advance_loop(C.byref(t2), 0.1, 1e6, 1 << 40) and t.value == t2.value
