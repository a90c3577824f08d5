"""Deterministic synthetic inputs for the BASELINE.json configs.

numpy only; no dependency on the CUDA library or on the oracle, so tests, bench.py and the
C++ demo all consume identical bytes.
"""
import math

import numpy as np


def yaw_quat(yaw_rad):
    """Unit quaternion (w, x, y, z) for a rotation about +z."""
    return np.array([math.cos(0.5 * yaw_rad), 0.0, 0.0, math.sin(0.5 * yaw_rad)], dtype=np.float64)


def random_unit_quat(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    return q.astype(np.float64)


def quat_to_matrix(q_wxyz, t):
    """Eigen::Quaterniond::toRotationMatrix() association order + translation -> 4x4 double.

    Same arithmetic (operation order) as QuatTransformation::getTransformationMatrix() so the
    host shim, the oracle and this helper agree bit for bit.
    """
    w, x, y, z = (float(v) for v in q_wxyz)
    tx, ty, tz = 2.0 * x, 2.0 * y, 2.0 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    T = np.eye(4, dtype=np.float64)
    T[0, 0] = 1.0 - (tyy + tzz); T[0, 1] = txy - twz; T[0, 2] = txz + twy
    T[1, 0] = txy + twz; T[1, 1] = 1.0 - (txx + tzz); T[1, 2] = tyz - twx
    T[2, 0] = txz - twy; T[2, 1] = tyz + twx; T[2, 2] = 1.0 - (txx + tyy)
    T[0, 3], T[1, 3], T[2, 3] = (float(v) for v in t)
    return T


# ---------------------------------------------------------------------------------------
# cfg1: single 10k-point cloud, res 0.15, max_range 5
# ---------------------------------------------------------------------------------------
def cfg1_pointcloud(n=10000, seed=0):
    rng = np.random.default_rng(seed)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    r = rng.uniform(0.5, 8.0, size=n)
    pts = (d * r[:, None]).astype(np.float32)
    q = np.array([1.0, 0.0, 0.0, 0.0])
    t = np.zeros(3)
    return {"points": pts, "q": q, "t": t, "T": quat_to_matrix(q, t),
            "params": {"resolution": 0.15, "sensor_max_range": 5.0}}


# ---------------------------------------------------------------------------------------
# cfg2: 640x480 disparity image + Q (generateQ layout, world_base.cc:146-180)
# ---------------------------------------------------------------------------------------
def generate_q(Tx, left_cx, left_cy, left_fx, left_fy, right_cx):
    Q = np.zeros((4, 4), dtype=np.float64)
    Q[0, 0] = left_fy * Tx
    Q[0, 3] = -left_fy * left_cx * Tx
    Q[1, 1] = left_fx * Tx
    Q[1, 3] = -left_fx * left_cy * Tx
    Q[2, 3] = left_fx * left_fy * Tx
    Q[3, 2] = -left_fy
    Q[3, 3] = left_fy * (left_cx - right_cx)
    return Q


def cfg2_disparity(rows=480, cols=640, seed=1):
    rng = np.random.default_rng(seed)
    disp = rng.uniform(0.5, 128.0, size=(rows, cols)).astype(np.float32)
    sel = rng.uniform(size=(rows, cols))
    disp[sel < 0.10] = -1.0          # invalid marker == image minimum
    disp[(sel >= 0.10) & (sel < 0.11)] = 0.0   # zero disparity -> point at infinity
    Q = generate_q(Tx=-0.110074, left_cx=367.215, left_cy=248.375, left_fx=458.654,
                   left_fy=458.654, right_cx=366.0)
    # camera looking along +x of the world: optical (z fwd, x right, y down) -> world
    q = np.array([0.5, -0.5, 0.5, -0.5])
    t = np.array([0.3, -0.2, 1.1])
    return {"disparity": disp, "Q": Q, "full_image_size": (float(cols), float(rows)), "q": q,
            "t": t, "T": quat_to_matrix(q, t),
            "params": {"resolution": 0.15, "sensor_max_range": 5.0}}


# ---------------------------------------------------------------------------------------
# cfg3 / cfg4: HDL-64-style scans inside an axis-aligned walled box with a ground plane
# ---------------------------------------------------------------------------------------
HDL64_ELEVATIONS = 64
HDL64_AZIMUTHS = 2083


def hdl64_scan(position=(0.0, 0.0, 1.8), yaw_rad=0.0, box=(-40.0, 40.0, -40.0, 40.0), seed=0,
               n_elev=HDL64_ELEVATIONS, n_azim=HDL64_AZIMUTHS, sigma=0.02):
    """One revolution: n_azim firings x n_elev lasers (azimuth-major point order).

    Returns points in the SENSOR frame (float32) and the sensor->world transform.
    The world is a ground plane z=0 plus four infinitely tall walls at box=(xmin,xmax,ymin,ymax).
    """
    rng = np.random.default_rng(seed)
    elev = np.deg2rad(np.linspace(-24.8, 2.0, n_elev))
    azim = np.linspace(0.0, 2.0 * np.pi, n_azim, endpoint=False)
    az, el = np.meshgrid(azim, elev, indexing="ij")      # (n_azim, n_elev)
    az = az.reshape(-1)
    el = el.reshape(-1)
    d_s = np.stack([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)], axis=1)
    c, s = math.cos(yaw_rad), math.sin(yaw_rad)
    d_w = np.stack([c * d_s[:, 0] - s * d_s[:, 1], s * d_s[:, 0] + c * d_s[:, 1], d_s[:, 2]], axis=1)
    px, py, pz = position
    xmin, xmax, ymin, ymax = box
    with np.errstate(divide="ignore", invalid="ignore"):
        tx = np.where(d_w[:, 0] > 0, (xmax - px) / d_w[:, 0],
                      np.where(d_w[:, 0] < 0, (xmin - px) / d_w[:, 0], np.inf))
        ty = np.where(d_w[:, 1] > 0, (ymax - py) / d_w[:, 1],
                      np.where(d_w[:, 1] < 0, (ymin - py) / d_w[:, 1], np.inf))
        tz = np.where(d_w[:, 2] < 0, (0.0 - pz) / d_w[:, 2], np.inf)
    rng_true = np.minimum(np.minimum(tx, ty), tz)
    rng_meas = rng_true + rng.normal(0.0, sigma, size=rng_true.shape)
    pts = (d_s * rng_meas[:, None]).astype(np.float32)
    q = yaw_quat(yaw_rad)
    t = np.array(position, dtype=np.float64)
    return {"points": pts, "q": q, "t": t, "T": quat_to_matrix(q, t)}


def cfg3_scan(seed=0):
    s = hdl64_scan(seed=seed)
    s["params"] = {"resolution": 0.10, "sensor_max_range": 50.0}
    return s


def cfg4_scan(index, resolution=0.05):
    """Scan `index` of the 1000-scan corridor trajectory (0.5 m/scan along +x, yaw 0.2 deg/scan)."""
    s = hdl64_scan(position=(0.5 * index, 0.0, 1.8), yaw_rad=math.radians(0.2 * index),
                   box=(-50.0, 550.0, -40.0, 40.0), seed=index)
    s["params"] = {"resolution": resolution, "sensor_max_range": 50.0}
    return s


def trajectory_scan(index, resolution=0.10):
    """cfg3-style scan `index` from a moving sensor (bench.py steps; same scene as cfg4)."""
    return cfg4_scan(index, resolution=resolution)


# ---------------------------------------------------------------------------------------
# cfg5: batched queries
# ---------------------------------------------------------------------------------------
def cfg5_queries(n_segments=1_000_000, n_points=1_000_000, seed=2,
                 bounds=((-45.0, 45.0), (-45.0, 45.0), (-0.5, 4.0))):
    rng = np.random.default_rng(seed)
    lo = np.array([b[0] for b in bounds])
    hi = np.array([b[1] for b in bounds])
    a = rng.uniform(lo, hi, size=(n_segments, 3))
    d = rng.normal(size=(n_segments, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    length = rng.uniform(0.5, 20.0, size=n_segments)
    b = a + d * length[:, None]
    segments = np.concatenate([a, b], axis=1).astype(np.float64)
    points = rng.uniform(lo, hi, size=(n_points, 3)).astype(np.float64)
    return {"segments": segments, "points": points}
