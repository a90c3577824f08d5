"""ctypes binding of the CPU oracle (oracle/oracle.cc).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")


def build(force=False):
    """Compile liboracle.so with the committed Makefile (g++ only)."""
    src = os.path.join(_HERE, "oracle.cc")
    hdr = os.path.join(_HERE, "oracle.h")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"],
                          stdout=subprocess.DEVNULL)
    return _LIB_PATH


class Params(C.Structure):
    """Mirror of oracle_params == OctomapParameters (octomap_world.h:54-109)."""
    _fields_ = [
        ("resolution", C.c_double),
        ("probability_hit", C.c_double),
        ("probability_miss", C.c_double),
        ("threshold_min", C.c_double),
        ("threshold_max", C.c_double),
        ("threshold_occupancy", C.c_double),
        ("filter_speckles", C.c_int32),
        ("_pad0", C.c_int32),
        ("max_free_space", C.c_double),
        ("min_height_free_space", C.c_double),
        ("sensor_max_range", C.c_double),
        ("visualize_min_z", C.c_double),
        ("visualize_max_z", C.c_double),
        ("treat_unknown_as_occupied", C.c_int32),
        ("change_detection_enabled", C.c_int32),
    ]


class ScanStats(C.Structure):
    _fields_ = [
        ("points_in", C.c_uint64),
        ("points_valid", C.c_uint64),
        ("points_in_range", C.c_uint64),
        ("rays_cast", C.c_uint64),
        ("traversals", C.c_uint64),
        ("occupied_emitted", C.c_uint64),
        ("unique_free", C.c_uint64),
        ("unique_occupied", C.c_uint64),
        ("longest_ray", C.c_uint64),
        ("t_keyset_s", C.c_double),
        ("t_update_s", C.c_double),
        ("t_inner_s", C.c_double),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        build()
    L = C.CDLL(_LIB_PATH)
    vp = C.c_void_p
    dp = C.POINTER(C.c_double)
    L.oracle_default_params.argtypes = [C.POINTER(Params)]
    L.oracle_create.argtypes = [C.POINTER(Params)]
    L.oracle_create.restype = vp
    L.oracle_destroy.argtypes = [vp]
    L.oracle_reset.argtypes = [vp]
    L.oracle_set_params.argtypes = [vp, C.POINTER(Params)]
    L.oracle_get_params.argtypes = [vp, C.POINTER(Params)]
    L.oracle_set_faithful_inner_update.argtypes = [vp, C.c_int]
    L.oracle_insert_pointcloud.argtypes = [vp, vp, C.c_size_t, C.c_size_t, C.c_int, dp,
                                           C.POINTER(C.c_size_t)]
    L.oracle_insert_projected.argtypes = [vp, vp, C.c_int, C.c_int, C.c_size_t, dp, dp]
    L.oracle_insert_disparity.argtypes = [vp, vp, C.c_int, C.c_int, C.c_size_t, dp, C.c_double,
                                          dp, dp]
    L.oracle_reproject.argtypes = [vp, C.c_int, C.c_int, C.c_size_t, dp, vp]
    L.oracle_fixup_q.argtypes = [dp, C.c_double, C.c_int, dp]
    L.oracle_apply_keys.argtypes = [vp, vp, C.c_size_t, vp, C.c_size_t]
    L.oracle_query_points.argtypes = [vp, vp, C.c_size_t, vp]
    L.oracle_query_lines.argtypes = [vp, vp, C.c_size_t, vp]
    L.oracle_query_points_probability.argtypes = [vp, vp, C.c_size_t, vp, vp]
    L.oracle_query_visibility.argtypes = [vp, vp, C.c_size_t, C.c_int, vp]
    L.oracle_query_lines_bbox.argtypes = [vp, vp, C.c_size_t, dp, vp]
    L.oracle_query_bbox_status.argtypes = [vp, vp, C.c_size_t, dp, vp]
    L.oracle_check_path_collisions.argtypes = [vp, vp, C.c_size_t, dp, C.POINTER(C.c_size_t)]
    L.oracle_check_path_collisions.restype = C.c_int
    L.oracle_coord_to_key_checked.argtypes = [vp, C.c_double, C.POINTER(C.c_uint16)]
    L.oracle_coord_to_key_checked.restype = C.c_int
    L.oracle_coord_to_key.argtypes = [vp, C.c_double]
    L.oracle_coord_to_key.restype = C.c_uint16
    L.oracle_key_to_coord.argtypes = [vp, C.c_uint16]
    L.oracle_key_to_coord.restype = C.c_double
    L.oracle_compute_ray_keys.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float), vp,
                                          C.c_size_t]
    L.oracle_compute_ray_keys.restype = C.c_long
    L.oracle_quat_to_matrix.argtypes = [dp, dp, dp]
    L.oracle_logodds_constants.argtypes = [vp, C.POINTER(C.c_float)]
    for name in ("oracle_last_free_count", "oracle_last_occupied_count", "oracle_leaf_count",
                 "oracle_tree_node_count"):
        getattr(L, name).argtypes = [vp]
        getattr(L, name).restype = C.c_size_t
    L.oracle_last_free_keys.argtypes = [vp, vp]
    L.oracle_last_occupied_keys.argtypes = [vp, vp]
    L.oracle_last_scan_stats.argtypes = [vp, C.POINTER(ScanStats)]
    L.oracle_export_leaves.argtypes = [vp, vp, vp]
    L.oracle_prune.argtypes = [vp]
    L.oracle_get_changed.argtypes = [vp, vp, vp, C.c_size_t]
    L.oracle_get_changed.restype = C.c_size_t
    L.oracle_set_log_odds_bbox.argtypes = [vp, vp, C.c_size_t, dp, C.c_double, C.c_int]
    L.oracle_write_bt.argtypes = [vp, C.c_int, vp, C.c_size_t]
    L.oracle_write_bt.restype = C.c_size_t
    L.oracle_write_ot.argtypes = [vp, C.c_int, vp, C.c_size_t]
    L.oracle_write_ot.restype = C.c_size_t
    _lib = L
    return L


def default_params(**overrides):
    p = Params()
    lib().oracle_default_params(C.byref(p))
    for k, v in overrides.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def quat_to_matrix(q_wxyz, t):
    q = np.ascontiguousarray(q_wxyz, dtype=np.float64)
    tt = np.ascontiguousarray(t, dtype=np.float64)
    T = np.zeros(16, dtype=np.float64)
    lib().oracle_quat_to_matrix(_dptr(q), _dptr(tt), _dptr(T))
    return T.reshape(4, 4)


def pack_keys(k3):
    """(n,3) uint16 -> uint64 k0 | k1<<16 | k2<<32 (sortable, for exact set comparison)."""
    k = np.asarray(k3, dtype=np.uint64).reshape(-1, 3)
    return k[:, 0] | (k[:, 1] << np.uint64(16)) | (k[:, 2] << np.uint64(32))


class OracleWorld:
    """The reference's OctomapWorld, restated on the CPU (same method names)."""

    def __init__(self, params=None, **overrides):
        self._L = lib()
        if params is None:
            params = default_params(**overrides)
        self._h = self._L.oracle_create(C.byref(params))

    def close(self):
        if self._h:
            self._L.oracle_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- map management -----------------------------------------------------------------
    def resetMap(self):
        self._L.oracle_reset(self._h)

    def setOctomapParameters(self, params):
        self._L.oracle_set_params(self._h, C.byref(params))

    def getOctomapParameters(self):
        p = Params()
        self._L.oracle_get_params(self._h, C.byref(p))
        return p

    def set_faithful_inner_update(self, on):
        self._L.oracle_set_faithful_inner_update(self._h, int(bool(on)))

    # -- insertion ----------------------------------------------------------------------
    def insertPointcloud(self, T, xyz, is_dense=True):
        """T: 4x4 (getTransformationMatrix).  Returns the mutated cloud (world frame)."""
        cloud = np.array(xyz, dtype=np.float32, order="C").reshape(-1, 3)
        Tm = np.ascontiguousarray(T, dtype=np.float64).reshape(16)
        n_out = C.c_size_t(0)
        self._L.oracle_insert_pointcloud(self._h, cloud.ctypes.data, cloud.shape[0], 12,
                                         int(bool(is_dense)), _dptr(Tm), C.byref(n_out))
        return cloud[: n_out.value]

    def insertProjectedDisparity(self, q_wxyz, t, xyz_hw3):
        img = np.ascontiguousarray(xyz_hw3, dtype=np.float32)
        rows, cols = img.shape[:2]
        q = np.ascontiguousarray(q_wxyz, dtype=np.float64)
        tt = np.ascontiguousarray(t, dtype=np.float64)
        self._L.oracle_insert_projected(self._h, img.ctypes.data, rows, cols, cols * 12,
                                        _dptr(q), _dptr(tt))

    def insertDisparityImage(self, q_wxyz, t, disparity, Q_full, full_image_width):
        d = np.ascontiguousarray(disparity, dtype=np.float32)
        rows, cols = d.shape
        Q = np.ascontiguousarray(Q_full, dtype=np.float64).reshape(16)
        q = np.ascontiguousarray(q_wxyz, dtype=np.float64)
        tt = np.ascontiguousarray(t, dtype=np.float64)
        self._L.oracle_insert_disparity(self._h, d.ctypes.data, rows, cols, cols * 4, _dptr(Q),
                                        float(full_image_width), _dptr(q), _dptr(tt))

    def updateOccupancy(self, free_k3, occ_k3):
        f = np.ascontiguousarray(free_k3, dtype=np.uint16).reshape(-1, 3)
        o = np.ascontiguousarray(occ_k3, dtype=np.uint16).reshape(-1, 3)
        self._L.oracle_apply_keys(self._h, f.ctypes.data, f.shape[0], o.ctypes.data, o.shape[0])

    def getChangedPoints(self):
        n = self._L.oracle_get_changed(self._h, None, None, 0)
        k = np.zeros((n, 3), dtype=np.uint16)
        o = np.zeros(n, dtype=np.uint8)
        if n:
            self._L.oracle_get_changed(self._h, k.ctypes.data, o.ctypes.data, n)
        return k, o.astype(bool)

    def setLogOddsBoundingBox(self, positions, bounding_box_size, log_odds_value, insertion_method=0):
        p = np.ascontiguousarray(positions, dtype=np.float64).reshape(-1, 3)
        b = np.ascontiguousarray(bounding_box_size, dtype=np.float64).reshape(3)
        self._L.oracle_set_log_odds_bbox(self._h, p.ctypes.data, p.shape[0], _dptr(b),
                                         float(log_odds_value), int(insertion_method))

    def setFree(self, positions, bounding_box_size, insertion_method=0):
        self.setLogOddsBoundingBox(positions, bounding_box_size, float(self.logodds_constants()[2]),
                                   insertion_method)

    def setOccupied(self, positions, bounding_box_size, insertion_method=0):
        self.setLogOddsBoundingBox(positions, bounding_box_size, float(self.logodds_constants()[3]),
                                   insertion_method)

    # -- queries ------------------------------------------------------------------------
    def getCellStatusPoint(self, xyz):
        p = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
        out = np.zeros(p.shape[0], dtype=np.uint8)
        self._L.oracle_query_points(self._h, p.ctypes.data, p.shape[0], out.ctypes.data)
        return out

    def getLineStatus(self, start_end_xyz6):
        p = np.ascontiguousarray(start_end_xyz6, dtype=np.float64).reshape(-1, 6)
        out = np.zeros(p.shape[0], dtype=np.uint8)
        self._L.oracle_query_lines(self._h, p.ctypes.data, p.shape[0], out.ctypes.data)
        return out

    def getCellProbabilityPoint(self, xyz):
        p = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
        st = np.zeros(p.shape[0], dtype=np.uint8)
        pr = np.zeros(p.shape[0], dtype=np.float64)
        self._L.oracle_query_points_probability(self._h, p.ctypes.data, p.shape[0], st.ctypes.data,
                                                pr.ctypes.data)
        return st, pr

    def getVisibility(self, view_voxel_xyz6, stop_at_unknown_cell):
        p = np.ascontiguousarray(view_voxel_xyz6, dtype=np.float64).reshape(-1, 6)
        st = np.zeros(p.shape[0], dtype=np.uint8)
        self._L.oracle_query_visibility(self._h, p.ctypes.data, p.shape[0],
                                        int(bool(stop_at_unknown_cell)), st.ctypes.data)
        return st

    def getLineStatusBoundingBox(self, start_end_xyz6, bounding_box_size):
        p = np.ascontiguousarray(start_end_xyz6, dtype=np.float64).reshape(-1, 6)
        b = np.ascontiguousarray(bounding_box_size, dtype=np.float64).reshape(3)
        st = np.zeros(p.shape[0], dtype=np.uint8)
        self._L.oracle_query_lines_bbox(self._h, p.ctypes.data, p.shape[0], _dptr(b), st.ctypes.data)
        return st

    def getCellStatusBoundingBox(self, xyz, bounding_box_size):
        p = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
        b = np.ascontiguousarray(bounding_box_size, dtype=np.float64).reshape(3)
        st = np.zeros(p.shape[0], dtype=np.uint8)
        self._L.oracle_query_bbox_status(self._h, p.ctypes.data, p.shape[0], _dptr(b), st.ctypes.data)
        return st

    def checkPathForCollisionsWithRobot(self, xyz, robot_size):
        """(collides, index of the earliest colliding pose or None)"""
        p = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
        b = np.ascontiguousarray(robot_size, dtype=np.float64).reshape(3)
        idx = C.c_size_t(0)
        hit = self._L.oracle_check_path_collisions(self._h, p.ctypes.data, p.shape[0], _dptr(b), C.byref(idx))
        return bool(hit), (int(idx.value) if hit else None)

    # -- primitives ---------------------------------------------------------------------
    def coordToKey(self, c):
        return int(self._L.oracle_coord_to_key(self._h, float(c)))

    def coordToKeyChecked(self, c):
        k = C.c_uint16(0)
        ok = self._L.oracle_coord_to_key_checked(self._h, float(c), C.byref(k))
        return (bool(ok), int(k.value))

    def keyToCoord(self, k):
        return float(self._L.oracle_key_to_coord(self._h, int(k)))

    def computeRayKeys(self, origin, end, cap=100000):
        o = (C.c_float * 3)(*[float(np.float32(v)) for v in origin])
        e = (C.c_float * 3)(*[float(np.float32(v)) for v in end])
        buf = np.zeros((cap, 3), dtype=np.uint16)
        n = self._L.oracle_compute_ray_keys(self._h, o, e, buf.ctypes.data, cap)
        if n < 0:
            return None
        return buf[:n].copy()

    def logodds_constants(self):
        out = (C.c_float * 5)()
        self._L.oracle_logodds_constants(self._h, out)
        return np.array(list(out), dtype=np.float32)

    # -- parity hooks -------------------------------------------------------------------
    def last_scan_keys(self):
        nf = self._L.oracle_last_free_count(self._h)
        no = self._L.oracle_last_occupied_count(self._h)
        f = np.zeros((nf, 3), dtype=np.uint16)
        o = np.zeros((no, 3), dtype=np.uint16)
        if nf:
            self._L.oracle_last_free_keys(self._h, f.ctypes.data)
        if no:
            self._L.oracle_last_occupied_keys(self._h, o.ctypes.data)
        return f, o

    def last_scan_stats(self):
        s = ScanStats()
        self._L.oracle_last_scan_stats(self._h, C.byref(s))
        return s.as_dict()

    def export_leaves(self):
        n = self._L.oracle_leaf_count(self._h)
        k = np.zeros((n, 3), dtype=np.uint16)
        v = np.zeros(n, dtype=np.float32)
        if n:
            self._L.oracle_export_leaves(self._h, k.ctypes.data, v.ctypes.data)
        return k, v

    def tree_node_count(self):
        return int(self._L.oracle_tree_node_count(self._h))

    def prune(self):
        self._L.oracle_prune(self._h)

    def write_ot(self, with_header=False):
        n = self._L.oracle_write_ot(self._h, int(bool(with_header)), None, 0)
        buf = (C.c_ubyte * max(1, n))()
        self._L.oracle_write_ot(self._h, int(bool(with_header)), buf, n)
        return bytes(buf[:n])

    def write_bt(self, live=False):
        n = self._L.oracle_write_bt(self._h, int(bool(live)), None, 0) if not live else None
        if live:
            buf = (C.c_ubyte * (1 << 26))()
            n = self._L.oracle_write_bt(self._h, 1, buf, len(buf))
            return bytes(buf[:n])
        buf = (C.c_ubyte * max(1, n))()
        self._L.oracle_write_bt(self._h, 0, buf, n)
        return bytes(buf[:n])


def reproject(disparity, Q):
    d = np.ascontiguousarray(disparity, dtype=np.float32)
    rows, cols = d.shape
    Qm = np.ascontiguousarray(Q, dtype=np.float64).reshape(16)
    out = np.zeros((rows, cols, 3), dtype=np.float32)
    lib().oracle_reproject(d.ctypes.data, rows, cols, cols * 4, _dptr(Qm), out.ctypes.data)
    return out


def fixup_q(Q_full, full_image_width, disparity_cols):
    Qm = np.ascontiguousarray(Q_full, dtype=np.float64).reshape(16)
    out = np.zeros(16, dtype=np.float64)
    lib().oracle_fixup_q(_dptr(Qm), float(full_image_width), int(disparity_cols), _dptr(out))
    return out.reshape(4, 4)
