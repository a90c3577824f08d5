"""Host-side mirror of volumetric_mapping::OctomapWorld for the scan-insertion path.

Method names and argument meaning follow the reference class
(octomap_world/include/octomap_world/octomap_world.h:119-256 and
volumetric_map_base/include/volumetric_map_base/world_base.h:50-235); every call goes through
the C ABI of include/vmap/vmap.h into the CUDA kernels.  numpy arrays stand in for
pcl::PointCloud / cv::Mat / Eigen types, a (q_wxyz, t) pair or a 4x4 matrix for
kindr::minimal::QuatTransformation.
"""
import ctypes as C
import os

import numpy as np

from . import _capi
from ._capi import VmapError

DEDUPE_BITMASK = 0
DEDUPE_SORT = 1

# CellStatus (world_base.h:54)
kFree, kOccupied, kUnknown = 0, 1, 2


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def quat_to_matrix(q_wxyz, t):
    """QuatTransformation::getTransformationMatrix(): Eigen::Quaterniond::toRotationMatrix()
    (same association order) with the translation in the last column."""
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


class OctomapWorld:
    """GPU-resident OctomapWorld (insertion + point/line queries)."""

    def __init__(self, params=None, device=-1, dedupe_mode=DEDUPE_BITMASK, initial_blocks=0,
                 max_blocks=0, dense_window_bytes=0, **param_overrides):
        self._L = _capi.lib()
        if params is None:
            params = _capi.default_params(**param_overrides)
        elif param_overrides:
            raise TypeError("pass either params or keyword overrides")
        cfg = _capi.default_config(device=device, dedupe_mode=dedupe_mode,
                                   initial_blocks=initial_blocks, max_blocks=max_blocks,
                                   dense_window_bytes=dense_window_bytes)
        h = C.c_void_p()
        st = self._L.vmap_create(C.byref(params), C.byref(cfg), C.byref(h))
        if st != 0:
            raise VmapError(st, (self._L.vmap_last_error(None) or b"").decode())
        self._h = h

    # -- plumbing -----------------------------------------------------------------------
    def _check(self, st):
        if st != 0:
            raise VmapError(st, (self._L.vmap_last_error(self._h) or b"").decode())

    def close(self):
        if getattr(self, "_h", None):
            self._L.vmap_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    # -- map management (octomap_world.cc:75-108) -----------------------------------------
    def resetMap(self):
        self._check(self._L.vmap_reset(self._h))

    def setOctomapParameters(self, params):
        self._check(self._L.vmap_set_params(self._h, C.byref(params)))

    def getOctomapParameters(self):
        p = _capi.Params()
        self._check(self._L.vmap_get_params(self._h, C.byref(p)))
        return p

    # -- insertion ----------------------------------------------------------------------
    def insertPointcloud(self, T, xyz, is_dense=True):
        """WorldBase::insertPointcloud -> insertPointcloudIntoMapImpl (octomap_world.cc:110-141).
        T: 4x4 getTransformationMatrix().  Returns the cloud as the reference leaves it in the
        caller's buffer (NaN-compacted when !is_dense, world frame)."""
        cloud = np.array(xyz, dtype=np.float32, order="C").reshape(-1, 3)
        Tm = np.ascontiguousarray(T, dtype=np.float64).reshape(16)
        n_out = C.c_size_t(0)
        self._check(self._L.vmap_insert_pointcloud(self._h, cloud.ctypes.data, cloud.shape[0], 12,
                                                   int(bool(is_dense)), _dptr(Tm),
                                                   cloud.ctypes.data, C.byref(n_out)))
        return cloud[: n_out.value]

    def insertPointcloudFast(self, T, xyz):
        """Same insertion without the in-place write-back of the transformed cloud."""
        cloud = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        Tm = np.ascontiguousarray(T, dtype=np.float64).reshape(16)
        self._check(self._L.vmap_insert_pointcloud(self._h, cloud.ctypes.data, cloud.shape[0], 12,
                                                   1, _dptr(Tm), None, None))

    def insertPointcloudHostPtr(self, T_c16, host_ptr, n, stride_bytes=12):
        """Same call with the arguments already marshalled (a (c_double*16) transform and the
        address of n packed float32 points, e.g. pinned memory): what a C++ caller pays."""
        self._check(self._L.vmap_insert_pointcloud(self._h, C.c_void_p(int(host_ptr)), int(n),
                                                   int(stride_bytes), 1, T_c16, None, None))

    def insertPointcloudDevice(self, T, data_ptr, n, stride_bytes=12):
        """Points already in device memory (e.g. a torch tensor's data_ptr())."""
        Tm = np.ascontiguousarray(T, dtype=np.float64).reshape(16)
        self._check(self._L.vmap_insert_pointcloud_dev(self._h, C.c_void_p(int(data_ptr)), int(n),
                                                       int(stride_bytes), _dptr(Tm)))

    # -- the scan pipeline (include/vmap/vmap.h, "the scan pipeline") ----------------------------
    def flush(self):
        """Drain the pipelined scans; deferred failures surface here."""
        self._check(self._L.vmap_flush(self._h))

    def set_pipeline_depth(self, depth):
        self._check(self._L.vmap_set_pipeline_depth(self._h, int(depth)))

    def totals(self):
        """Sums of the per-scan counters since creation / resetMap, plus `scans` and `replayed`."""
        s = _capi.ScanStats()
        n, r = C.c_uint64(0), C.c_uint64(0)
        self._check(self._L.vmap_totals(self._h, C.byref(s), C.byref(n), C.byref(r)))
        d = s.as_dict()
        d["scans"], d["replayed"] = int(n.value), int(r.value)
        return d

    def region_begin(self):
        self._check(self._L.vmap_region_begin(self._h))

    def region_end(self):
        """Milliseconds on the device since region_begin (everything drained on both sides)."""
        ms = C.c_double(0.0)
        self._check(self._L.vmap_region_end(self._h, C.byref(ms)))
        return float(ms.value)

    def insertProjectedDisparity(self, q_wxyz, t, xyz_hw3):
        """insertProjectedDisparityIntoMapImpl (octomap_world.cc:143-175)."""
        img = np.ascontiguousarray(xyz_hw3, dtype=np.float32)
        rows, cols = img.shape[:2]
        q = np.ascontiguousarray(q_wxyz, dtype=np.float64)
        tt = np.ascontiguousarray(t, dtype=np.float64)
        self._check(self._L.vmap_insert_projected(self._h, img.ctypes.data, rows, cols, cols * 12,
                                                  _dptr(q), _dptr(tt)))

    def insertDisparityImage(self, q_wxyz, t, disparity, Q_full, full_image_width):
        """WorldBase::insertDisparityImage(cv::Mat) (world_base.cc:51-87)."""
        d = np.ascontiguousarray(disparity, dtype=np.float32)
        rows, cols = d.shape
        Q = np.ascontiguousarray(Q_full, dtype=np.float64).reshape(16)
        q = np.ascontiguousarray(q_wxyz, dtype=np.float64)
        tt = np.ascontiguousarray(t, dtype=np.float64)
        self._check(self._L.vmap_insert_disparity(self._h, d.ctypes.data, rows, cols, cols * 4,
                                                  _dptr(Q), float(full_image_width), _dptr(q),
                                                  _dptr(tt)))

    def insertDisparityImageDevice(self, q_wxyz, t, data_ptr, rows, cols, Q_full, full_image_width,
                                   row_pitch_bytes=None):
        Q = np.ascontiguousarray(Q_full, dtype=np.float64).reshape(16)
        q = np.ascontiguousarray(q_wxyz, dtype=np.float64)
        tt = np.ascontiguousarray(t, dtype=np.float64)
        pitch = cols * 4 if row_pitch_bytes is None else int(row_pitch_bytes)
        self._check(self._L.vmap_insert_disparity_dev(self._h, C.c_void_p(int(data_ptr)), rows,
                                                      cols, pitch, _dptr(Q),
                                                      float(full_image_width), _dptr(q), _dptr(tt)))

    def reprojectImageTo3D(self, disparity, Q):
        d = np.ascontiguousarray(disparity, dtype=np.float32)
        rows, cols = d.shape
        Qm = np.ascontiguousarray(Q, dtype=np.float64).reshape(16)
        out = np.zeros((rows, cols, 3), dtype=np.float32)
        self._check(self._L.vmap_reproject(self._h, d.ctypes.data, rows, cols, cols * 4, _dptr(Qm),
                                           out.ctypes.data))
        return out

    def updateOccupancy(self, free_k3, occ_k3):
        """updateOccupancy(KeySet*, KeySet*) (octomap_world.cc:238-264)."""
        f = np.ascontiguousarray(free_k3, dtype=np.uint16).reshape(-1, 3)
        o = np.ascontiguousarray(occ_k3, dtype=np.uint16).reshape(-1, 3)
        self._check(self._L.vmap_apply_keys(self._h, f.ctypes.data, f.shape[0], o.ctypes.data,
                                            o.shape[0]))

    def getChangedPoints(self):
        """(keys, occupied) collected by change detection since the last call; resets it
        (octomap_world.cc:1413-1440).  Points are keyToCoord(keys)."""
        n = C.c_size_t(0)
        st = self._L.vmap_get_changed(self._h, None, None, 0, C.byref(n))
        if st not in (0, 5):
            self._check(st)
        k = np.zeros((n.value, 3), dtype=np.uint16)
        o = np.zeros(n.value, dtype=np.uint8)
        if n.value:
            self._check(self._L.vmap_get_changed(self._h, k.ctypes.data, o.ctypes.data, n.value, C.byref(n)))
        return k, o.astype(bool)

    # -- manual edits (octomap_world.cc:497-523, 781-818) --------------------------------
    def setLogOddsBoundingBox(self, positions, bounding_box_size, log_odds_value, insertion_method=0):
        p = np.ascontiguousarray(positions, dtype=np.float64).reshape(-1, 3)
        b = np.ascontiguousarray(bounding_box_size, dtype=np.float64).reshape(3)
        self._check(self._L.vmap_set_log_odds_bbox(self._h, p.ctypes.data, p.shape[0], _dptr(b),
                                                   float(log_odds_value), int(insertion_method)))

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
        self._check(self._L.vmap_query_points(self._h, p.ctypes.data, p.shape[0], out.ctypes.data))
        return out

    def getLineStatus(self, start_end_xyz6):
        p = np.ascontiguousarray(start_end_xyz6, dtype=np.float64).reshape(-1, 6)
        out = np.zeros(p.shape[0], dtype=np.uint8)
        self._check(self._L.vmap_query_lines(self._h, p.ctypes.data, p.shape[0], out.ctypes.data))
        return out

    def getCellProbabilityPoint(self, xyz):
        """(true status, probability) per point (octomap_world.cc:363-393)."""
        p = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
        st = np.zeros(p.shape[0], dtype=np.uint8)
        pr = np.zeros(p.shape[0], dtype=np.float64)
        self._check(self._L.vmap_query_points_probability(self._h, p.ctypes.data, p.shape[0],
                                                          st.ctypes.data, pr.ctypes.data))
        return st, pr

    def getCellTrueStatusPoint(self, xyz):
        p = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
        st = np.zeros(p.shape[0], dtype=np.uint8)
        self._check(self._L.vmap_query_points_probability(self._h, p.ctypes.data, p.shape[0],
                                                          st.ctypes.data, None))
        return st

    def getVisibility(self, view_voxel_xyz6, stop_at_unknown_cell):
        p = np.ascontiguousarray(view_voxel_xyz6, dtype=np.float64).reshape(-1, 6)
        st = np.zeros(p.shape[0], dtype=np.uint8)
        self._check(self._L.vmap_query_visibility(self._h, p.ctypes.data, p.shape[0],
                                                  int(bool(stop_at_unknown_cell)), st.ctypes.data))
        return st

    def getLineStatusBoundingBox(self, start_end_xyz6, bounding_box_size):
        p = np.ascontiguousarray(start_end_xyz6, dtype=np.float64).reshape(-1, 6)
        b = np.ascontiguousarray(bounding_box_size, dtype=np.float64).reshape(3)
        st = np.zeros(p.shape[0], dtype=np.uint8)
        self._check(self._L.vmap_query_lines_bbox(self._h, p.ctypes.data, p.shape[0], _dptr(b),
                                                  st.ctypes.data))
        return st

    def stagePointcloudHostPtr(self, host_ptr, n, stride_bytes=12):
        """Start the asynchronous host->device copy of a scan (see vmap_stage_pointcloud)."""
        self._check(self._L.vmap_stage_pointcloud(self._h, C.c_void_p(int(host_ptr)), int(n), int(stride_bytes)))

    def insertStaged(self, T, is_dense=True):
        Tm = np.ascontiguousarray(T, dtype=np.float64).reshape(16)
        self._check(self._L.vmap_insert_staged(self._h, int(bool(is_dense)), _dptr(Tm)))

    def loadOccupiedPointcloudFile(self, filename):
        """loadOctomapCallback's .pcd / .ply branch (octomap_manager.cc:313-344); returns the point count."""
        n = C.c_size_t(0)
        self._check(self._L.vmap_load_occupied_pointcloud_file(self._h, os.fsencode(filename), C.byref(n)))
        return int(n.value)

    def insertOccupiedPoints(self, xyz):
        p = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        self._check(self._L.vmap_insert_occupied_points(self._h, p.ctypes.data, p.shape[0], 12))

    def getCellStatusBoundingBox(self, xyz, bounding_box_size):
        """octomap_world.cc:274-344, batched: one box size, many centres."""
        p = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
        b = np.ascontiguousarray(bounding_box_size, dtype=np.float64).reshape(3)
        st = np.zeros(p.shape[0], dtype=np.uint8)
        self._check(self._L.vmap_query_bbox_status(self._h, p.ctypes.data, p.shape[0], _dptr(b), st.ctypes.data))
        return st

    def setRobotSize(self, robot_size):
        self._robot_size = np.ascontiguousarray(robot_size, dtype=np.float64).reshape(3).copy()

    def getRobotSize(self):
        return getattr(self, "_robot_size", np.ones(3)).copy()

    def checkCollisionWithRobot(self, robot_position):
        return self.checkPathForCollisionsWithRobot(np.asarray(robot_position, np.float64).reshape(1, 3))[0]

    def checkPathForCollisionsWithRobot(self, robot_positions):
        """octomap_world.cc:1384-1400: (collides, index of the earliest colliding pose or None)."""
        p = np.ascontiguousarray(robot_positions, dtype=np.float64).reshape(-1, 3)
        b = self.getRobotSize()
        hit, idx = C.c_int(0), C.c_size_t(0)
        self._check(self._L.vmap_check_path_collisions(self._h, p.ctypes.data, p.shape[0], _dptr(b), C.byref(hit),
                                                       C.byref(idx)))
        return bool(hit.value), (int(idx.value) if hit.value else None)

    def getCellStatusPointDevice(self, xyz_ptr, n, status_ptr):
        self._check(self._L.vmap_query_points_dev(self._h, C.c_void_p(int(xyz_ptr)), int(n),
                                                  C.c_void_p(int(status_ptr))))

    def getLineStatusDevice(self, ab_ptr, n, status_ptr):
        self._check(self._L.vmap_query_lines_dev(self._h, C.c_void_p(int(ab_ptr)), int(n),
                                                 C.c_void_p(int(status_ptr))))

    # -- multi-GPU sharding (include/vmap/vmap.h, "Multi-GPU sharding") ------------------
    def set_shard(self, rank, world):
        self._check(self._L.vmap_set_shard(self._h, int(rank), int(world)))
        self._world = int(world)

    def shard_generate(self, T, xyz, is_dense=True):
        """Generate one scan here (nothing is applied); returns records per destination rank."""
        cloud = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        Tm = np.ascontiguousarray(T, dtype=np.float64).reshape(16)
        counts = (C.c_uint64 * max(1, getattr(self, "_world", 1)))()
        self._check(self._L.vmap_shard_generate(self._h, cloud.ctypes.data, cloud.shape[0], 12,
                                                int(bool(is_dense)), _dptr(Tm), counts))
        return list(counts)

    def shard_generate_device(self, T, data_ptr, n, stride_bytes=12):
        Tm = np.ascontiguousarray(T, dtype=np.float64).reshape(16)
        counts = (C.c_uint64 * max(1, getattr(self, "_world", 1)))()
        self._check(self._L.vmap_shard_generate_dev(self._h, C.c_void_p(int(data_ptr)), int(n),
                                                    int(stride_bytes), _dptr(Tm), counts))
        return list(counts)

    def shard_records(self, dest):
        ptr, n = C.c_void_p(), C.c_size_t(0)
        self._check(self._L.vmap_shard_records(self._h, int(dest), C.byref(ptr), C.byref(n)))
        return (ptr.value or 0), int(n.value)

    def shard_apply(self, lists):
        """lists: [(device_ptr, n_records), ...] in scan order."""
        k = len(lists)
        ptrs = (C.c_void_p * max(k, 1))(*[C.c_void_p(int(p)) for p, _ in lists])
        cnts = (C.c_size_t * max(k, 1))(*[int(n) for _, n in lists])
        self._check(self._L.vmap_shard_apply_dev(self._h, ptrs, cnts, k))

    def shard_query_points_device(self, xyz_ptr, n, status_ptr):
        self._check(self._L.vmap_shard_query_points_dev(self._h, C.c_void_p(int(xyz_ptr)), int(n),
                                                        C.c_void_p(int(status_ptr))))

    def shard_query_lines_device(self, ab_ptr, n, packed_ptr):
        self._check(self._L.vmap_shard_query_lines_dev(self._h, C.c_void_p(int(ab_ptr)), int(n),
                                                       C.c_void_p(int(packed_ptr))))

    @staticmethod
    def dist_unique_id():
        buf = (C.c_uint8 * 128)()
        L = _capi.lib()
        st = L.vmap_dist_unique_id(buf)
        if st != 0:
            raise VmapError(st, (L.vmap_last_error(None) or b"").decode())
        return bytes(buf)

    def dist_init(self, id128, rank, world):
        buf = (C.c_uint8 * 128)(*bytes(id128))
        self._check(self._L.vmap_dist_init(self._h, buf, int(rank), int(world)))
        self._world = int(world)

    @staticmethod
    def dist_attach_local(worlds, records_per_round=0):
        """Single-process stand-in for a multi-GPU job: `worlds[r]` becomes rank r of len(worlds)
        (peer-memory transport with plain pointers; see vmap_dist_attach_local)."""
        L = _capi.lib()
        arr = (C.c_void_p * len(worlds))(*[w._h for w in worlds])
        st = L.vmap_dist_attach_local(arr, len(worlds), int(records_per_round))
        if st != 0:
            raise VmapError(st, (L.vmap_last_error(worlds[0]._h) or b"").decode())
        for w in worlds:
            w._world = len(worlds)

    def dist_flush_begin(self):
        self._check(self._L.vmap_dist_flush_begin(self._h))

    def dist_flush(self):
        self._check(self._L.vmap_dist_flush(self._h))

    def dist_set_batch(self, scans_per_rank):
        self._check(self._L.vmap_dist_set_batch(self._h, int(scans_per_rank)))

    def set_debug_l2_flush(self, n_bytes):
        self._check(self._L.vmap_set_debug_l2_flush(self._h, int(n_bytes)))

    def dist_shutdown(self):
        self._check(self._L.vmap_dist_shutdown(self._h))

    def dist_insert_pointcloud(self, T, xyz, is_dense=True):
        """COLLECTIVE: this rank's scan becomes the rank-th scan of the round on every shard."""
        cloud = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        Tm = np.ascontiguousarray(T, dtype=np.float64).reshape(16)
        self._check(self._L.vmap_dist_insert_pointcloud(self._h, cloud.ctypes.data, cloud.shape[0],
                                                        12, int(bool(is_dense)), _dptr(Tm)))

    def dist_insert_pointcloud_device(self, T, data_ptr, n, stride_bytes=12):
        Tm = np.ascontiguousarray(T, dtype=np.float64).reshape(16)
        self._check(self._L.vmap_dist_insert_pointcloud_dev(self._h, C.c_void_p(int(data_ptr)),
                                                            int(n), int(stride_bytes), _dptr(Tm)))

    def dist_query_points_device(self, xyz_ptr, n, status_ptr):
        self._check(self._L.vmap_dist_query_points_dev(self._h, C.c_void_p(int(xyz_ptr)), int(n),
                                                       C.c_void_p(int(status_ptr))))

    def dist_query_lines_device(self, ab_ptr, n, status_ptr):
        self._check(self._L.vmap_dist_query_lines_dev(self._h, C.c_void_p(int(ab_ptr)), int(n),
                                                      C.c_void_p(int(status_ptr))))

    # -- leaves / introspection ---------------------------------------------------------
    def leaf_count(self):
        n = C.c_size_t(0)
        self._check(self._L.vmap_leaf_count(self._h, C.byref(n)))
        return int(n.value)

    def export_leaves(self):
        n = self.leaf_count()
        k = np.zeros((n, 3), dtype=np.uint16)
        v = np.zeros(n, dtype=np.float32)
        got = C.c_size_t(0)
        self._check(self._L.vmap_export_leaves(self._h, k.ctypes.data, v.ctypes.data, n,
                                               C.byref(got)))
        return k[: got.value], v[: got.value]

    def export_leaves_device(self, keys_ptr, logodds_ptr, capacity):
        """Leaves straight into device buffers (capacity x 3 uint16 keys, capacity float32 values) of
        this handle's GPU; returns the leaf count (VMAP_ERR_BUFFER_TOO_SMALL when capacity is too small)."""
        n = C.c_size_t(0)
        self._check(self._L.vmap_export_leaves_dev(self._h, C.c_void_p(int(keys_ptr)), C.c_void_p(int(logodds_ptr)),
                                                   int(capacity), C.byref(n)))
        return int(n.value)

    def import_leaves_device(self, keys_ptr, logodds_ptr, n):
        self._check(self._L.vmap_import_leaves_dev(self._h, C.c_void_p(int(keys_ptr)), C.c_void_p(int(logodds_ptr)), int(n)))

    def leaf_digest(self):
        """(number of leaves, order-independent 64-bit digest of (key, value bits)); see
        sharded.leaf_digest for the host mirror."""
        n, d = C.c_uint64(0), C.c_uint64(0)
        self._check(self._L.vmap_leaf_digest(self._h, C.byref(n), C.byref(d)))
        return int(n.value), int(d.value)

    def getLineStatusProbesDevice(self, ab_ptr, n, status_ptr):
        """getLineStatusDevice that also returns the total number of voxels probed."""
        probes = C.c_uint64(0)
        self._check(self._L.vmap_query_lines_probes_dev(self._h, C.c_void_p(int(ab_ptr)), int(n),
                                                        C.c_void_p(int(status_ptr)), C.byref(probes)))
        return int(probes.value)

    def import_leaves(self, keys_k3, logodds):
        k = np.ascontiguousarray(keys_k3, dtype=np.uint16).reshape(-1, 3)
        v = np.ascontiguousarray(logodds, dtype=np.float32).reshape(-1)
        assert k.shape[0] == v.shape[0]
        self._check(self._L.vmap_import_leaves(self._h, k.ctypes.data, v.ctypes.data, k.shape[0]))

    # -- on-demand export to octomap's tree form (octomap_world.cc:820-856) ----------------
    def tree_size(self, max_likelihood=False):
        """octree_->prune(); octree_->size()."""
        n = C.c_size_t(0)
        self._check(self._L.vmap_tree_size(self._h, int(bool(max_likelihood)), C.byref(n)))
        return int(n.value)

    def writeOctomapToBinaryConst(self):
        """The .bt stream as bytes (octomap_world.cc:854-856)."""
        n = C.c_size_t(0)
        self._check(self._L.vmap_serialize_bt(self._h, None, 0, C.byref(n)))
        buf = (C.c_uint8 * max(1, n.value))()
        self._check(self._L.vmap_serialize_bt(self._h, buf, n.value, C.byref(n)))
        return bytes(buf[: n.value])

    def getOctomapFullMsgData(self, with_header=False):
        """Payload of getOctomapFullMsg (octomap_world.cc:826-831): OcTree::writeData bytes; with
        the header it is a complete .ot stream."""
        n = C.c_size_t(0)
        self._check(self._L.vmap_serialize_ot(self._h, int(bool(with_header)), None, 0, C.byref(n)))
        buf = (C.c_uint8 * max(1, n.value))()
        self._check(self._L.vmap_serialize_ot(self._h, int(bool(with_header)), buf, n.value, C.byref(n)))
        return bytes(buf[: n.value])

    def setOctomapFromFullMsgData(self, data, with_header=False):
        b = bytes(data)
        buf = (C.c_uint8 * max(1, len(b))).from_buffer_copy(b or b"\0")
        self._check(self._L.vmap_deserialize_ot(self._h, int(bool(with_header)), buf, len(b)))

    def writeOctomapToFile(self, filename):
        """octomap_world.cc:849-852; like OcTree::writeBinary it thresholds the live map."""
        self._check(self._L.vmap_write_bt(self._h, str(filename).encode(), 1))
        return True

    def loadOctomapFromFile(self, filename):
        self._check(self._L.vmap_load_bt(self._h, str(filename).encode()))
        return True

    def setOctomapFromBinaryMsg(self, data):
        b = bytes(data)
        buf = (C.c_uint8 * max(1, len(b))).from_buffer_copy(b or b"\0")
        self._check(self._L.vmap_deserialize_bt(self._h, buf, len(b)))

    def set_debug_capture(self, enabled=True):
        self._check(self._L.vmap_set_debug_capture(self._h, int(bool(enabled))))

    def last_scan_keys(self):
        nf, no = C.c_size_t(0), C.c_size_t(0)
        self._check(self._L.vmap_debug_last_scan_keys(self._h, None, 0, C.byref(nf), None, 0,
                                                      C.byref(no)))
        f = np.zeros((nf.value, 3), dtype=np.uint16)
        o = np.zeros((no.value, 3), dtype=np.uint16)
        if nf.value or no.value:
            self._check(self._L.vmap_debug_last_scan_keys(
                self._h, f.ctypes.data if nf.value else None, nf.value, C.byref(nf),
                o.ctypes.data if no.value else None, no.value, C.byref(no)))
        return f, o

    def last_scan_stats(self):
        s = _capi.ScanStats()
        self._check(self._L.vmap_last_scan_stats(self._h, C.byref(s)))
        return s.as_dict()

    def logodds_constants(self):
        out = (C.c_float * 5)()
        self._check(self._L.vmap_logodds_constants(self._h, out))
        return np.array(list(out), dtype=np.float32)

    def stream(self):
        return int(self._L.vmap_stream(self._h) or 0)

    def last_walk_kernel_ms(self):
        """Duration of the segment walker's launch alone in the last synchronous scan (0: not recorded)."""
        return float(self._L.vmap_last_walk_kernel_ms(self._h))

    def kernel_launches(self):
        return int(self._L.vmap_kernel_launches(self._h))

    def device_bytes(self):
        return int(self._L.vmap_device_bytes(self._h))
