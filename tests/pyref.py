"""Second, independent restatement of the path in pure Python (numpy scalar float32/float64).

Small cases only.  It exists so the C++ oracle is checked by something that shares no code
with it: DDA key lists, the skip rule, key-set precedence and the clamped log-odds update are
re-derived here from SURVEY.md section 3 (octomap_world.cc:110-264 + octomap semantics).
"""
import math

import numpy as np

F = np.float32
D = np.float64
TREE_MAX_VAL = 32768
INT_MIN = -2147483648


def x86_int(v):
    v = float(v)
    if not (v >= -2147483648.0 and v < 2147483648.0):
        return INT_MIN
    return int(v)  # trunc; callers pass floor()ed values


def coord_to_key_checked(c, res):
    factor = D(1.0) / D(res)
    v = D(factor) * D(c)
    scaled = x86_int(math.floor(v) if math.isfinite(v) else v) + TREE_MAX_VAL
    if 0 <= scaled < 2 * TREE_MAX_VAL:
        return True, scaled
    return False, 0


def coord_to_key(c, res):
    factor = D(1.0) / D(res)
    v = D(factor) * D(c)
    scaled = x86_int(math.floor(v) if math.isfinite(v) else v) + TREE_MAX_VAL
    return scaled & 0xFFFF


def key_to_coord(k, res):
    return (D(int(k) - TREE_MAX_VAL) + D(0.5)) * D(res)


def point_key(p, res):
    return tuple(coord_to_key(D(F(c)), res) for c in p)


def point_key_checked(p, res):
    out = []
    for c in p:
        ok, k = coord_to_key_checked(D(F(c)), res)
        if not ok:
            return None
        out.append(k)
    return tuple(out)


def fnorm_sq(v):
    return D(F(F(F(v[0]) * F(v[0])) + F(F(v[1]) * F(v[1]))) + F(F(v[2]) * F(v[2])))


def compute_ray_keys(origin, end, res, max_keys=100000):
    """Returns None when computeRayKeys would return false, else the list of key tuples."""
    origin = [F(c) for c in origin]
    end = [F(c) for c in end]
    ko = point_key_checked(origin, res)
    ke = point_key_checked(end, res)
    if ko is None or ke is None:
        return None
    if ko == ke:
        return []
    ray = [ko]
    direction = [F(end[i] - origin[i]) for i in range(3)]
    with np.errstate(all="ignore"):
        length = F(np.sqrt(fnorm_sq(direction)))
        direction = [F(d / length) for d in direction]
    step = [0, 0, 0]
    t_max = [D(0)] * 3
    t_delta = [D(0)] * 3
    cur = list(ko)
    for i in range(3):
        if direction[i] > 0.0:
            step[i] = 1
        elif direction[i] < 0.0:
            step[i] = -1
        if step[i] != 0:
            border = key_to_coord(cur[i], res)
            border = D(border + D(F(D(step[i]) * D(res) * D(0.5))))
            t_max[i] = D(D(border - D(origin[i])) / D(direction[i]))
            t_delta[i] = D(D(res) / D(abs(D(direction[i]))))
        else:
            t_max[i] = D(np.finfo(np.float64).max)
            t_delta[i] = D(np.finfo(np.float64).max)
    while True:
        if t_max[0] < t_max[1]:
            dim = 0 if t_max[0] < t_max[2] else 2
        else:
            dim = 1 if t_max[1] < t_max[2] else 2
        cur[dim] = (cur[dim] + step[dim]) & 0xFFFF
        with np.errstate(over="ignore"):
            t_max[dim] = D(t_max[dim] + t_delta[dim])
        if tuple(cur) == ke:
            break
        dist = min(min(t_max[0], t_max[1]), t_max[2])
        if dist > D(length):
            break
        if len(ray) >= max_keys - 1:
            break
        ray.append(tuple(cur))
    return ray


def logodds(p):
    return F(math.log(p / (1.0 - p)))


class PyWorld:
    """dict-based flat map: key tuple -> float32 log-odds (absent == unknown)."""

    def __init__(self, resolution=0.15, probability_hit=0.65, probability_miss=0.4,
                 threshold_min=0.12, threshold_max=0.97, threshold_occupancy=0.7,
                 sensor_max_range=5.0, max_free_space=0.0, min_height_free_space=0.0,
                 treat_unknown_as_occupied=True):
        self.res = resolution
        self.hit = logodds(probability_hit)
        self.miss = logodds(probability_miss)
        self.cmin = logodds(threshold_min)
        self.cmax = logodds(threshold_max)
        self.occ = logodds(threshold_occupancy)
        self.max_range = sensor_max_range
        self.max_free_space = max_free_space
        self.min_height_free_space = min_height_free_space
        self.treat_unknown_as_occupied = treat_unknown_as_occupied
        self.leaves = {}
        self.last_free = set()
        self.last_occ = set()
        self.traversals = 0
        self.rays_cast = 0

    # -- updateNode (with the early abort) -----------------------------------------------
    def update_node(self, key, occupied):
        u = self.hit if occupied else self.miss
        v = self.leaves.get(key)
        if v is not None and ((u >= 0 and v >= self.cmax) or (u <= 0 and v <= self.cmin)):
            return
        v = F(0.0) if v is None else v
        v = F(v + u)
        if v < self.cmin:
            v = self.cmin
        elif v > self.cmax:
            v = self.cmax
        self.leaves[key] = v

    def _free_filter(self, key, origin):
        if self.max_free_space == 0.0:
            return True
        vc = [F(key_to_coord(k, self.res)) for k in key]
        d = [F(vc[i] - origin[i]) for i in range(3)]
        n = np.sqrt(fnorm_sq(d))
        return bool(n < self.max_free_space or
                    D(vc[2]) > D(F(origin[2])) - D(self.min_height_free_space))

    def cast_ray(self, origin, point, free, occ):
        self.rays_cast += 1
        d = [F(F(point[i]) - F(origin[i])) for i in range(3)]
        with np.errstate(all="ignore"):
            norm = np.sqrt(fnorm_sq(d))
        if self.max_range < 0.0 or norm <= self.max_range:
            ray = compute_ray_keys(origin, point, self.res)
            if ray is not None:
                self.traversals += len(ray)
                free.update(k for k in ray if self._free_filter(k, origin))
            kc = point_key_checked(point, self.res)
            if kc is not None:
                occ.add(kc)
        else:
            with np.errstate(all="ignore"):
                if norm > 0:
                    ln = F(norm)
                    d = [F(c / ln) for c in d]
                mr = F(self.max_range)
                new_end = [F(F(origin[i]) + F(d[i] * mr)) for i in range(3)]
            ray = compute_ray_keys(origin, new_end, self.res)
            if ray is not None:
                self.traversals += len(ray)
                free.update(k for k in ray if self._free_filter(k, origin))

    def update_occupancy(self, free, occ):
        for k in occ:
            self.update_node(k, True)
            free.discard(k)
        for k in free:
            self.update_node(k, False)
        self.last_free, self.last_occ = free, occ

    def insert_pointcloud(self, T, xyz, is_dense=True):
        T = np.asarray(T, dtype=np.float64)
        pts = [tuple(F(c) for c in p) for p in np.asarray(xyz, dtype=np.float32)]
        if not is_dense:
            pts = [p for p in pts if all(math.isfinite(float(c)) for c in p)]
        world = []
        for p in pts:
            x, y, z = (D(c) for c in p)
            with np.errstate(all="ignore"):
                world.append(tuple(
                    F(D(D(D(T[r, 0] * x) + D(T[r, 1] * y)) + D(T[r, 2] * z)) + T[r, 3])
                    for r in range(3)))
        origin = tuple(F(T[r, 3]) for r in range(3))
        free, occ = set(), set()
        self.traversals = 0
        self.rays_cast = 0
        for p in world:
            if point_key(p, self.res) not in occ:
                self.cast_ray(origin, p, free, occ)
        self.update_occupancy(free, occ)
        return np.array(world, dtype=np.float32).reshape(-1, 3)

    def cell_status(self, key):
        v = self.leaves.get(key)
        if v is None:
            return 1 if self.treat_unknown_as_occupied else 2
        return 1 if v >= self.occ else 0

    def get_cell_status_point(self, p):
        key = []
        for c in p:
            ok, k = coord_to_key_checked(D(c), self.res)
            if not ok:
                return 1 if self.treat_unknown_as_occupied else 2
            key.append(k)
        return self.cell_status(tuple(key))

    def get_line_status(self, a, b):
        ray = compute_ray_keys([F(c) for c in a], [F(c) for c in b], self.res)
        for k in ray or []:
            st = self.cell_status(k)
            if st != 0:
                return st
        return 0
