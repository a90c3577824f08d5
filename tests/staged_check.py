"""Stand-alone GPU check of vmap_stage_pointcloud / vmap_insert_staged: the pipelined calls must
build exactly the map the plain calls build (and that map equals the oracle's on small scans)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle                                   # noqa: E402
import volumetric_mapping_b200 as vm            # noqa: E402
from volumetric_mapping_b200 import scenes      # noqa: E402


def leaves(w):
    k, v = w.export_leaves()
    o = np.argsort(oracle.pack_keys(k))
    return oracle.pack_keys(k)[o], np.asarray(v, np.float32).view(np.uint32)[o]


def main():
    kw = dict(resolution=0.1, sensor_max_range=15.0)
    scans = [scenes.hdl64_scan(position=(0.5 * i, 0.2 * i, 1.5), yaw_rad=0.2 * i, box=(-8.0, 9.0, -7.0, 8.0), seed=i,
                               n_elev=32, n_azim=300 + 40 * i) for i in range(6)]     # different sizes: buffers grow
    pts = [np.ascontiguousarray(s["points"], np.float32) for s in scans]
    plain, piped, ow = vm.OctomapWorld(**kw), vm.OctomapWorld(**kw), oracle.OracleWorld(**kw)
    for s, p in zip(scans, pts):
        plain.insertPointcloud(s["T"], p)
        ow.insertPointcloud(s["T"], p)
    piped.stagePointcloudHostPtr(pts[0].ctypes.data, len(pts[0]))
    for i, s in enumerate(scans):
        if i + 1 < len(scans):
            piped.stagePointcloudHostPtr(pts[i + 1].ctypes.data, len(pts[i + 1]))
        piped.insertStaged(s["T"])
    a, b, c = leaves(plain), leaves(piped), leaves(ow)
    ok = all(np.array_equal(x, y) for x, y in zip(a, b)) and all(np.array_equal(x, y) for x, y in zip(a, c))
    # misuse is an error, not a crash
    try:
        piped.insertStaged(scans[0]["T"])
        ok = False
    except vm.VmapError:
        pass
    print("%d leaves" % len(a[0]))
    print("STAGED_OK" if ok else "STAGED_FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
