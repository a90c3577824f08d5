"""Stand-alone GPU check of getCellStatusBoundingBox / checkPathForCollisionsWithRobot against the
oracle (run in its own process by tests/test_gpu_zz_first_hardware_run.py, so that a device fault in this
not-yet-hardware-proven kernel cannot disturb the rest of the GPU suite)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle                                   # noqa: E402
import volumetric_mapping_b200 as vm            # noqa: E402
from volumetric_mapping_b200 import scenes      # noqa: E402


def main():
    bad_total = 0
    for res, unk in ((0.1, 1), (0.15, 0), (0.25, 0)):
        kw = dict(resolution=res, sensor_max_range=12.0, treat_unknown_as_occupied=unk)
        ow = oracle.OracleWorld(**kw)
        w = vm.OctomapWorld(**kw)
        for i in range(3):
            s = scenes.hdl64_scan(position=(0.4 * i, 0.1 * i, 1.2), yaw_rad=0.3 * i, box=(-6.0, 7.0, -5.0, 6.0),
                                  seed=i, n_elev=24, n_azim=240)
            for _ in range(2):
                ow.insertPointcloud(s["T"], s["points"])
                w.insertPointcloud(s["T"], s["points"])
        rng = np.random.default_rng(3)
        pts = rng.uniform([-7, -6, -0.5], [8, 7, 3.0], size=(4000, 3))
        g = np.arange(-16, 17)
        gx, gy = np.meshgrid(g, g)
        grid = np.stack([gx.ravel() * res, gy.ravel() * res, np.full(gx.size, 1.0)], 1)
        for q in (pts, grid):
            for size in ([0.05, 0.05, 0.05], [0.3, 0.3, 0.3], [0.5, 0.4, 0.25], [1.0, 1.0, 1.0], [2 * res] * 3):
                want = ow.getCellStatusBoundingBox(q, size)
                got = w.getCellStatusBoundingBox(q, size)
                bad = int(np.count_nonzero(want != got))
                if bad:
                    i = int(np.flatnonzero(want != got)[0])
                    print("MISMATCH res=%g unk=%d size=%s: %d of %d, first %s got %d want %d" %
                          (res, unk, size, bad, len(q), q[i], got[i], want[i]))
                bad_total += bad
        path = np.stack([np.linspace(0.0, 9.0, 181), np.full(181, 0.3), np.full(181, 1.2)], 1)
        w.setRobotSize([0.5, 0.5, 0.5])
        a = ow.checkPathForCollisionsWithRobot(path, [0.5, 0.5, 0.5])
        b = w.checkPathForCollisionsWithRobot(path)
        if a != b:
            print("PATH MISMATCH", a, b)
            bad_total += 1
        w.close()
    print("BBOX_OK" if bad_total == 0 else "BBOX_FAILED %d" % bad_total)
    return 0 if bad_total == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
