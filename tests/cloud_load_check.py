"""Stand-alone GPU check of the occupied-only point-cloud loader (loadOctomapCallback's .pcd / .ply
branch, octomap_manager.cc:313-344) against the oracle's updateOccupancy on the same keys."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle                                   # noqa: E402
import volumetric_mapping_b200 as vm            # noqa: E402


def main():
    rng = np.random.default_rng(4)
    pts = (rng.normal(size=(20000, 3)) * 4.0).astype(np.float32)
    pts[::97] = pts[1::97][: len(pts[::97])]            # duplicates collapse in the KeySet
    pts[5] = [np.nan, 0, 0]
    pts[6] = [1e9, 0, 0]                                 # outside the map: skipped
    ok = True
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "cloud.pcd")
        with open(path, "wb") as o:
            o.write(("VERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH %d\nHEIGHT 1\nPOINTS %d\nDATA binary\n"
                     % (len(pts), len(pts))).encode())
            o.write(pts.astype("<f4").tobytes())
        for res in (0.1, 0.25):
            ow = oracle.OracleWorld(resolution=res)
            w = vm.OctomapWorld(resolution=res)
            keys = []
            for p in pts:
                k = [ow.coordToKeyChecked(float(c)) for c in p]
                if all(a for a, _ in k):
                    keys.append([b for _, b in k])
            occ = np.asarray(keys, np.uint16)
            for rep in range(2):                         # twice: 0.62 then 1.24 log-odds
                ow.updateOccupancy(np.zeros((0, 3), np.uint16), occ)
                n = w.loadOccupiedPointcloudFile(path) if rep == 0 else (w.insertOccupiedPoints(pts), len(pts))[1]
                ok = ok and n == len(pts)
            ka, va = ow.export_leaves()
            kb, vb = w.export_leaves()
            oa, ob = np.argsort(oracle.pack_keys(ka)), np.argsort(oracle.pack_keys(kb))
            same = (np.array_equal(oracle.pack_keys(ka)[oa], oracle.pack_keys(kb)[ob]) and
                    np.array_equal(np.asarray(va, np.float32).view(np.uint32)[oa], np.asarray(vb, np.float32).view(np.uint32)[ob]))
            print("res %g: %d leaves, identical=%s" % (res, len(va), same))
            ok = ok and same
            w.close()
    print("CLOUD_LOAD_OK" if ok else "CLOUD_LOAD_FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
