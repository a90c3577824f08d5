"""End-to-end scans/s from pinned host memory, plain vs pipelined (not yet run on hardware):
  plain      vmap_insert_pointcloud per scan (H2D copy, kernels, counter read-back in sequence)
  pipelined  vmap_stage_pointcloud(scan k+1) before vmap_insert_staged(scan k): the copy of the next
             scan overlaps the kernels of the current one
Both build the same map (leaves compared bit for bit).   usage: tools/e2e_pipeline_bench.py [n_scans]"""
import ctypes as C
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import volumetric_mapping_b200 as vm           # noqa: E402
from volumetric_mapping_b200 import scenes     # noqa: E402
import oracle                                   # noqa: E402  (pack_keys only)

n_scans = int(sys.argv[1]) if len(sys.argv) > 1 else 40
scans = [scenes.trajectory_scan(i, resolution=0.1) for i in range(n_scans)]
pinned = [torch.from_numpy(s["points"]).pin_memory() for s in scans]
n_pts = [p.shape[0] for p in pinned]
T_c = [(C.c_double * 16)(*np.ascontiguousarray(s["T"], dtype=np.float64).reshape(16)) for s in scans]


def leaves(w):
    k, v = w.export_leaves()
    o = np.argsort(oracle.pack_keys(k))
    return oracle.pack_keys(k)[o], np.asarray(v, np.float32).view(np.uint32)[o]


def run(pipelined):
    w = vm.OctomapWorld(resolution=0.1, sensor_max_range=50.0, initial_blocks=1 << 20)
    best = 1e9
    for rep in range(3):
        w.resetMap()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if pipelined:
            w.stagePointcloudHostPtr(pinned[0].data_ptr(), n_pts[0])
            for i in range(n_scans):
                if i + 1 < n_scans:
                    w.stagePointcloudHostPtr(pinned[i + 1].data_ptr(), n_pts[i + 1])
                w.insertStaged(scans[i]["T"])
        else:
            for i in range(n_scans):
                w.insertPointcloudHostPtr(T_c[i], pinned[i].data_ptr(), n_pts[i])
        best = min(best, time.perf_counter() - t0)
    out = leaves(w)
    w.close()
    return best, out


t_plain, a = run(False)
t_pipe, b = run(True)
print("plain     %.3f ms/scan  %.0f scans/s" % (1e3 * t_plain / n_scans, n_scans / t_plain))
print("pipelined %.3f ms/scan  %.0f scans/s" % (1e3 * t_pipe / n_scans, n_scans / t_pipe))
print("IDENTICAL" if np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) else "MISMATCH")
