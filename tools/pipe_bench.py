"""Throughput of the scan pipeline on the bench scene (cfg3): K scans inserted back to back, region-timed
on the device (vmap_region_begin / _end), for pipeline depth 0 / 1 / 2, device-resident and pinned-host
points, with and without the in-stream 256 MiB L2 flush before every scan.  All variants must build the
same map (leaf digest).      usage: tools/pipe_bench.py [n_scans] [resolution]"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import volumetric_mapping_b200 as vm           # noqa: E402
from volumetric_mapping_b200 import scenes     # noqa: E402

n_scans = int(sys.argv[1]) if len(sys.argv) > 1 else 40
res = float(sys.argv[2]) if len(sys.argv) > 2 else 0.1
scans = [scenes.trajectory_scan(i, resolution=res) for i in range(n_scans)]
dev = [torch.from_numpy(s["points"]).cuda() for s in scans]
pin = [torch.from_numpy(s["points"]).pin_memory() for s in scans]
n = scans[0]["points"].shape[0]
import ctypes as C
T_c = [(C.c_double * 16)(*np.ascontiguousarray(s["T"], dtype=np.float64).reshape(16)) for s in scans]
warm = 5
out = []
ref = None
for flush in (0, 256 << 20):
    for kind in ("dev", "host"):
        for depth in (0, 1, 2):
            w = vm.OctomapWorld(resolution=res, sensor_max_range=50.0, initial_blocks=1 << 20)
            w.set_pipeline_depth(depth)
            w.set_debug_l2_flush(flush)
            for i in range(warm):
                w.insertPointcloudDevice(scans[i]["T"], dev[i].data_ptr(), n)
            w.region_begin()
            t0 = time.perf_counter()
            for i in range(warm, n_scans):
                if kind == "dev":
                    w.insertPointcloudDevice(scans[i]["T"], dev[i].data_ptr(), n)
                else:
                    w.insertPointcloudHostPtr(T_c[i], pin[i].data_ptr(), n)
            t_enq = time.perf_counter() - t0
            ms = w.region_end()
            t = w.totals()
            dg = w.leaf_digest()
            if ref is None:
                ref = dg
            k = n_scans - warm
            row = {"flush_mb": flush >> 20, "input": kind, "depth": depth, "ms_per_scan": ms / k,
                   "host_enqueue_ms_per_scan": 1e3 * t_enq / k, "replayed": t["replayed"],
                   "same_map": dg == ref, "sum_stage_ms_per_scan": {x: t[x] / t["scans"] for x in ("gpu_ms", "walk_ms", "update_ms")}}
            out.append(row)
            print(json.dumps(row), flush=True)
            w.close()
json.dump(out, open("gpurun_out/pipe_bench.json", "w"), indent=1)
print("ALL_SAME" if all(r["same_map"] for r in out) else "MISMATCH")
