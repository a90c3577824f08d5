"""Small driver for ncu captures of the scan kernels: inserts a few cfg3 scans synchronously (pipeline
depth 0) so that every kernel of a scan appears once per scan in the launch list.
   usage: tools/ncu_walk.py [n_scans] [resolution]"""
import sys

sys.path.insert(0, ".")
import volumetric_mapping_b200 as vm           # noqa: E402
from volumetric_mapping_b200 import scenes     # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
res = float(sys.argv[2]) if len(sys.argv) > 2 else 0.1
w = vm.OctomapWorld(resolution=res, sensor_max_range=50.0, initial_blocks=1 << 20)
w.set_pipeline_depth(0)
for i in range(n):
    s = scenes.trajectory_scan(i, resolution=res)
    w.insertPointcloud(s["T"], s["points"])
print(w.last_scan_stats())
w.close()
