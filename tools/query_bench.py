"""cfg5: batched getLineStatus / getCellStatusPoint (1M each) against a 0.1 m map built from N scans."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import volumetric_mapping_b200 as vm
from volumetric_mapping_b200 import scenes

n_scans = int(sys.argv[1]) if len(sys.argv) > 1 else 100
w = vm.OctomapWorld(resolution=0.1, sensor_max_range=50.0, initial_blocks=1 << 20)
for i in range(n_scans):
    s = scenes.trajectory_scan(i, resolution=0.1)
    d = torch.from_numpy(s["points"]).cuda()
    w.insertPointcloudDevice(s["T"], d.data_ptr(), d.shape[0])
print("map built: blocks", w.last_scan_stats()["blocks_total"], "leaves", w.leaf_count())
q = scenes.cfg5_queries(bounds=((-45.0, 45.0 + 0.5 * n_scans), (-45.0, 45.0), (-0.5, 4.0)))
seg = torch.from_numpy(q["segments"]).cuda(); pts = torch.from_numpy(q["points"]).cuda()
sl = torch.zeros(seg.shape[0], dtype=torch.uint8, device="cuda"); sp = torch.zeros(pts.shape[0], dtype=torch.uint8, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for rep in range(3):
    flush.fill_(rep); torch.cuda.synchronize()
    t0 = time.perf_counter(); w.getLineStatusDevice(seg.data_ptr(), seg.shape[0], sl.data_ptr()); t1 = time.perf_counter()
    flush.fill_(rep); torch.cuda.synchronize()
    t2 = time.perf_counter(); w.getCellStatusPointDevice(pts.data_ptr(), pts.shape[0], sp.data_ptr()); t3 = time.perf_counter()
    print("lines: %.3f ms (%.1f M seg/s)   points: %.3f ms (%.1f M pts/s)" % (1e3*(t1-t0), seg.shape[0]/(t1-t0)/1e6, 1e3*(t3-t2), pts.shape[0]/(t3-t2)/1e6))
print("line status histogram", np.bincount(sl.cpu().numpy(), minlength=3), "point", np.bincount(sp.cpu().numpy(), minlength=3))
