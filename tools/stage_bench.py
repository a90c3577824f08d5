"""Per-stage timings under bench conditions (L2 flushed between scans): mean over n scans."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import volumetric_mapping_b200 as vm
from volumetric_mapping_b200 import scenes
res = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16
w = vm.OctomapWorld(resolution=res, sensor_max_range=50.0, initial_blocks=1 << 20)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
acc = {}
for i in range(n + 4):
    s = scenes.trajectory_scan(i, resolution=res)
    d = torch.from_numpy(s["points"]).cuda()
    flush.fill_(i & 255); torch.cuda.synchronize()
    w.insertPointcloudDevice(s["T"], d.data_ptr(), d.shape[0])
    st = w.last_scan_stats()
    if i >= 4:
        for k in ("gpu_ms", "front_ms", "resolve_ms", "walk_ms", "update_ms"):
            acc.setdefault(k, []).append(st[k])
print("res %.2f: " % res + "  ".join("%s %.3f" % (k[:-3], np.mean(v)) for k, v in acc.items()))
