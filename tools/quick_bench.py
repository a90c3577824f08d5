"""Scratch timing of the cfg3 scan (device-resident points): per-scan gpu_ms and stats."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import volumetric_mapping_b200 as vm
from volumetric_mapping_b200 import scenes

res = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
n_scans = int(sys.argv[2]) if len(sys.argv) > 2 else 20
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
w = vm.OctomapWorld(resolution=res, sensor_max_range=50.0, initial_blocks=1 << 20, dedupe_mode=(1 if mode == 1 else 0), dense_window_bytes=(1 if mode == 2 else 0))
scans = [scenes.trajectory_scan(i, resolution=res) for i in range(n_scans)]
dev = [torch.from_numpy(s["points"]).cuda() for s in scans]
torch.cuda.synchronize()
for rep in range(2):
    w.resetMap()
    t0 = time.perf_counter()
    tot = 0
    for s, d in zip(scans, dev):
        w.insertPointcloudDevice(s["T"], d.data_ptr(), d.shape[0])
        st = w.last_scan_stats()
        tot += st["traversals"]
        if rep == 1:
            print("scan gpu_ms=%.3f [f %.3f r %.3f w %.3f u %.3f] trav=%d rays=%d ufree=%d uocc=%d blocks_touched=%d total=%d retries=%d" % (
                st["gpu_ms"], st["front_ms"], st["resolve_ms"], st["walk_ms"], st["update_ms"], st["traversals"], st["rays_cast"], st["unique_free"], st["unique_occupied"],
                st["blocks_touched"], st["blocks_total"], st["retries"]))
    dt = time.perf_counter() - t0
    print("rep %d: %d scans %.3f ms/scan wall, %.3e traversals/s" % (rep, n_scans, 1e3 * dt / n_scans, tot / dt))
print("device MB", w.device_bytes() / 1e6, "launches", w.kernel_launches())
