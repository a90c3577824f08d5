"""Generation + pack time of one rank of a G-way sharded map (1 GPU; nothing is exchanged)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import volumetric_mapping_b200 as vm
from volumetric_mapping_b200 import scenes
G = int(sys.argv[1]) if len(sys.argv) > 1 else 8
w = vm.OctomapWorld(resolution=0.1, sensor_max_range=50.0)
w.set_shard(0, G)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
g, wl = [], []
for i in range(14):
    s = scenes.trajectory_scan(i, resolution=0.1)
    d = torch.from_numpy(s["points"]).cuda()
    flush.fill_(i); torch.cuda.synchronize()
    t0 = time.perf_counter()
    counts = w.shard_generate_device(s["T"], d.data_ptr(), d.shape[0])
    dt = time.perf_counter() - t0
    st = w.last_scan_stats()
    if i >= 4:
        g.append(st["gpu_ms"]); wl.append(1e3 * dt)
print("G=%d generate+pack: device %.3f ms, host wall %.3f ms; records/dest %s" % (G, np.mean(g), np.mean(wl), counts[:G]))
