"""A/B of the segment walker's variants on the bench scene (torch-free, starts fast).  Every
configuration (VMAP_WALK_ADDR, VMAP_WALK_VARIANT, VMAP_WALK_CTAS) must produce the same map bit
for bit as the default one (leaves + per-scan statistics); prints the event-timed stage times.
   usage: tools/ab_walk.py [n_scans] [--oracle] [--filter]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import volumetric_mapping_b200 as vm           # noqa: E402
from volumetric_mapping_b200 import scenes     # noqa: E402
import oracle                                   # noqa: E402  (pack_keys; OracleWorld only with --oracle)

os.environ["VMAP_TRACE"] = "1"
t_start = time.time()
args = [a for a in sys.argv[1:] if not a.startswith("--")]
n_scans = int(args[0]) if args else 10
scans = [scenes.trajectory_scan(i, resolution=0.1) for i in range(n_scans)]
out = {"n_scans": n_scans, "configs": []}
KEYS = ("walk_ms", "gpu_ms", "update_ms", "front_ms", "resolve_ms")


def run(addr, variant, ctas, n_use=None, reps=3, **params):
    os.environ["VMAP_WALK_ADDR"] = str(addr)
    os.environ["VMAP_WALK_VARIANT"] = str(variant)
    os.environ["VMAP_WALK_CTAS"] = str(ctas)
    w = vm.OctomapWorld(resolution=0.1, sensor_max_range=50.0, initial_blocks=1 << 20, **params)
    stats = []
    for rep in range(reps):
        w.resetMap()
        stats = []
        for s in scans[:n_use]:
            w.insertPointcloud(s["T"], s["points"])
            stats.append(w.last_scan_stats())
    k, v = w.export_leaves()
    p = oracle.pack_keys(k)
    o = np.argsort(p)
    summary = {key: float(np.median([st[key] for st in stats])) for key in KEYS}
    counts = [[int(st[c]) for st in stats] for c in ("traversals", "unique_free", "unique_occupied", "longest_ray")]
    w.close()
    return summary, counts, p[o], np.asarray(v, np.float32).view(np.uint32)[o]


CONFIGS = [(0, 48, 5), (1, 48, 5), (2, 48, 5), (3, 48, 6), (4, 48, 6), (1, 2, 5), (1, 16, 5), (1, 128, 5),
           (2, 16, 5), (1, 48, 10), (2, 48, 10)]
params = {"max_free_space": 6.0, "min_height_free_space": 0.3} if "--filter" in sys.argv else {}
base = None
ok_all = True
for cfg in CONFIGS:
    summary, counts, k, v = run(*cfg, **params)
    if base is None:
        base = (counts, k, v)
    same = bool(counts == base[0] and np.array_equal(k, base[1]) and np.array_equal(v, base[2]))
    ok_all = ok_all and same
    row = {"addr": cfg[0], "near_limit": cfg[1], "ctas": cfg[2], "identical": same, **summary}
    out["configs"].append(row)
    print(json.dumps(row), flush=True)
out["leaves"] = int(len(base[1]))

if "--oracle" in sys.argv:      # bit-exact leaves against the CPU oracle on the first two scans
    ow = oracle.OracleWorld(resolution=0.1, sensor_max_range=50.0)
    for s in scans[:2]:
        ow.insertPointcloud(s["T"], s["points"])
    ko, vo = ow.export_leaves()
    po = oracle.pack_keys(ko)
    oo = np.argsort(po)
    for addr in (0, 1, 2, 3, 4):
        _, _, k, v = run(addr, 48, 6 if addr >= 3 else 5, n_use=2, reps=1)
        ok = bool(np.array_equal(k, po[oo]) and np.array_equal(v, np.asarray(vo, np.float32).view(np.uint32)[oo]))
        out["oracle_parity_addr%d" % addr] = ok
        ok_all = ok_all and ok
        print("oracle parity (VMAP_WALK_ADDR=%d):" % addr, ok, len(po), flush=True)
out["seconds"] = time.time() - t_start
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/ab_walk.json", "w"), indent=1)
print("ALL_IDENTICAL" if ok_all else "MISMATCH")
