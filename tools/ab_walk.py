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


def run(addr, variant, ctas, order=0, counted=1, far_blind=0, n_use=None, reps=3, **params):
    os.environ["VMAP_WALK_COUNTED"] = str(counted)
    os.environ["VMAP_WALK_FAR_BLIND"] = str(far_blind)
    os.environ["VMAP_WALK_ADDR"] = str(addr)
    os.environ["VMAP_WALK_VARIANT"] = str(variant)
    os.environ["VMAP_WALK_CTAS"] = str(ctas)
    os.environ["VMAP_WALK_ORDER"] = str(order)
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


# (VMAP_WALK_ADDR, near limit, CTAs per SM, VMAP_WALK_ORDER); round-1 results for ADDR 0..4 are in
# profiles/r01_ab_walk_variants.json -- ADDR 5 (near-field combining) and ORDER 1 (golden-ratio
# chunk order) have not run on hardware yet
# round 2: + counted loops (VMAP_WALK_COUNTED) and the 64-register build (ADDR 6)
# + fine-bin ray order (always on) and blind far-field touched REDs (VMAP_WALK_FAR_BLIND)
CONFIGS = [(0, 48, 5, 0, 1, 0), (7, 48, 5, 0, 1, 0), (9, 48, 5, 0, 1, 0), (10, 48, 5, 0, 1, 0), (9, 64, 5, 0, 1, 0),
           (10, 64, 5, 0, 1, 0), (9, 48, 5, 0, 1, 0), (10, 48, 5, 0, 1, 0)]
params = {"max_free_space": 6.0, "min_height_free_space": 0.3} if "--filter" in sys.argv else {}
base = None
ok_all = True
for cfg in CONFIGS:
    summary, counts, k, v = run(*cfg, **params)
    if base is None:
        base = (counts, k, v)
    same = bool(counts == base[0] and np.array_equal(k, base[1]) and np.array_equal(v, base[2]))
    ok_all = ok_all and same
    row = {"addr": cfg[0], "near_limit": cfg[1], "ctas": cfg[2], "order": cfg[3], "counted": cfg[4], "far_blind": cfg[5],
           "identical": same, **summary}
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
    for addr in (0, 9, 10):
        _, _, k, v = run(addr, 48, 5, order=0, counted=1, far_blind=0, n_use=2, reps=1)
        ok = bool(np.array_equal(k, po[oo]) and np.array_equal(v, np.asarray(vo, np.float32).view(np.uint32)[oo]))
        out["oracle_parity_addr%d" % addr] = ok
        ok_all = ok_all and ok
        print("oracle parity (VMAP_WALK_ADDR=%d):" % addr, ok, len(po), flush=True)
out["seconds"] = time.time() - t_start
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/ab_walk.json", "w"), indent=1)
print("ALL_IDENTICAL" if ok_all else "MISMATCH")
