"""A/B of the scan's helper-kernel chain (torch-free): VMAP_HELPERS (folded resolve / per-axis checkpoints)
x VMAP_PDL (programmatic dependent launch) x VMAP_STAGE_MARKS, synchronous (event-timed gpu_ms and stage
times per scan) and pipelined (region-timed ms per scan, with and without the in-stream 256 MiB L2
flush).  Every configuration must build the same map (leaf digest + per-scan statistics); the first
two scans are also compared with the CPU oracle when --oracle is given.
   usage: tools/ab_helpers.py [n_scans] [--oracle] [--res 0.1]"""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, ".")
import volumetric_mapping_b200 as vm           # noqa: E402
from volumetric_mapping_b200 import scenes     # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
n_scans = int(args[0]) if args else 24
res = float(sys.argv[sys.argv.index("--res") + 1]) if "--res" in sys.argv else 0.1
scans = [scenes.trajectory_scan(i, resolution=res) for i in range(n_scans)]
n = scans[0]["points"].shape[0]

rt = None
for name in ("libcudart.so.12", "libcudart.so", "/usr/local/cuda/lib64/libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"):
    try:
        rt = C.CDLL(name)
        break
    except OSError:
        pass
if rt is None:
    import glob
    import torch  # noqa: F401  (last resort: torch bundles the runtime)
    rt = C.CDLL(glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*"))[0])
rt.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
dev = []
for s in scans:
    p = C.c_void_p()
    assert rt.cudaMalloc(C.byref(p), n * 12) == 0
    pts = np.ascontiguousarray(s["points"], np.float32)
    assert rt.cudaMemcpy(p, pts.ctypes.data, n * 12, 1) == 0
    dev.append(p.value)

KEYS = ("gpu_ms", "front_ms", "resolve_ms", "walk_ms", "update_ms")
COUNTS = ("traversals", "unique_free", "unique_occupied", "longest_ray", "rays_cast", "occupied_emitted", "blocks_touched")


EXTRA = {}                 # further environment switches for the worlds made next


def make(helpers, pdl, marks, ctx=3, flush_at=1):
    os.environ.update(EXTRA)
    os.environ["VMAP_FLUSH_AT"] = str(flush_at)
    os.environ["VMAP_PIPE_CTX"] = str(ctx)
    os.environ["VMAP_HELPERS"] = str(helpers)
    os.environ["VMAP_PDL"] = str(pdl)
    if marks is None:
        os.environ.pop("VMAP_STAGE_MARKS", None)
    else:
        os.environ["VMAP_STAGE_MARKS"] = str(marks)
    return vm.OctomapWorld(resolution=res, sensor_max_range=50.0, initial_blocks=1 << 20)


def run_sync(helpers, pdl, marks, reps=3, flush=0):
    w = make(helpers, pdl, marks)
    w.set_pipeline_depth(0)
    w.set_debug_l2_flush(flush)
    stats = []
    for _ in range(reps):
        w.resetMap()
        stats = []
        for i in range(n_scans):
            w.insertPointcloudDevice(scans[i]["T"], dev[i], n)
            stats.append(w.last_scan_stats())
    dg = w.leaf_digest()
    counts = [[int(st[c]) for c in COUNTS] for st in stats]
    summary = {k: round(float(np.median([st[k] for st in stats[2:]])), 5) for k in KEYS}
    w.close()
    return summary, counts, dg


def run_pipe(helpers, pdl, marks, ctx, flush, warm=4, reps=3, flush_at=1):
    w = make(helpers, pdl, marks, ctx, flush_at)
    w.set_pipeline_depth(2)
    w.set_debug_l2_flush(flush)
    best = None
    for _ in range(reps):
        w.resetMap()
        for i in range(warm):
            w.insertPointcloudDevice(scans[i]["T"], dev[i], n)
        w.region_begin()
        for i in range(warm, n_scans):
            w.insertPointcloudDevice(scans[i]["T"], dev[i], n)
        ms = w.region_end() / (n_scans - warm)
        best = ms if best is None else min(best, ms)
    dg = w.leaf_digest()
    rep = w.totals()["replayed"]
    w.close()
    return round(best, 5), dg, rep


# ---- main (tools/ab_cosched.py executes everything above this line) ----
out = {"n_scans": n_scans, "res": res, "sync": [], "pipelined": []}
base = None
ok = True
# every configuration is visited `rounds` times in turn (clock / thermal drift hits all alike); medians over the visits
rounds = 3
SYNC = [(0, 0, None), (1, 123, None), (0, 0, 0), (1, 123, 0)]
if "--quick" in sys.argv:
    SYNC = SYNC[:2]
    PIPE_SKIP = True
acc = {c: [] for c in SYNC}
for _ in range(rounds):
    for cfg in SYNC:
        summary, counts, dg = run_sync(*cfg, reps=2)
        if base is None:
            base = (counts, dg)
        ok = ok and counts == base[0] and dg == base[1]
        acc[cfg].append(summary)
for cfg in SYNC:
    row = {"helpers": cfg[0], "pdl": cfg[1], "stage_marks": cfg[2],
           **{k: round(float(np.median([a[k] for a in acc[cfg]])), 5) for k in KEYS}}
    out["sync"].append(row)
    print(json.dumps(row), flush=True)
PIPE = [(0, 0, None, 2), (1, 123, None, 3)]   # (helpers, pdl, marks, scan contexts)
for flush in (0, 256 << 20):
    accp = {c: [] for c in PIPE}
    for _ in range(rounds):
        for cfg in PIPE:
            ms, dg, rep = run_pipe(*cfg, flush, reps=2)
            ok = ok and dg == base[1]
            accp[cfg].append(ms)
    for cfg in PIPE:
        row = {"helpers": cfg[0], "pdl": cfg[1], "contexts": cfg[3], "flush_mb": flush >> 20, "ms_per_scan": round(float(np.median(accp[cfg])), 5),
               "min": min(accp[cfg]), "max": max(accp[cfg])}
        out["pipelined"].append(row)
        print(json.dumps(row), flush=True)
# where and how the benchmark's L2 flush runs (three contexts, folded helpers)
out["flush_placement"] = []
FL = [(0,), (1,)]
accf = {c: [] for c in FL}
for _ in range(rounds):
    for c in FL:
        ms, dg, rep = run_pipe(1, 123, None, 3, 256 << 20, reps=2, flush_at=c[0])
        ok = ok and dg == base[1]
        accf[c].append(ms)
for c in FL:
    row = {"flush_at": c[0], "ms_per_scan": round(float(np.median(accf[c])), 5), "min": min(accf[c]), "max": max(accf[c])}
    out["flush_placement"].append(row)
    print(json.dumps(row), flush=True)
# K4 variants: k_update<0> (row by row) vs k_update_mlp (rows in flight together, blocks pipelined)
out["update_variants"] = []
accu = {v: {"sync": [], "sync_flush": [], "pipe": [], "pipe_flush": []} for v in (0, 2)}
for _ in range(rounds):
    for v in (0, 2):
        EXTRA["VMAP_UPDATE_VARIANT"] = str(v)
        s0, counts, dg = run_sync(1, 123, None, reps=2)
        ok = ok and counts == base[0] and dg == base[1]
        s1, counts, dg = run_sync(1, 123, None, reps=2, flush=256 << 20)
        ok = ok and counts == base[0] and dg == base[1]
        accu[v]["sync"].append(s0["update_ms"]); accu[v]["sync_flush"].append(s1["update_ms"])
        for key, fl in (("pipe", 0), ("pipe_flush", 256 << 20)):
            ms, dg, rep = run_pipe(1, 123, None, 3, fl, reps=2)
            ok = ok and dg == base[1]
            accu[v][key].append(ms)
EXTRA.pop("VMAP_UPDATE_VARIANT", None)
for v in (0, 2):
    row = {"update_variant": v, "update_ms_sync": float(np.median(accu[v]["sync"])), "update_ms_sync_flushed": float(np.median(accu[v]["sync_flush"])),
           "pipelined_ms": float(np.median(accu[v]["pipe"])), "pipelined_flushed_ms": float(np.median(accu[v]["pipe_flush"]))}
    out["update_variants"].append(row)
    print(json.dumps(row), flush=True)
print(json.dumps({"all_identical": ok}), flush=True)

if "--oracle" in sys.argv:
    import oracle                                # noqa: E402
    ow = oracle.OracleWorld(resolution=res, sensor_max_range=50.0)
    for s in scans[:2]:
        ow.insertPointcloud(s["T"], s["points"])
    ko, vo = ow.export_leaves()
    po = oracle.pack_keys(ko)
    oo = np.argsort(po)
    w = make(1, 1, None)
    for i in range(2):
        w.insertPointcloudDevice(scans[i]["T"], dev[i], n)
    k, v = w.export_leaves()
    p = oracle.pack_keys(k)
    o = np.argsort(p)
    eq = bool(np.array_equal(p[o], po[oo]) and
              np.array_equal(np.asarray(v, np.float32).view(np.uint32)[o], np.asarray(vo, np.float32).view(np.uint32)[oo]))
    out["oracle_equal"] = eq
    ok = ok and eq
    print(json.dumps({"oracle_equal": eq, "leaves": int(len(p))}), flush=True)
    w.close()
out["all_identical"] = ok
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/ab_helpers.json", "w"), indent=1)
print("AB_OK" if ok else "AB_MISMATCH")
sys.exit(0 if ok else 1)
