#!/usr/bin/env python
"""bench.py -- the headline benchmark of the scan-insertion hot path.

Metric (BASELINE.json): ray-voxel traversals/s (and scans/s) for Velodyne HDL-64-style scans
inserted at 0.1 m resolution, 50 m max range.  One STEP = one scan (133,312 points, about
24.9 M ray-voxel traversals) inserted into the growing map; successive steps use successive
scans of the synthetic corridor trajectory (scenes.trajectory_scan), so every step has new
input and new map blocks.

  python bench.py --gpus N --steps K --warmup W            the CUDA path (this repository)
  python bench.py --impl reference --steps K --warmup W    the reference's CPU algorithm
                                                           (oracle/ restatement, 1 thread)

Per-step timing is done on the device with CUDA events recorded on the library's own stream
(vmap_scan_stats.gpu_ms, bracketing exactly the scan's kernels); a 256 MiB buffer is
overwritten between steps so no step starts with its map blocks in L2.  N > 1: one process per
GPU (torchrun); ONE map, sharded by voxel-block owner over the ranks.  A step is one ROUND of N
scans: rank r generates scan (step*N + r), block records are exchanged all-to-all over NCCL
inside the library, and every shard applies the round's scans in order (DESIGN.md "Multi-GPU").
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from volumetric_mapping_b200 import scenes  # noqa: E402  (numpy only)

RESOLUTION = 0.10
MAX_RANGE = 50.0
METRIC = "ray_voxel_traversals_per_s"
UNIT = "traversals/s"
WORKLOAD = "cfg3: HDL-64-style scan, 64x2083=133312 points, 50 m range, 0.1 m voxels, 1 scan/step"


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def algorithmic_bytes(st, n_points):
    """B_alg = 12 N + 16 (T + O) + 24 U per scan."""
    T, O = st["traversals"], st["occupied_emitted"]
    U = st["unique_free"] + st["unique_occupied"]
    return 12 * n_points + 16 * (T + O) + 24 * U


def walk_bytes(st):
    """The DDA walk's share of B_alg (DESIGN.md "Kernels"): read one 12-byte ray end per cast
    ray, produce one 8-byte key per traversal."""
    return 12 * st["rays_cast"] + 8 * st["traversals"]


# ---------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe) during the timed region
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle (CPU restatement of the reference), single thread
# ---------------------------------------------------------------------------------------------
def oracle_world():
    import oracle  # test infrastructure; legal here (bench.py cpu_baseline / --impl reference)
    oracle.build()
    return oracle.OracleWorld(resolution=RESOLUTION, sensor_max_range=MAX_RANGE)


def decimate(scan, d):
    """Every d-th azimuth firing of the scan (all 64 lasers of the kept firings)."""
    if d <= 1:
        return scan["points"]
    pts = scan["points"].reshape(scenes.HDL64_AZIMUTHS, scenes.HDL64_ELEVATIONS, 3)
    return np.ascontiguousarray(pts[::d]).reshape(-1, 3)


def cpu_baseline_sample(budget_s=20.0):
    """Times the oracle on a bounded sample of the same workload: whole scans of the trajectory
    until about `budget_s` seconds of CPU work are done (at least one scan)."""
    ow = oracle_world()
    t_total, trav, scans = 0.0, 0, 0
    split = [0.0, 0.0, 0.0]
    while scans < 1 or (t_total < budget_s and scans < 4):
        s = scenes.trajectory_scan(scans, resolution=RESOLUTION)
        t0 = time.perf_counter()
        ow.insertPointcloud(s["T"], s["points"])
        t_total += time.perf_counter() - t0
        st = ow.last_scan_stats()
        trav += st["traversals"]
        split[0] += st["t_keyset_s"]; split[1] += st["t_update_s"]; split[2] += st["t_inner_s"]
        scans += 1
    # cfg5 on the CPU: a bounded sample of the same query distribution against this 4-scan map
    q = scenes.cfg5_queries(n_segments=100000, n_points=100000,
                            bounds=((-45.0, 45.0 + 0.5 * scans), (-45.0, 45.0), (-0.5, 4.0)))
    t0 = time.perf_counter()
    ow.getLineStatus(q["segments"])
    t1 = time.perf_counter()
    ow.getCellStatusPoint(q["points"])
    t2 = time.perf_counter()
    return {"value": trav / t_total, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "%d full cfg3 scans of the same trajectory, %.1f s of CPU work "
                      "(keyset %.1f s / updateNode %.1f s / updateInnerOccupancy %.1f s)"
                      % (scans, t_total, split[0], split[1], split[2]),
            "scans_per_s": scans / t_total, "host_cpus": os.cpu_count(),
            "queries": {"segments_per_s": 100000 / (t1 - t0), "points_per_s": 100000 / (t2 - t1),
                        "sample": "100k segments + 100k points vs the map of %d scans" % scans}}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    total = args.steps + args.warmup
    d = max(1, int(np.ceil(total / 12.0)))      # keep the whole run within a few minutes
    ow = oracle_world()
    times, trav = [], 0
    for i in range(total):
        s = scenes.trajectory_scan(i, resolution=RESOLUTION)
        pts = decimate(s, d)
        t0 = time.perf_counter()
        ow.insertPointcloud(s["T"], pts)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
            trav += ow.last_scan_stats()["traversals"]
    t = sum(times)
    value = trav / t
    sample = ("each step = every %d-th azimuth firing (%d of 133312 points) of the step's scan"
              % (d, decimate(scenes.trajectory_scan(0, resolution=RESOLUTION), d).shape[0])
              if d > 1 else "each step = one full cfg3 scan")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t / max(1, len(times)), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64/f32/u16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "resolution": RESOLUTION, "max_range": MAX_RANGE,
                   "path": "oracle/ C++ restatement of OctomapWorld (reference is unbuildable "
                           "here: octomap/PCL/OpenCV/Eigen absent), 1 thread like the reference"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": sample, "host_cpus": os.cpu_count()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "scans_per_s": len(times) / t, "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------
# the CUDA arm
# ---------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    import volumetric_mapping_b200 as vm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    K, W = args.steps, args.warmup
    total = K + W
    # weak scaling: a step is one round of `world` scans of the SAME trajectory and map; rank r
    # generates scan step*world + r, every shard applies the round in scan order.
    scans = [scenes.trajectory_scan(rank + world * i, resolution=RESOLUTION) for i in range(total)]
    n_pts = scans[0]["points"].shape[0]
    dev_pts = [torch.from_numpy(s["points"]).cuda() for s in scans]
    pinned = [torch.from_numpy(s["points"]).pin_memory() for s in scans]
    pinned_np = [p.numpy() for p in pinned]
    T_c = [(C.c_double * 16)(*np.ascontiguousarray(s["T"], dtype=np.float64).reshape(16)) for s in scans]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(kind):
        """kind: 'dev' (inputs resident in HBM) or 'host' (pinned host buffers, H2D inside)."""
        if world > 1:
            from volumetric_mapping_b200 import sharded
            sw = sharded.ShardedOctomapWorld(rank, world, local_rank, batch=args.dist_batch,
                                             resolution=RESOLUTION, sensor_max_range=MAX_RANGE,
                                             initial_blocks=1 << 18)
            w = sw.local
            w.set_debug_l2_flush(256 << 20)   # in-stream L2 flush before every round's apply
        else:
            w = vm.OctomapWorld(resolution=RESOLUTION, sensor_max_range=MAX_RANGE,
                                initial_blocks=1 << 19, device=local_rank)
        launches0 = None
        stats, gpu_ms, wall_ms = [], [], []
        sampler = ClockSampler(local_rank)
        if kind == "dev" and rank == 0:
            sampler.start()          # nvidia-smi needs ~100 ms to come up: start before warm-up
            time.sleep(0.3)
        t_region = None
        for i in range(total):
            if i == W:
                if world > 1:
                    w.dist_flush()
                barrier()
                launches0 = w.kernel_launches()
                t_region = time.perf_counter()
            if world > 1:
                # rounds are pipelined inside the library (exchange + apply of round k overlap the
                # generation of round k+1): the timed region is the K rounds as a whole
                if kind == "dev":
                    w.dist_insert_pointcloud_device(scans[i]["T"], dev_pts[i].data_ptr(), n_pts)
                else:
                    w.dist_insert_pointcloud(scans[i]["T"], pinned_np[i])
                st = w.last_scan_stats()
                if i >= W:
                    stats.append(st)
                continue
            flush.fill_(i & 0xFF)            # evict the map from L2 between steps
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if kind == "dev":
                w.insertPointcloudDevice(scans[i]["T"], dev_pts[i].data_ptr(), n_pts)
            else:
                w.insertPointcloudHostPtr(T_c[i], pinned[i].data_ptr(), n_pts)
            dt = time.perf_counter() - t0
            st = w.last_scan_stats()
            if i >= W:
                stats.append(st)
                gpu_ms.append(st["gpu_ms"])
                wall_ms.append(1e3 * dt)
        if world > 1:
            w.dist_flush()
        barrier()
        if world > 1:
            region_ms = 1e3 * (time.perf_counter() - t_region)
            gpu_ms = wall_ms = [region_ms / K] * K
        clocks = sampler.stop() if (kind == "dev" and rank == 0) else None
        launches = w.kernel_launches() - launches0
        out = {"stats": stats, "gpu_ms": gpu_ms, "wall_ms": wall_ms, "launches": launches,
               "clocks": clocks, "device_bytes": w.device_bytes()}
        if kind == "dev" and world == 1:
            out["queries"] = time_queries(w, total)
        if world > 1:
            sw.close()
        else:
            w.close()
        return out

    def time_queries(w, n_scans):
        """cfg5: 1M getLineStatus segments + 1M getCellStatusPoint points against the map just
        built (device-resident queries, L2 flushed, wall clock around the C ABI call)."""
        q = scenes.cfg5_queries(bounds=((-45.0, 45.0 + 0.5 * n_scans), (-45.0, 45.0), (-0.5, 4.0)))
        seg = torch.from_numpy(q["segments"]).cuda()
        pts = torch.from_numpy(q["points"]).cuda()
        sl = torch.zeros(seg.shape[0], dtype=torch.uint8, device="cuda")
        sp = torch.zeros(pts.shape[0], dtype=torch.uint8, device="cuda")
        best_l, best_p = 1e9, 1e9
        for rep in range(4):
            flush.fill_(rep)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            w.getLineStatusDevice(seg.data_ptr(), seg.shape[0], sl.data_ptr())
            t1 = time.perf_counter()
            flush.fill_(rep + 7)
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            w.getCellStatusPointDevice(pts.data_ptr(), pts.shape[0], sp.data_ptr())
            t3 = time.perf_counter()
            if rep:
                best_l, best_p = min(best_l, t1 - t0), min(best_p, t3 - t2)
        return {"workload": "cfg5: 1M segments (0.5-20 m) + 1M points vs the 0.1 m map of %d scans" % n_scans,
                "segments_per_s": seg.shape[0] / best_l, "points_per_s": pts.shape[0] / best_p,
                "line_ms": 1e3 * best_l, "point_ms": 1e3 * best_p,
                "non_free_fraction": float((sl != 0).float().mean().item())}

    dev = run("dev")
    host = run("host")

    trav = sum(s["traversals"] for s in dev["stats"])
    t_dev = sum(dev["gpu_ms"]) * 1e-3
    t_host = sum(host["wall_ms"]) * 1e-3
    trav_host = sum(s["traversals"] for s in host["stats"])
    if world > 1:
        # whole-job aggregate: all ranks' traversals over the slowest rank's time
        tt = torch.tensor([t_dev, t_host], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        tv = torch.tensor([trav, trav_host], dtype=torch.float64, device="cuda")
        dist.all_reduce(tv, op=dist.ReduceOp.SUM)
        ln = torch.tensor([dev["launches"]], dtype=torch.float64, device="cuda")
        dist.all_reduce(ln, op=dist.ReduceOp.SUM)
        t_dev, t_host = tt.tolist()
        trav, trav_host = tv.tolist()
        launches = int(ln.item())
    else:
        launches = dev["launches"]

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        st = dev["stats"]
        # dominant kernel: the DDA walk (K2).  Average launch duration from CUDA events on the
        # library's stream over the timed region; algorithmic bytes per launch per DESIGN.md.
        walk_ms = float(np.mean([s["walk_ms"] for s in st]))
        walk_b = float(np.mean([walk_bytes(s) for s in st]))
        scan_b = float(np.mean([algorithmic_bytes(s, n_pts) for s in st]))
        scan_ms = float(np.mean([s["gpu_ms"] for s in st]))
        ach = walk_b / (walk_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "K2 = k_checkpoints + k_walk_segments (+ list/locate): DDA walk + key-set build",
                    "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": None, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": walk_b, "avg_launch_ms": walk_ms,
                    "whole_scan": {"algorithmic_bytes": scan_b, "gpu_ms": scan_ms,
                                   "achieved": scan_b / (scan_ms * 1e-3) / 1e9,
                                   "frac": scan_b / (scan_ms * 1e-3) / 1e9 / peak},
                    "stage_ms": {k: float(np.mean([s[k] for s in st]))
                                 for k in ("front_ms", "resolve_ms", "walk_ms", "update_ms")}}
        traffic_file = os.path.join(ROOT, "profiles", "walk_traffic.json")
        if os.path.exists(traffic_file):
            try:
                roofline["traffic"] = json.load(open(traffic_file))["dram_bytes_per_launch"]
            except Exception:
                pass
        cpu = cpu_baseline_sample() if world == 1 and not args.no_cpu_baseline else None
        line = {
            "metric": METRIC, "value": trav / t_dev, "unit": UNIT, "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": 1e3 * t_dev / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64/f32/u16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "resolution": RESOLUTION, "max_range": MAX_RANGE,
                       "points_per_step": n_pts,
                       "traversals_per_step": trav / K / world,
                       "l2": ("256 MiB buffer overwritten between steps (L2 flushed)" if world == 1 else
                              "256 MiB buffer overwritten on the apply stream before every round"),
                       "timing": ("CUDA events on the library stream around each scan's kernels"
                                  if world == 1 else "wall clock over the K pipelined rounds, barrier + "
                                  "device sync on both sides, max over ranks"),
                       "sharding": ("one map sharded by 8^3-block owner over %d ranks; step = "
                                    "round of %d scans, NCCL all-to-all of block records once per "
                                    "%d rounds" % (world, world, args.dist_batch)) if world > 1 else "single GPU"},
            "scans_per_s": K * world / t_dev,
            "e2e": {"value": trav_host / t_host, "unit": UNIT,
                    "h2d_bytes_per_step": n_pts * 12, "d2h_bytes_per_step": 80,
                    "scans_per_s": K * world / t_host,
                    "how": "vmap_insert_pointcloud (C ABI) on pinned host points; wall clock "
                           "around each call (H2D copy, 4 kernels, counter read-back)"},
            "gpu_launches": launches,
            "clocks": dev["clocks"],
            "roofline": roofline,
            "cpu_baseline": cpu,
            "queries": dev.get("queries"),
            "device_bytes": dev["device_bytes"],
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dist-batch", type=int, default=4,
                    help="N > 1: scans per rank that share one exchange round (1..8)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
