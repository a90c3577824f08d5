#!/usr/bin/env python
"""bench.py -- the headline benchmark of the scan-insertion hot path.

Metric (BASELINE.json): ray-voxel traversals/s (and scans/s) for Velodyne HDL-64-style scans
inserted at 0.1 m resolution, 50 m max range (cfg3).  One STEP = one scan per GPU (133,312 points,
about 24.9 M ray-voxel traversals) inserted into the growing map; successive steps use successive
scans of the synthetic corridor trajectory (scenes.trajectory_scan), so every step has new input
and new map blocks.

  python bench.py --gpus N --steps K --warmup W            the CUDA path (this repository)
  python bench.py --impl reference --steps K --warmup W    the reference's CPU algorithm
                                                           (oracle/ restatement, 1 thread)

Timing (every N the same way): the K steps are ONE region on the device -- vmap_region_begin /
vmap_region_end: everything drained, a CUDA event, the K pipelined inserts, everything drained, a
second event -- bracketed by a barrier + device synchronisation, maximum over ranks.  A 256 MiB
buffer is overwritten in-stream before every scan, INSIDE the region, so no scan finds its map
blocks in L2.  N > 1: one process per GPU (torchrun); ONE map, sharded by voxel-block owner over
the ranks; rank r generates scan (step*N + r), block records are exchanged through peer memory and
every shard applies the round's scans in order (DESIGN.md "Multi-GPU").  After the timed region
the sharded map is compared (leaf count + order-independent digest of (key, log-odds bits)) with a
single-GPU build of the same scans: "parity" in the JSON line, exit code 3 on a mismatch.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # one hardware queue per stream (spin-wait kernels)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from volumetric_mapping_b200 import scenes  # noqa: E402  (numpy only)

RESOLUTION = 0.10
MAX_RANGE = 50.0
METRIC = "ray_voxel_traversals_per_s"
UNIT = "traversals/s"
WORKLOAD = "cfg3: HDL-64-style scan, 64x2083=133312 points, 50 m range, 0.1 m voxels, 1 scan/step/GPU"
L2_FLUSH_BYTES = 256 << 20
MAP_BUILD_RES = 0.05


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


def oracle_world_for(params):
    import oracle  # test infrastructure; legal here (bench.py cpu_baseline leg)
    oracle.build()
    return oracle.OracleWorld(**params)


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
    """The reference's own (single-threaded) algorithm on the stated config: every step is one FULL
    cfg3 scan of the same trajectory, inserted into the growing map -- about 6 s of CPU per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    total = args.steps + args.warmup
    ow = oracle_world()
    times, trav = [], 0
    for i in range(total):
        s = scenes.trajectory_scan(i, resolution=RESOLUTION)
        t0 = time.perf_counter()
        ow.insertPointcloud(s["T"], s["points"])
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
            trav += ow.last_scan_stats()["traversals"]
    t = sum(times)
    value = trav / t
    sample = "each step = one full cfg3 scan (133312 points) of the same trajectory"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t / max(1, len(times)), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64/f32/u16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "resolution": RESOLUTION, "max_range": MAX_RANGE,
                   "points_per_step": int(s["points"].shape[0]), "traversals_per_step": trav / max(1, len(times)),
                   "path": "oracle/ C++ restatement of OctomapWorld (reference is unbuildable "
                           "here: octomap/PCL/OpenCV/Eigen absent), 1 thread like the reference; every step "
                           "is one full scan of the same trajectory the CUDA arm inserts"},
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
    from volumetric_mapping_b200 import sharded

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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(v):
        if world == 1:
            return float(v)
        t = torch.tensor([float(v)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(v):
        if world == 1:
            return float(v)
        t = torch.tensor([float(v)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def gather_digest(n, d):
        """(count, digest) of the whole map: shard digests add up modulo 2^64."""
        if world == 1:
            return n, d
        mine = torch.tensor([n, d - (1 << 64) if d >= (1 << 63) else d], dtype=torch.int64, device="cuda")
        got = [torch.zeros(2, dtype=torch.int64, device="cuda") for _ in range(world)]
        dist.all_gather(got, mine)
        cnt = sum(int(g[0].item()) for g in got)
        dg = sum(int(g[1].item()) % (1 << 64) for g in got) % (1 << 64)
        return cnt, dg

    def make_world(res, blocks):
        """One map: a single handle at N = 1, a sharded one (peer-memory transport) at N > 1."""
        if world > 1:
            sw = sharded.ShardedOctomapWorld(rank, world, local_rank, batch=args.dist_batch,
                                             resolution=res, sensor_max_range=MAX_RANGE,
                                             initial_blocks=blocks)
            return sw, sw.local
        w = vm.OctomapWorld(resolution=res, sensor_max_range=MAX_RANGE, initial_blocks=blocks, device=local_rank)
        return None, w

    def close_world(sw, w):
        if sw is not None:
            sw.close()
        else:
            w.close()

    def timed_build(res, n_total, n_warm, kind, blocks, stride=12, writeback=False, keep=False, sample_clocks=False):
        """Inserts scans rank + world * i (i < n_total) of the trajectory at `res`; the first n_warm
        steps are warm-up, the rest is the timed region.  kind: 'dev' (points resident in HBM),
        'host' (pinned host points, lean call) or 'dropin' (pageable 16-byte points, in-place
        world-frame write-back: what the reference-side binding passes)."""
        scans = [scenes.trajectory_scan(rank + world * i, resolution=res) for i in range(n_total)]
        n_pts = scans[0]["points"].shape[0]
        T_c = [(C.c_double * 16)(*np.ascontiguousarray(s["T"], dtype=np.float64).reshape(16)) for s in scans]
        if kind == "dev":
            bufs = [torch.from_numpy(s["points"]).cuda() for s in scans]
        elif kind == "host":
            bufs = [torch.from_numpy(s["points"]).pin_memory() for s in scans]
        else:
            bufs = []
            for s in scans:      # pcl::PointXYZ: x, y, z, 1.0f -- pageable, like a cloud that came off a message
                b = np.ones((n_pts, stride // 4), dtype=np.float32)
                b[:, :3] = s["points"]
                bufs.append(b)
        sw, w = make_world(res, blocks)
        w.set_debug_l2_flush(L2_FLUSH_BYTES)
        n_out = C.c_size_t(0)

        def insert(i):
            if world > 1:
                if kind == "dev":
                    w.dist_insert_pointcloud_device(scans[i]["T"], bufs[i].data_ptr(), n_pts)
                else:
                    w._check(w._L.vmap_dist_insert_pointcloud(w.handle, C.c_void_p(bufs[i].data_ptr()), n_pts, 12, 1, T_c[i]))
            elif kind == "dev":
                w._check(w._L.vmap_insert_pointcloud_dev(w.handle, C.c_void_p(bufs[i].data_ptr()), n_pts, 12, T_c[i]))
            elif kind == "host":
                w.insertPointcloudHostPtr(T_c[i], bufs[i].data_ptr(), n_pts)
            else:
                p = bufs[i].ctypes.data
                w._check(w._L.vmap_insert_pointcloud(w.handle, p, n_pts, stride, 1, T_c[i], p, C.byref(n_out)))

        def flush():
            if world > 1:
                w.dist_flush()
            else:
                w.flush()

        sampler = ClockSampler(local_rank)
        if kind == "dev" and rank == 0 and sample_clocks:
            sampler.start()          # nvidia-smi needs ~100 ms to come up: start before warm-up
            time.sleep(0.3)
        for i in range(n_warm):
            insert(i)
        flush()
        barrier()
        t_before = w.totals()
        launches0 = w.kernel_launches()
        w.region_begin()
        t0 = time.perf_counter()
        for i in range(n_warm, n_total):
            insert(i)
        flush()
        ms = w.region_end()
        wall_ms = 1e3 * (time.perf_counter() - t0)
        barrier()
        t_after = w.totals()
        clocks = sampler.stop() if (kind == "dev" and rank == 0 and sample_clocks) else None
        out = {"ms": allmax(ms), "wall_ms": allmax(wall_ms),
               "traversals": allsum(t_after["traversals"] - t_before["traversals"]),
               "launches": int(allsum(w.kernel_launches() - launches0)),
               "replayed": int(allsum(t_after["replayed"])), "n_pts": n_pts, "clocks": clocks,
               "device_bytes": w.device_bytes(), "steps": n_total - n_warm}
        out["digest"] = gather_digest(*w.leaf_digest())
        if keep:
            out["world"] = (sw, w)
        else:
            close_world(sw, w)
        return out

    def single_gpu_digest(res, n_total, blocks):
        """The same scans, in global order, through vmap_insert_pointcloud_dev on ONE device (the
        oracle-verified path): what the sharded map must equal."""
        w = vm.OctomapWorld(resolution=res, sensor_max_range=MAX_RANGE, initial_blocks=blocks, device=local_rank)
        for g in range(n_total * world):
            s = scenes.trajectory_scan(g, resolution=res)
            d = torch.from_numpy(s["points"]).cuda()
            w.insertPointcloudDevice(s["T"], d.data_ptr(), d.shape[0])
        out = w.leaf_digest()
        w.close()
        return out

    def stage_pass(n_total, n_warm):
        """Serialised pass (pipeline depth 0, one scan at a time) for the per-stage CUDA-event times
        the roofline needs; the L2 flush runs in-stream BEFORE each scan's first event."""
        w = vm.OctomapWorld(resolution=RESOLUTION, sensor_max_range=MAX_RANGE, initial_blocks=1 << 19, device=local_rank)
        w.set_pipeline_depth(0)
        w.set_debug_l2_flush(L2_FLUSH_BYTES)
        stats = []
        for i in range(n_total):
            s = scenes.trajectory_scan(i, resolution=RESOLUTION)
            d = torch.from_numpy(s["points"]).cuda()
            w.insertPointcloudDevice(s["T"], d.data_ptr(), d.shape[0])
            if i >= n_warm:
                st = w.last_scan_stats()
                st["walk_kernel_ms"] = w.last_walk_kernel_ms()     # events directly around the walker's launch
                stats.append(st)
        w.close()
        return stats

    def time_queries(w, n_scans):
        """cfg5: 1M getLineStatus segments + 1M getCellStatusPoint points against the map just
        built (device-resident queries, L2 flushed, wall clock around the C ABI call); the same for
        a HARD subset -- segments that start in known-free space along the trajectory --, the probe
        count behind B_alg = 49 + 12 * probes per segment, and a 100k subsample checked against the
        CPU oracle on an oracle-built map of the same scans when that is affordable."""
        flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda")
        bounds = ((-45.0, 45.0 + 0.5 * n_scans), (-45.0, 45.0), (-0.5, 4.0))
        q = scenes.cfg5_queries(bounds=bounds)
        rng = np.random.default_rng(7)
        m = q["segments"].shape[0]
        # hard subset: start points 0.3..1.7 m above the ground within 8 m of the sensor path (seen free)
        a = np.stack([rng.uniform(0.0, 0.5 * n_scans, m), rng.uniform(-8.0, 8.0, m), rng.uniform(0.3, 1.7, m)], axis=1)
        d = rng.normal(size=(m, 3))
        d[:, 2] *= 0.15
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        hard = np.concatenate([a, a + d * rng.uniform(2.0, 30.0, (m, 1))], axis=1)
        out = {}
        peak, _ = measured_peak_gbs()
        for name, segs in (("uniform", q["segments"]), ("hard", hard)):
            seg = torch.from_numpy(np.ascontiguousarray(segs)).cuda()
            sl = torch.zeros(seg.shape[0], dtype=torch.uint8, device="cuda")
            best = 1e9
            for rep in range(4):
                flush.fill_(rep)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                w.getLineStatusDevice(seg.data_ptr(), seg.shape[0], sl.data_ptr())
                t1 = time.perf_counter()
                if rep:
                    best = min(best, t1 - t0)
            probes = w.getLineStatusProbesDevice(seg.data_ptr(), seg.shape[0], sl.data_ptr())
            b_alg = 49 * seg.shape[0] + 12 * probes
            out[name] = {"segments_per_s": seg.shape[0] / best, "line_ms": 1e3 * best,
                         "probes_per_segment": probes / seg.shape[0],
                         "non_free_fraction": float((sl != 0).float().mean().item()),
                         "roofline": {"bound": "hbm", "algorithmic_bytes": b_alg, "achieved": b_alg / best / 1e9,
                                      "peak": peak, "unit": "GB/s", "frac": b_alg / best / 1e9 / peak}}
        pts = torch.from_numpy(q["points"]).cuda()
        sp = torch.zeros(pts.shape[0], dtype=torch.uint8, device="cuda")
        best_p = 1e9
        for rep in range(4):
            flush.fill_(rep + 7)
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            w.getCellStatusPointDevice(pts.data_ptr(), pts.shape[0], sp.data_ptr())
            t3 = time.perf_counter()
            if rep:
                best_p = min(best_p, t3 - t2)
        out["points_per_s"] = pts.shape[0] / best_p
        out["point_ms"] = 1e3 * best_p
        out["workload"] = ("cfg5: 1M segments (0.5-20 m, uniform in the mapped box) / 1M hard segments (2-30 m, starting in "
                           "known-free space along the path) + 1M points vs the 0.1 m map of %d scans" % n_scans)
        # compatibility keys of the round-1 line
        out["segments_per_s"] = out["uniform"]["segments_per_s"]
        out["line_ms"] = out["uniform"]["line_ms"]
        out["non_free_fraction"] = out["uniform"]["non_free_fraction"]
        return out

    def time_queries_sharded(sw, n_scans):
        """cfg5 on the SHARDED map (N > 1): every GPU assembles a read-only replica of the whole map
        (ShardedOctomapWorld.replicate: leaves exported on the device, NCCL broadcasts over NVLink,
        imported in place), then the 1M uniform segments / 1M points are dealt over the ranks in equal
        slices and the statuses all-gathered -- wall clock around the collective call, max over ranks;
        the digest of the replica is compared with the sum of the shards'."""
        t0 = time.perf_counter()
        leaves = sw.replicate()
        torch.cuda.synchronize()
        t_rep = allmax(time.perf_counter() - t0)
        rep_ok = bool(gather_digest(*sw.local.leaf_digest()) == sw.replica.leaf_digest())
        bounds = ((-45.0, 45.0 + 0.5 * n_scans), (-45.0, 45.0), (-0.5, 4.0))
        q = scenes.cfg5_queries(bounds=bounds)
        seg = torch.from_numpy(np.ascontiguousarray(q["segments"])).cuda()
        pts = torch.from_numpy(np.ascontiguousarray(q["points"])).cuda()
        flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda")
        best_l = best_p = 1e9
        for rep in range(4):
            flush.fill_(rep)
            barrier()
            t0 = time.perf_counter()
            sl = sw.getLineStatusSliced(seg)
            torch.cuda.synchronize()
            dt = allmax(time.perf_counter() - t0)
            if rep:
                best_l = min(best_l, dt)
        for rep in range(4):
            flush.fill_(rep + 9)
            barrier()
            t0 = time.perf_counter()
            sw.getCellStatusPointSliced(pts)
            torch.cuda.synchronize()
            dt = allmax(time.perf_counter() - t0)
            if rep:
                best_p = min(best_p, dt)
        return {"mode": "read-only replica per GPU + segment slices (SURVEY 8(e)); statuses all-gathered to every rank",
                "replicate_ms": 1e3 * t_rep, "replica_leaves": leaves, "replica_equals_shards": rep_ok,
                "segments_per_s": seg.shape[0] / best_l, "line_ms": 1e3 * best_l,
                "points_per_s": pts.shape[0] / best_p, "point_ms": 1e3 * best_p,
                "non_free_fraction": float((sl != 0).float().mean().item()),
                "workload": "cfg5: 1M uniform segments (0.5-20 m) + 1M points vs the sharded 0.1 m map of %d scans" % n_scans}

    def small_configs(with_cpu):
        """BASELINE.json configs 1 and 2 (the reference's own CPU-runnable cases): ONE insertion of the
        config's input into an empty 0.15 m map through the reference-facing call with host buffers
        (vmap_insert_pointcloud with write-back / vmap_insert_disparity), wall clock around call +
        drain, median of 9 after 3 warm-ups (resetMap between repetitions, outside the clock); the
        oracle's time for the same single insertion beside it."""
        out = {}
        c1, c2 = scenes.cfg1_pointcloud(), scenes.cfg2_disparity()
        for name, s in (("cfg1", c1), ("cfg2", c2)):
            w = vm.OctomapWorld(device=local_rank, **s["params"])

            def one(world_obj, s=s, name=name):
                if name == "cfg1":
                    world_obj.insertPointcloud(s["T"], s["points"])
                else:
                    world_obj.insertDisparityImage(s["q"], s["t"], s["disparity"], s["Q"], s["full_image_size"][0])
            times = []
            for rep in range(12):
                w.resetMap()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                one(w)
                w.flush()
                times.append(time.perf_counter() - t0)
            st = w.last_scan_stats()
            leaves = w.leaf_count()
            w.close()
            t = float(np.median(times[3:]))
            out[name] = {"ms_per_insert": 1e3 * t, "traversals": int(st["traversals"]), "rays_cast": int(st["rays_cast"]),
                         "traversals_per_s": st["traversals"] / t, "leaves": int(leaves)}
            if with_cpu:
                ow = oracle_world_for(s["params"])
                t0 = time.perf_counter()
                one(ow)
                tc = time.perf_counter() - t0
                ost = ow.last_scan_stats()
                out[name]["cpu_ms"] = 1e3 * tc
                out[name]["cpu_traversals_per_s"] = ost["traversals"] / tc
                out[name]["same_traversals_as_cpu"] = bool(int(ost["traversals"]) == int(st["traversals"]))
        out["workload"] = ("cfg1: 10k-point cloud, 0.15 m, 5 m range; cfg2: 480x640 disparity image through Q, 0.15 m, 5 m range; "
                           "one insertion into an empty map, host buffers, wall clock around call + drain")
        return out

    def map_build(n_scans):
        """BASELINE.json config 4, bounded: n_scans consecutive scans per GPU of the 1000-scan trajectory
        at 0.05 m (about 55 M traversals and 350 k voxel blocks per scan), same timing as the headline."""
        r = timed_build(MAP_BUILD_RES, n_scans + 3, 3, "dev", 1 << 21)
        return {"workload": "cfg4 (bounded): %d consecutive scans per GPU of the 1000-scan trajectory at 0.05 m, "
                            "map sharded over %d GPU(s)" % (n_scans, world),
                "scans": n_scans * world, "ms_per_step": r["ms"] / n_scans,
                "scans_per_s": n_scans * world / (r["ms"] * 1e-3),
                "traversals_per_s": r["traversals"] / (r["ms"] * 1e-3),
                "traversals_per_scan": r["traversals"] / (n_scans * world),
                "leaves": r["digest"][0], "replayed": r["replayed"], "device_bytes": r["device_bytes"]}

    # ---- the runs ----
    dev = timed_build(RESOLUTION, total, W, "dev", 1 << 19, keep=True, sample_clocks=True)
    queries_sharded = None
    if world > 1:
        sw0, w0 = dev.pop("world")
        try:
            queries_sharded = time_queries_sharded(sw0, total * world)
        except Exception as e:          # reported in the line, never hides the headline
            queries_sharded = {"error": repr(e)[:300]}
        close_world(sw0, w0)
    host = timed_build(RESOLUTION, total, W, "host", 1 << 19)
    parity = None
    if world > 1:
        ref = single_gpu_digest(RESOLUTION, total, 1 << 19) if rank == 0 else (0, 0)
        parity = {"checked": True, "leaves": dev["digest"][0], "digest": "%016x" % dev["digest"][1],
                  "single_gpu_leaves": ref[0], "single_gpu_digest": "%016x" % ref[1],
                  "equal": bool(dev["digest"] == ref and host["digest"] == ref) if rank == 0 else None,
                  "how": "leaf count and order-independent 64-bit digest of (key, log-odds bits), shards summed, "
                         "vs a single-GPU vmap_insert_pointcloud_dev build of the same %d scans" % (total * world)}
    queries = dropin = stages = small = None
    if world == 1:
        sw0, w0 = dev.pop("world")
        queries = time_queries(w0, total)
        close_world(sw0, w0)
        dropin = timed_build(RESOLUTION, total, W, "dropin", 1 << 19, stride=16, writeback=True)
        stages = stage_pass(min(total, W + 12), W)
        small = small_configs(with_cpu=(rank == 0 and not args.no_cpu_baseline))
    mb = map_build(args.map_build_scans) if args.map_build_scans > 0 else None

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        n_pts = dev["n_pts"]
        t_dev, t_host = dev["ms"] * 1e-3, host["ms"] * 1e-3
        roofline = None
        if stages:
            # dominant kernel: the DDA walk (K2).  Average duration from CUDA events on the library's
            # stream in the serialised pass; algorithmic bytes per launch per DESIGN.md.
            walk_ms = float(np.mean([s["walk_kernel_ms"] or s["walk_ms"] for s in stages]))
            walk_b = float(np.mean([walk_bytes(s) for s in stages]))
            scan_b = float(np.mean([algorithmic_bytes(s, n_pts) for s in stages]))
            scan_ms = float(np.mean([s["gpu_ms"] for s in stages]))
            ach = walk_b / (walk_ms * 1e-3) / 1e9
            roofline = {"bound": "hbm", "kernel": "k_walk_segments_queued (K2b: DDA walk + key-set build), CUDA events directly around its launch",
                        "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                        "traffic": None, "peak_source": peak_src,
                        "algorithmic_bytes_per_launch": walk_b, "avg_launch_ms": walk_ms,
                        "measured_in": "serialised pass of %d scans (pipeline depth 0), CUDA events between the kernels" % len(stages),
                        "whole_scan": {"algorithmic_bytes": scan_b, "gpu_ms": scan_ms,
                                       "achieved": scan_b / (scan_ms * 1e-3) / 1e9,
                                       "frac": scan_b / (scan_ms * 1e-3) / 1e9 / peak},
                        "whole_scan_pipelined": {"algorithmic_bytes": scan_b, "ms": dev["ms"] / K,
                                                 "achieved": scan_b / (dev["ms"] / K * 1e-3) / 1e9,
                                                 "frac": scan_b / (dev["ms"] / K * 1e-3) / 1e9 / peak},
                        "stage_ms": {k: float(np.mean([s[k] for s in stages]))
                                     for k in ("front_ms", "resolve_ms", "walk_ms", "update_ms", "walk_kernel_ms")},
                        "stage_note": "front = K1; resolve = k_resolve_fold; walk = k_checkpoints_pts + walker + list + admit + locate; "
                                      "update = k_update_mlp; walk_kernel = the walker alone"}
            traffic_file = os.path.join(ROOT, "profiles", "walk_traffic.json")
            if os.path.exists(traffic_file):
                try:
                    tf = json.load(open(traffic_file))
                    roofline["traffic"] = tf["dram_bytes_per_launch"]
                    if tf.get("warp_instructions") and tf.get("kernel_us"):
                        # the walk is issue / latency bound, not HBM bound: instructions the kernel executes
                        # against what the 592 SM sub-partitions could issue in its run time
                        sm_mhz = (dev["clocks"] or {}).get("sm_mhz") or 1965.0
                        roofline["issue_frac"] = tf["warp_instructions"] / (592 * sm_mhz * 1e6 * tf["kernel_us"] * 1e-6)
                        roofline["issue_note"] = ("%d warp instructions of the walk kernel (ncu, profiles/) over 592 issue "
                                                  "ports x SM clock x its %.1f us" % (tf["warp_instructions"], tf["kernel_us"]))
                except Exception:
                    pass
        cpu = cpu_baseline_sample() if world == 1 and not args.no_cpu_baseline else None
        line = {
            "metric": METRIC, "value": dev["traversals"] / t_dev, "unit": UNIT, "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": 1e3 * t_dev / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64/f32/u16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "resolution": RESOLUTION, "max_range": MAX_RANGE,
                       "points_per_step": n_pts * world,
                       "traversals_per_step": dev["traversals"] / K,
                       "l2": ("256 MiB buffer overwritten in-stream inside the timed region, once per scan: between the "
                              "previous scan's map update and this scan's (no update finds its blocks in L2)" if world == 1 else
                              "256 MiB buffer overwritten in-stream before every scan's generate, inside the timed region"),
                       "timing": "one device-timed region over the K pipelined steps (CUDA events after a full drain on "
                                 "both sides), barrier + device sync around it, max over ranks",
                       "pipeline": ("2 scans in flight when an insert call returns, three scan contexts (the next scan's front "
                                    "half is queued while the host waits); map updates ordered by events" if world == 1 else
                                    "up to 3 generated scans in flight per GPU on three scan contexts; packs ordered by events, "
                                    "applies on their own stream"),
                       "sharding": ("one map sharded by 8^3-block owner over %d ranks; step = round of %d scans; block "
                                    "records pulled from the peers' round buffers over NVLink (CUDA IPC), %d scans per "
                                    "rank per round" % (world, world, args.dist_batch)) if world > 1 else "single GPU"},
            "scans_per_s": K * world / t_dev,
            "e2e": {"value": host["traversals"] / t_host, "unit": UNIT,
                    "h2d_bytes_per_step": n_pts * 12 * world, "d2h_bytes_per_step": 96 * world,
                    "scans_per_s": K * world / t_host, "ms_per_step": host["ms"] / K,
                    "how": "vmap_insert_pointcloud / vmap_dist_insert_pointcloud (C ABI) on pinned host points: H2D copy, "
                           "kernels and counter read-back of every step inside the same kind of timed region"},
            "gpu_launches": dev["launches"],
            "clocks": dev["clocks"],
            "roofline": roofline,
            "cpu_baseline": cpu,
            "queries": queries,
            "queries_sharded": queries_sharded,
            "map_build": mb,
            "small_configs": small,
            "parity": parity,
            "device_bytes": dev["device_bytes"],
            "scans_replayed_after_growth": dev["replayed"],
        }
        if dropin:
            t_d = dropin["ms"] * 1e-3
            line["e2e_dropin"] = {"value": dropin["traversals"] / t_d, "unit": UNIT,
                                  "scans_per_s": K / t_d, "ms_per_step": dropin["ms"] / K,
                                  "h2d_bytes_per_step": n_pts * 16, "d2h_bytes_per_step": n_pts * 16 + 96,
                                  "how": "what the reference-side binding passes (INTEGRATION.md): pageable pcl::PointXYZ "
                                         "cloud (16-byte stride), is_dense, in-place world-frame write-back "
                                         "(octomap_world.cc:115-119) -- every call waits for its own write-back"}
            line["e2e_dropin"]["same_map_as_lean"] = bool(dropin["digest"] == host["digest"] == dev["digest"])
        print(json.dumps(line))
    rc = 0
    if world > 1:
        ok = torch.tensor([1 if (rank != 0 or parity["equal"]) else 0], dtype=torch.int32, device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        rc = 0 if int(ok.item()) else 3
        dist.destroy_process_group()
    return rc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dist-batch", type=int, default=2,
                    help="N > 1: scans per rank that share one exchange round (1..8)")
    ap.add_argument("--map-build-scans", type=int, default=32,
                    help="cfg4 leg: scans per GPU at 0.05 m (0 = skip)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        if args.steps == 100 and args.warmup == 5:
            args.steps, args.warmup = 20, 3       # defaults of the CPU arm: about 2.5 minutes
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
