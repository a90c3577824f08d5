"""BASELINE.json config 4 in full: the 1000-scan corridor trajectory (HDL-64-style scans, 0.5 m per
scan) inserted at 0.05 m resolution, 50 m range, on 1..8 GPUs.

  python tools/map_build_bench.py [--scans 1000] [--res 0.05]                          one GPU
  python -m torch.distributed.run --nproc-per-node N ... tools/map_build_bench.py      N GPUs, one map
                                                                                        sharded by block owner

Timing as in bench.py: all scans resident in HBM, ONE device-timed region over the pipelined inserts
(vmap_region_begin / _end, full drain on both sides), barrier + synchronize around it, maximum over
ranks.  N > 1: rank r inserts scans r, r + N, ... (collective rounds); afterwards the sharded map's
leaf count and digest are compared with a single-GPU build of the same scans made on rank 0.
Prints one JSON line (rank 0).
"""
import argparse
import json
import os
import sys
import time

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from volumetric_mapping_b200 import scenes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scans", type=int, default=1000)
    ap.add_argument("--res", type=float, default=0.05)
    ap.add_argument("--batch", type=int, default=2, help="N > 1: scans per rank per exchange round")
    ap.add_argument("--blocks", type=int, default=1 << 23, help="initial voxel-block capacity per GPU")
    ap.add_argument("--warmup", type=int, default=8, help="scans (per GPU) inserted before the timed region")
    ap.add_argument("--no-check", action="store_true")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    import volumetric_mapping_b200 as vm
    from volumetric_mapping_b200 import sharded

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce(v, op):
        if world == 1:
            return float(v)
        t = torch.tensor([float(v)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    rounds = (args.scans + world - 1) // world
    mine = [k * world + rank for k in range(rounds)]          # index >= scans: no scan in that round
    t0 = time.time()
    scans = {i: scenes.trajectory_scan(i, resolution=args.res) for i in mine if i < args.scans}
    dev = {i: torch.from_numpy(s["points"]).cuda() for i, s in scans.items()}
    t_gen = time.time() - t0
    if world > 1:
        sw = sharded.ShardedOctomapWorld(rank, world, local, batch=args.batch, resolution=args.res,
                                         sensor_max_range=50.0, initial_blocks=args.blocks)
        w = sw.local
    else:
        sw, w = None, vm.OctomapWorld(resolution=args.res, sensor_max_range=50.0, initial_blocks=args.blocks, device=local)

    def insert(i):
        if i < args.scans:
            s, d = scans[i], dev[i]
            if world > 1:
                w.dist_insert_pointcloud_device(s["T"], d.data_ptr(), d.shape[0])
            else:
                w.insertPointcloudDevice(s["T"], d.data_ptr(), d.shape[0])
        elif world > 1:
            w.dist_insert_pointcloud_device(np.eye(4), 0, 0)

    def flush():
        if world > 1:
            w.dist_flush()
        else:
            w.flush()

    n_warm = min(args.warmup, max(rounds - 1, 0))
    for i in mine[:n_warm]:
        insert(i)
    flush()
    barrier()
    before = w.totals()
    w.region_begin()
    t0 = time.perf_counter()
    for i in mine[n_warm:]:
        insert(i)
    flush()
    ms = w.region_end()
    wall = time.perf_counter() - t0
    barrier()
    after = w.totals()
    ms = reduce(ms, dist.ReduceOp.MAX if world > 1 else None)
    wall = reduce(wall, dist.ReduceOp.MAX if world > 1 else None)
    trav = reduce(after["traversals"] - before["traversals"], dist.ReduceOp.SUM if world > 1 else None)
    timed_scans = reduce(sum(1 for i in mine[n_warm:] if i < args.scans), dist.ReduceOp.SUM if world > 1 else None)
    replayed = reduce(after["replayed"], dist.ReduceOp.SUM if world > 1 else None)
    n, d = w.leaf_digest()
    dev_bytes = reduce(w.device_bytes(), dist.ReduceOp.MAX if world > 1 else None)
    if world > 1:
        mine_t = torch.tensor([n, d - (1 << 64) if d >= (1 << 63) else d], dtype=torch.int64, device="cuda")
        got = [torch.zeros(2, dtype=torch.int64, device="cuda") for _ in range(world)]
        dist.all_gather(got, mine_t)
        n = sum(int(g[0].item()) for g in got)
        d = sum(int(g[1].item()) % (1 << 64) for g in got) % (1 << 64)
    parity = None
    if world > 1 and not args.no_check:
        ref = (0, 0)
        if rank == 0:
            sw1 = vm.OctomapWorld(resolution=args.res, sensor_max_range=50.0, initial_blocks=args.blocks * 2, device=local)
            for g in range(args.scans):
                s = scenes.trajectory_scan(g, resolution=args.res)
                p = torch.from_numpy(s["points"]).cuda()
                sw1.insertPointcloudDevice(s["T"], p.data_ptr(), p.shape[0])
            ref = sw1.leaf_digest()
            sw1.close()
        parity = {"checked": True, "leaves": n, "digest": "%016x" % d, "single_gpu_leaves": ref[0],
                  "single_gpu_digest": "%016x" % ref[1], "equal": bool((n, d) == ref) if rank == 0 else None}
    if rank == 0:
        print(json.dumps({
            "workload": "cfg4: %d scans of the corridor trajectory (64x2083 points each), %.3f m voxels, 50 m range; "
                        "%d warm-up scans per GPU outside the timed region" % (args.scans, args.res, n_warm),
            "n_gpus": world, "timed_scans": int(timed_scans), "ms": ms, "host_wall_ms": 1e3 * wall,
            "scans_per_s": timed_scans / (ms * 1e-3), "traversals_per_s": trav / (ms * 1e-3),
            "traversals_per_scan": trav / max(timed_scans, 1), "leaves": n, "scans_replayed_after_growth": int(replayed),
            "device_bytes_max_rank": int(dev_bytes), "scan_synthesis_s": t_gen, "dist_batch": args.batch if world > 1 else None,
            "parity": parity}), flush=True)
    rc = 0
    if world > 1:
        ok = torch.tensor([1 if (parity is None or rank != 0 or parity["equal"]) else 0], dtype=torch.int32, device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        rc = 0 if int(ok.item()) else 3
        sw.close()
        dist.destroy_process_group()
    else:
        w.close()
    return rc


if __name__ == "__main__":
    sys.exit(main())
