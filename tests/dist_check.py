"""torchrun worker for test_nccl_two_ranks_if_two_gpus: builds a sharded map with one process per
GPU (peer-memory transport by default, VMAP_DIST_TRANSPORT=nccl for the send/recv one) and checks
the union of the shards against the oracle (rank 0 gathers the leaves)."""
import os
import sys

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # one hardware queue per stream (spin-wait kernels)

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from volumetric_mapping_b200 import scenes, sharded  # noqa: E402


def main():
    import faulthandler
    faulthandler.dump_traceback_later(75, exit=True)   # a stuck collective must not eat the GPU budget
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    params = dict(resolution=0.2, sensor_max_range=50.0)
    batch = int(os.environ.get("DIST_BATCH", "1"))
    n_scans = 3 * world + 1
    w = sharded.ShardedOctomapWorld(rank, world, local, batch=batch, initial_blocks=64, **params)
    for i in sharded.scans_of_rank(n_scans, rank, world):
        if i >= 0:
            s = scenes.trajectory_scan(2 * i, resolution=0.2)
            d = torch.from_numpy(s["points"]).cuda()
            w.insert_round_device(s["T"], d.data_ptr(), d.shape[0])
        else:
            w.insert_round_device(None, 0, 0)
    w.local.dist_flush()
    k, v = w.local.export_leaves()
    assert np.all(sharded.owner_of_keys(k, world) == rank)
    # collective queries
    q = scenes.cfg5_queries(n_segments=50000, n_points=50000)
    pts = torch.from_numpy(q["points"]).cuda()
    seg = torch.from_numpy(q["segments"]).cuda()
    sp = torch.zeros(pts.shape[0], dtype=torch.uint8, device="cuda")
    sl = torch.zeros(seg.shape[0], dtype=torch.uint8, device="cuda")
    w.local.dist_query_points_device(pts.data_ptr(), pts.shape[0], sp.data_ptr())
    w.local.dist_query_lines_device(seg.data_ptr(), seg.shape[0], sl.data_ptr())
    # the sliced query mode: a read-only replica of the whole map on every GPU, queries dealt over the ranks
    n_rep = w.replicate()
    sp2 = w.getCellStatusPointSliced(pts)
    sl2 = w.getLineStatusSliced(seg)
    rep_n, rep_d = w.replica.leaf_digest()
    # gather leaves on rank 0
    packed = torch.from_numpy(oracle.pack_keys(k).astype(np.int64)).cuda()
    vals = torch.from_numpy(v.view(np.int32).copy()).cuda()
    sizes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([packed.shape[0]], dtype=torch.int64, device="cuda"))
    mx = int(max(int(s.item()) for s in sizes))
    pk = torch.zeros(mx, dtype=torch.int64, device="cuda"); pk[: packed.shape[0]] = packed
    vl = torch.zeros(mx, dtype=torch.int32, device="cuda"); vl[: vals.shape[0]] = vals
    gp = [torch.zeros(mx, dtype=torch.int64, device="cuda") for _ in range(world)]
    gv = [torch.zeros(mx, dtype=torch.int32, device="cuda") for _ in range(world)]
    dist.all_gather(gp, pk)
    dist.all_gather(gv, vl)
    ok = True
    if rank == 0:
        ow = oracle.OracleWorld(**params)
        ow.set_faithful_inner_update(False)
        for i in range(n_scans):
            s = scenes.trajectory_scan(2 * i, resolution=0.2)
            ow.insertPointcloud(s["T"], s["points"])
        allk = np.concatenate([g[: int(n.item())].cpu().numpy() for g, n in zip(gp, sizes)]).astype(np.uint64)
        allv = np.concatenate([g[: int(n.item())].cpu().numpy() for g, n in zip(gv, sizes)]).view(np.uint32)
        o = np.argsort(allk)
        k0, v0 = ow.export_leaves()
        p0 = oracle.pack_keys(k0); o0 = np.argsort(p0)
        ok = (np.array_equal(allk[o], p0[o0]) and np.array_equal(allv[o], v0.view(np.uint32)[o0])
              and np.array_equal(sp.cpu().numpy(), ow.getCellStatusPoint(q["points"]))
              and np.array_equal(sl.cpu().numpy(), ow.getLineStatus(q["segments"]))
              and n_rep == allk.shape[0] and (rep_n, rep_d) == sharded.leaf_digest(k0, v0)
              and np.array_equal(sp2.cpu().numpy(), sp.cpu().numpy())
              and np.array_equal(sl2.cpu().numpy(), sl.cpu().numpy()))
        print("DIST_CHECK_OK" if ok else "DIST_CHECK_MISMATCH", allk.shape[0], flush=True)
    dist.barrier()
    w.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
