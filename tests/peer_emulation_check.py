"""Peer-memory transport of the sharded build on ONE GPU: G handles of this process stand for G
ranks (vmap_dist_attach_local: mailboxes shared by plain pointers instead of CUDA IPC), each driven
by its own host thread.  Everything
else -- pipelined generate on two scan contexts, packing into round buffers, publish / collect /
consumed flags, applies reading the peers' round buffers, batches, ragged last round, growth of
table and pool between rounds -- is the code the multi-process build runs.
Bar: the union of the shards equals the CPU oracle's map bit for bit (order of the clamped updates:
octomap_world.cc:244-262).       usage: peer_emulation_check.py G batch n_scans [resolution]"""
import os
import sys
import threading

# several spinning one-warp kernels of different handles are resident at once: every stream needs
# its own hardware queue, or a spinner could sit in front of the kernel it is waiting for
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
import volumetric_mapping_b200 as vm  # noqa: E402
from volumetric_mapping_b200 import scenes, sharded  # noqa: E402


def main():
    G, batch, n_scans = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    res = float(sys.argv[4]) if len(sys.argv) > 4 else 0.2
    params = dict(resolution=res, sensor_max_range=50.0)
    shards = [vm.OctomapWorld(initial_blocks=64, **params) for _ in range(G)]
    vm.OctomapWorld.dist_attach_local(shards, records_per_round=1 << 19)
    for w in shards:
        w.dist_set_batch(batch)
    scans = [scenes.trajectory_scan(2 * i, resolution=res) for i in range(n_scans)]
    ow = oracle.OracleWorld(**params)
    ow.set_faithful_inner_update(False)
    rounds = (n_scans + G - 1) // G
    empty = np.zeros((0, 3), np.float32)
    errors = [None] * G

    def rank_main(r):
        """What one rank's process does: its scan of every round, then the collective flush."""
        w = shards[r]
        try:
            for k in range(rounds):
                i = k * G + r
                if i < n_scans:
                    w.dist_insert_pointcloud(scans[i]["T"], scans[i]["points"])
                else:
                    w.dist_insert_pointcloud(np.eye(4), empty)      # ragged last round: no scan on this rank
            w.dist_flush()
        except Exception as e:      # noqa: BLE001
            errors[r] = e

    # one host thread per rank (ctypes drops the GIL inside the library): a rank's host may block on
    # its peers -- sizing the table needs the round's counts -- exactly like separate processes do
    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(G)]
    for t in threads:
        t.start()
    for s in scans:
        ow.insertPointcloud(s["T"], s["points"])
    for t in threads:
        t.join()
    for r, e in enumerate(errors):
        if e is not None:
            print("rank", r, "failed:", repr(e), flush=True)
    if any(e is not None for e in errors):
        return 1
    ks, vs, cnt, dig = [], [], 0, 0
    for r, w in enumerate(shards):
        k, v = w.export_leaves()
        assert np.all(sharded.owner_of_keys(k, G) == r), "a leaf sits on the wrong shard"
        n, d = w.leaf_digest()
        assert (n, d) == sharded.leaf_digest(k, v)
        cnt += n
        dig = (dig + d) % (1 << 64)
        ks.append(k); vs.append(v)
    gp = oracle.pack_keys(np.concatenate(ks))
    gv = np.concatenate(vs).view(np.uint32)
    o = np.argsort(gp)
    k0, v0 = ow.export_leaves()
    p0 = oracle.pack_keys(k0)
    o0 = np.argsort(p0)
    ok = np.array_equal(gp[o], p0[o0]) and np.array_equal(gv[o], v0.view(np.uint32)[o0])
    ok = ok and (cnt, dig) == sharded.leaf_digest(k0, v0)
    trav = sum(w.totals()["traversals"] for w in shards)
    print("PEER_EMULATION_OK" if ok else "PEER_EMULATION_MISMATCH", G, batch, n_scans, len(p0), trav, flush=True)
    for w in shards:
        w.dist_shutdown()
    for w in shards:
        w.close()
    return 0 if ok else 1


if __name__ == "__main__":
    import faulthandler
    faulthandler.dump_traceback_later(170, exit=True)
    sys.exit(main())
