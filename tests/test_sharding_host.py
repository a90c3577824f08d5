"""Host-side logic of the multi-GPU path, on CPU: ownership is a partition (C function ==
numpy mirror), scan dealing covers every scan once in order, and the NCCL-id broadcast plumbing
works over gloo with world_size 2."""
import os
import subprocess
import sys

import numpy as np

from volumetric_mapping_b200 import _capi, sharded

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_owner_function_matches_numpy_mirror_and_balances():
    L = _capi.lib()
    rng = np.random.default_rng(0)
    keys = rng.integers(0, 65536, (20000, 3)).astype(np.uint16)
    for world in (1, 2, 3, 4, 8):
        c = np.array([L.vmap_shard_owner(int(a), int(b), int(cc), world) for a, b, cc in keys[:3000]])
        m = sharded.owner_of_keys(keys[:3000], world)
        assert np.array_equal(c, m)
        counts = np.bincount(sharded.owner_of_keys(keys, world), minlength=world)
        assert counts.min() > 0.8 * len(keys) / world
    # all voxels of one 8^3 block share the owner
    base = np.array([[32768, 32768, 32768]], dtype=np.uint16)
    offs = rng.integers(0, 8, (100, 3)).astype(np.uint16)
    assert len(set(sharded.owner_of_keys(base + offs, 8).tolist())) == 1
    # a dense local neighbourhood (what one scan touches) is balanced too
    g = np.stack(np.meshgrid(*[np.arange(32768 - 200, 32768 + 200, 8)] * 3, indexing="ij"), -1).reshape(-1, 3)
    cnt = np.bincount(sharded.owner_of_keys(g, 8), minlength=8)
    assert cnt.min() > 0.85 * len(g) / 8


def test_scan_dealing_covers_all_scans_in_round_order():
    for n, world in ((1000, 8), (7, 3), (5, 8), (16, 4)):
        per_rank = [sharded.scans_of_rank(n, r, world) for r in range(world)]
        rounds = len(per_rank[0])
        assert all(len(p) == rounds for p in per_rank)
        order = [per_rank[r][k] for k in range(rounds) for r in range(world) if per_rank[r][k] >= 0]
        assert order == list(range(n))


def test_query_slices_partition_the_batch():
    for n in (0, 1, 7, 8, 9, 1000, 1_000_003):
        for world in (1, 2, 3, 8):
            cuts = [sharded.query_slice(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(lo <= hi for lo, hi in cuts)
            assert all(cuts[r][1] == cuts[r + 1][0] for r in range(world - 1))
            per = (n + world - 1) // world
            assert all(hi - lo <= per for lo, hi in cuts)


def test_id_broadcast_over_gloo_world_size_2():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29517",
           os.path.join(ROOT, "tests", "gloo_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("GLOO_CHECK_OK") == 2
