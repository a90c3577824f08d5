"""torchrun worker (gloo, CPU): the plumbing around the sharded path that does not need a GPU."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from volumetric_mapping_b200 import sharded  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    ident = sharded.broadcast_id(lambda: bytes(range(128)), rank, dist)
    assert ident == bytes(range(128))
    # each rank owns a disjoint part of a common key set; together they cover it
    rng = np.random.default_rng(5)
    keys = rng.integers(0, 65536, (5000, 3)).astype(np.uint16)
    mine = sharded.owner_of_keys(keys, world) == rank
    t = torch.from_numpy(mine.astype(np.int64))
    dist.all_reduce(t)
    assert bool((t == 1).all())
    # scan dealing: ranks agree on the number of rounds (collective calls stay matched)
    rounds = torch.tensor([len(sharded.scans_of_rank(11, rank, world))])
    lo, hi = rounds.clone(), rounds.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert int(lo) == int(hi)
    print("GLOO_CHECK_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
