"""Multi-GPU map build: one process per GPU, one map shard per process, torch.distributed for the
plumbing (rank discovery, the NCCL id broadcast, barriers), NCCL all-to-all of block records
inside libvmap_b200.so for the data path.

Scan order: scans are consumed in ROUNDS of `world` scans; in round k rank r generates scan
k*world + r, and every shard applies the round's scans in rank order, so every voxel sees its
updates in global scan order exactly like the single-threaded reference.
"""
import numpy as np

from .world import OctomapWorld


def scans_of_rank(n_scans, rank, world):
    """Indices of the scans rank `rank` generates, round by round (-1 = no scan in that round)."""
    rounds = (n_scans + world - 1) // world
    return [k * world + rank if k * world + rank < n_scans else -1 for k in range(rounds)]


def owner_of_keys(keys_k3, world):
    """Owner rank of each key (host mirror of vmap_shard_owner, vectorised)."""
    k = np.asarray(keys_k3, dtype=np.uint64).reshape(-1, 3) >> np.uint64(3)
    b = k[:, 0] | (k[:, 1] << np.uint64(13)) | (k[:, 2] << np.uint64(26))
    with np.errstate(over="ignore"):
        b ^= b >> np.uint64(30)
        b *= np.uint64(0xbf58476d1ce4e5b9)
        b ^= b >> np.uint64(27)
        b *= np.uint64(0x94d049bb133111eb)
        b ^= b >> np.uint64(31)
    return ((b >> np.uint64(40)) % np.uint64(world)).astype(np.int64)


def _mix64(x):
    with np.errstate(over="ignore"):
        x = x ^ (x >> np.uint64(30)); x = x * np.uint64(0xbf58476d1ce4e5b9)
        x = x ^ (x >> np.uint64(27)); x = x * np.uint64(0x94d049bb133111eb)
        x = x ^ (x >> np.uint64(31))
    return x


def leaf_digest(keys_k3, logodds):
    """Host mirror of vmap_leaf_digest: (count, sum over leaves of mix(mix(key48) ^ value bits)
    modulo 2^64).  Shard digests add up (modulo 2^64) to the digest of the whole map."""
    k = np.asarray(keys_k3, dtype=np.uint64).reshape(-1, 3)
    key = k[:, 0] | (k[:, 1] << np.uint64(16)) | (k[:, 2] << np.uint64(32))
    bits = np.ascontiguousarray(logodds, dtype=np.float32).view(np.uint32).astype(np.uint64)
    with np.errstate(over="ignore"):
        d = np.add.reduce(_mix64(_mix64(key) ^ bits), dtype=np.uint64) if len(bits) else np.uint64(0)
    return int(len(bits)), int(d)


def broadcast_id(make_id, rank, dist=None, device=None):
    """Rank 0 creates the 128-byte NCCL id, everybody receives it through torch.distributed."""
    import torch
    if dist is None:
        import torch.distributed as dist
    buf = torch.zeros(128, dtype=torch.uint8, device=device)
    if rank == 0:
        buf.copy_(torch.frombuffer(bytearray(make_id()), dtype=torch.uint8))
    dist.broadcast(buf, src=0)
    return bytes(buf.cpu().numpy().tobytes())


class ShardedOctomapWorld:
    """OctomapWorld sharded over torch.distributed ranks (backend nccl)."""

    def __init__(self, rank, world, device, batch=1, **kw):
        import torch.distributed as dist
        self.rank, self.world = rank, world
        self.local = OctomapWorld(device=device, **kw)
        import torch
        ident = broadcast_id(OctomapWorld.dist_unique_id, rank, dist,
                             device=torch.device("cuda", device))
        self.local.dist_init(ident, rank, world)
        if batch != 1:
            self.local.dist_set_batch(batch)

    def insert_round_device(self, T, data_ptr, n):
        """COLLECTIVE.  n == 0: this rank has no scan in the round."""
        self.local.dist_insert_pointcloud_device(T if T is not None else np.eye(4), data_ptr, n)

    def insert_round(self, T, xyz):
        self.local.dist_insert_pointcloud(T if T is not None else np.eye(4),
                                          xyz if xyz is not None else np.zeros((0, 3), np.float32))

    def close(self):
        self.local.dist_shutdown()
        self.local.close()
