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


def query_slice(n, rank, world):
    """[lo, hi) of n batched queries that rank `rank` answers in the sliced query mode (equal
    contiguous slices, the last ones may be short or empty)."""
    per = (n + world - 1) // world
    return min(rank * per, n), min((rank + 1) * per, n)


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


def build_replica(shards, device=-1, **kw):
    """Non-collective form of ShardedOctomapWorld.replicate for shards that live in ONE process on
    one GPU (the single-GPU emulation of the sharded path): a new single-GPU OctomapWorld holding the
    union of the shards' leaves, moved device to device."""
    import torch
    counts = [w.leaf_count() for w in shards]
    total = sum(counts)
    keys = torch.empty(6 * max(total, 1), dtype=torch.uint8, device="cuda")       # 3 x uint16 per leaf
    vals = torch.empty(max(total, 1), dtype=torch.float32, device="cuda")
    lo = 0
    for w, c in zip(shards, counts):
        if c:
            assert w.export_leaves_device(keys[6 * lo:].data_ptr(), vals[lo:].data_ptr(), c) == c
        lo += c
    torch.cuda.synchronize()
    rep = OctomapWorld(params=shards[0].getOctomapParameters(), device=device, **kw)
    if total:
        rep.import_leaves_device(keys.data_ptr(), vals.data_ptr(), total)
    return rep


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
        self._device = device
        self.replica = None
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

    # -- queries that scale: a read-only replica per GPU, segments sliced over the ranks ------------
    # ("gather then query on every GPU over a segment slice".  The collective queries
    # vmap_dist_query_* need no replica but make every rank walk every segment.)
    def replicate(self):
        """COLLECTIVE.  Every rank ends up with a read-only single-GPU copy of the WHOLE map
        (self.replica): each shard exports its leaves into its slice of one device buffer
        (vmap_export_leaves_dev), the slices travel by NCCL broadcast over NVLink, and the buffer is
        imported in place (vmap_import_leaves_dev).  Call again after further insertions."""
        import torch
        import torch.distributed as dist
        dev = torch.device("cuda", self._device)
        self.local.dist_flush()
        mine = self.local.leaf_count()
        counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(self.world)]
        dist.all_gather(counts, torch.tensor([mine], dtype=torch.int64, device=dev))
        counts = [int(c.item()) for c in counts]
        offs = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        total = int(offs[-1])
        keys = torch.empty(6 * max(total, 1), dtype=torch.uint8, device=dev)      # 3 x uint16 per leaf
        vals = torch.empty(max(total, 1), dtype=torch.float32, device=dev)
        lo = int(offs[self.rank])
        if mine:
            got = self.local.export_leaves_device(keys[6 * lo:].data_ptr(), vals[lo:].data_ptr(), mine)
            assert got == mine
        torch.cuda.synchronize(dev)
        for r in range(self.world):
            a, b = int(offs[r]), int(offs[r + 1])
            if b > a:
                dist.broadcast(keys[6 * a:6 * b], src=r)
                dist.broadcast(vals[a:b], src=r)
        torch.cuda.synchronize(dev)
        if self.replica is not None:
            self.replica.close()
        blocks = 1
        while blocks * 64 < total:        # a block rarely holds fewer than 64 known voxels; grows on demand anyway
            blocks *= 2
        self.replica = OctomapWorld(params=self.local.getOctomapParameters(), device=self._device,
                                    initial_blocks=max(blocks, 1 << 10))
        if total:
            self.replica.import_leaves_device(keys.data_ptr(), vals.data_ptr(), total)
        return total

    def slice_of(self, n):
        """[lo, hi) of n queries that this rank answers."""
        return query_slice(n, self.rank, self.world)

    def _sliced(self, call, q, n):
        import torch
        import torch.distributed as dist
        if self.replica is None:
            raise RuntimeError("replicate() first")
        per = (n + self.world - 1) // self.world
        out = torch.zeros(per * self.world, dtype=torch.uint8, device=q.device)
        lo, hi = self.slice_of(n)
        if hi > lo:
            call(q[lo:hi].data_ptr(), hi - lo, out[lo:].data_ptr())
        torch.cuda.synchronize(q.device)
        dist.all_gather_into_tensor(out, out[self.rank * per:(self.rank + 1) * per].clone())
        return out[:n]

    def getLineStatusSliced(self, segments):
        """COLLECTIVE.  getLineStatus (octomap_world.cc:395-418) for an (n, 6) float64 CUDA tensor that
        every rank holds: rank r walks segments [r*per, (r+1)*per) on its replica; all ranks get all n
        statuses (uint8 CUDA tensor)."""
        return self._sliced(self.replica.getLineStatusDevice, segments, segments.shape[0])

    def getCellStatusPointSliced(self, points):
        """COLLECTIVE.  getCellStatusPoint (octomap_world.cc:346-360), sliced the same way."""
        return self._sliced(self.replica.getCellStatusPointDevice, points, points.shape[0])

    def close(self):
        if self.replica is not None:
            self.replica.close()
            self.replica = None
        self.local.dist_shutdown()
        self.local.close()
