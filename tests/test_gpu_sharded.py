"""Sharded (multi-GPU) path on ONE GPU: G handles stand for G ranks, the test harness plays the
exchange by handing each shard the record lists the others generated for it (device pointers,
same GPU).  Everything else -- generate, pack, apply, the sharded queries -- is the code the
NCCL path runs.  Bar: the UNION of the shards equals the oracle's map bit for bit."""
import numpy as np
import pytest

import oracle
import volumetric_mapping_b200 as vm
from volumetric_mapping_b200 import scenes, sharded

pytestmark = pytest.mark.gpu


def leaves_packed(k, v):
    p = oracle.pack_keys(k)
    order = np.argsort(p, kind="stable")
    return p[order], v.view(np.uint32)[order]


def run_rounds(G, scans, params, capture=False):
    shards = [vm.OctomapWorld(initial_blocks=64, **params) for _ in range(G)]
    for r, w in enumerate(shards):
        w.set_shard(r, G)
    ow = oracle.OracleWorld(**params)
    ow.set_faithful_inner_update(False)
    n_rounds = (len(scans) + G - 1) // G
    for k in range(n_rounds):
        present = []
        for r in range(G):
            i = k * G + r
            if i < len(scans):
                shards[r].shard_generate(scans[i]["T"], scans[i]["points"])
                ow.insertPointcloud(scans[i]["T"], scans[i]["points"])
                present.append(r)
        for d in range(G):
            shards[d].shard_apply([shards[s].shard_records(d) for s in present])
    return shards, ow


@pytest.mark.parametrize("G", [2, 3])
def test_sharded_union_equals_oracle(G):
    rng = np.random.default_rng(21)
    params = dict(resolution=0.1, sensor_max_range=6.0)
    scans = []
    for i in range(7):          # not a multiple of G: the last round is ragged
        d = rng.normal(size=(3000, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        pts = (d * rng.uniform(0.3, 7.0, (3000, 1))).astype(np.float32)
        scans.append({"points": pts, "T": scenes.quat_to_matrix(scenes.random_unit_quat(rng),
                                                                 rng.uniform(-1, 1, 3))})
    shards, ow = run_rounds(G, scans, params)
    ks, vs = [], []
    for r, w in enumerate(shards):
        k, v = w.export_leaves()
        assert np.all(sharded.owner_of_keys(k, G) == r)      # every leaf sits on its owner
        ks.append(k); vs.append(v)
    gk, gv = leaves_packed(np.concatenate(ks), np.concatenate(vs))
    ok, ov = leaves_packed(*ow.export_leaves())
    assert np.array_equal(gk, ok)
    assert np.array_equal(gv, ov)
    assert min(len(k) for k in ks) > 0.5 * len(ok) / G        # balanced ownership


def test_sharded_hdl64_scans_and_queries():
    import torch
    G = 4
    params = dict(resolution=0.2, sensor_max_range=50.0)
    scans = [scenes.trajectory_scan(3 * i, resolution=0.2) for i in range(4)]
    shards, ow = run_rounds(G, scans, params)
    ks, vs = zip(*[w.export_leaves() for w in shards])
    gk, gv = leaves_packed(np.concatenate(ks), np.concatenate(vs))
    ok, ov = leaves_packed(*ow.export_leaves())
    assert np.array_equal(gk, ok) and np.array_equal(gv, ov)
    # queries: element-wise minimum of the shards' partial answers
    q = scenes.cfg5_queries(n_segments=100000, n_points=100000)
    pts = torch.from_numpy(q["points"]).cuda()
    seg = torch.from_numpy(q["segments"]).cuda()
    st = torch.full((G, pts.shape[0]), 255, dtype=torch.uint8, device="cuda")
    pk = torch.zeros((G, seg.shape[0]), dtype=torch.int64, device="cuda")
    tmp = torch.zeros(seg.shape[0], dtype=torch.int32, device="cuda")
    for r, w in enumerate(shards):
        w.shard_query_points_device(pts.data_ptr(), pts.shape[0], st[r].data_ptr())
        w.shard_query_lines_device(seg.data_ptr(), seg.shape[0], tmp.data_ptr())
        pk[r] = tmp.to(torch.int64) & 0xFFFFFFFF
    torch.cuda.synchronize()
    got_p = st.min(dim=0).values.cpu().numpy()
    m = pk.min(dim=0).values
    got_l = torch.where(m == 0xFFFFFFFF, torch.zeros_like(m), m & 3).to(torch.uint8).cpu().numpy()
    assert np.array_equal(got_p, ow.getCellStatusPoint(q["points"]))
    assert np.array_equal(got_l, ow.getLineStatus(q["segments"]))


def test_replica_of_the_shards_equals_the_oracle_map_and_answers_queries():
    """The sliced query mode's replica (SURVEY 8(e) "gather then query on every GPU over a segment
    slice"): the shards' leaves move device to device (vmap_export_leaves_dev /
    vmap_import_leaves_dev) into one single-GPU map, which must equal the oracle's map bit for bit
    and answer getCellStatusPoint / getLineStatus like it."""
    import torch
    G = 3
    params = dict(resolution=0.2, sensor_max_range=50.0)
    scans = [scenes.trajectory_scan(3 * i, resolution=0.2) for i in range(4)]
    shards, ow = run_rounds(G, scans, params)
    rep = sharded.build_replica(shards, initial_blocks=64)        # small on purpose: table and pool grow in the import
    n_sum = sum(w.leaf_digest()[0] for w in shards)
    d_sum = sum(w.leaf_digest()[1] for w in shards) % (1 << 64)
    assert rep.leaf_digest() == (n_sum, d_sum)
    gk, gv = leaves_packed(*rep.export_leaves())
    ok, ov = leaves_packed(*ow.export_leaves())
    assert np.array_equal(gk, ok) and np.array_equal(gv, ov)
    q = scenes.cfg5_queries(n_segments=60000, n_points=60000)
    pts = torch.from_numpy(q["points"]).cuda()
    seg = torch.from_numpy(q["segments"]).cuda()
    want_p, want_l = ow.getCellStatusPoint(q["points"]), ow.getLineStatus(q["segments"])
    got_p = torch.zeros(pts.shape[0], dtype=torch.uint8, device="cuda")
    got_l = torch.zeros(seg.shape[0], dtype=torch.uint8, device="cuda")
    for r in range(G):                   # every "rank" answers its slice on the replica
        lo, hi = sharded.query_slice(pts.shape[0], r, G)
        rep.getCellStatusPointDevice(pts[lo:hi].data_ptr(), hi - lo, got_p[lo:].data_ptr())
        lo, hi = sharded.query_slice(seg.shape[0], r, G)
        rep.getLineStatusDevice(seg[lo:hi].data_ptr(), hi - lo, got_l[lo:].data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(got_p.cpu().numpy(), want_p)
    assert np.array_equal(got_l.cpu().numpy(), want_l)
    # capacity too small: an error, nothing written past the buffer
    small = torch.zeros(6 * 8, dtype=torch.uint8, device="cuda")
    smallv = torch.zeros(8, dtype=torch.float32, device="cuda")
    with pytest.raises(vm.VmapError):
        rep.export_leaves_device(small.data_ptr(), smallv.data_ptr(), 8)
    empty = vm.OctomapWorld(**params)
    assert empty.export_leaves_device(0, 0, 0) == 0
    empty.import_leaves_device(0, 0, 0)
    assert empty.leaf_count() == 0


@pytest.mark.parametrize("G,batch,n_scans", [(2, 1, 5), (3, 2, 10), (4, 4, 19)])
def test_peer_memory_transport_emulated_on_one_gpu(G, batch, n_scans):
    """The peer-memory exchange (round buffers, publish / collect / consumed flags, applies that read
    the peers' buffers) with G handles of one process standing for G ranks; own process because it
    needs CUDA_DEVICE_MAX_CONNECTIONS set before the CUDA context exists."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "peer_emulation_check.py"), str(G), str(batch),
                        str(n_scans)], capture_output=True, text=True, timeout=240)
    assert r.returncode == 0 and "PEER_EMULATION_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-6000:]


@pytest.mark.parametrize("batch", [1, 2])
def test_nccl_two_ranks_if_two_gpus(batch):
    """The real exchange: needs >= 2 GPUs (gpurun --gpus 2); skipped on a 1-GPU box.  batch = scans
    per rank that share one exchange round (a ragged last batch is part of the case)."""
    import subprocess
    import sys
    import os
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(29531 + batch),
           os.path.join(root, "tests", "dist_check.py")]
    import signal
    proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                            env=dict(os.environ, DIST_BATCH=str(batch), VMAP_TRACE="1"), start_new_session=True)
    try:
        stdout, stderr = proc.communicate(timeout=150)
    except subprocess.TimeoutExpired:
        os.killpg(proc.pid, signal.SIGKILL)      # the whole process group: no stray ranks on the GPUs
        stdout, stderr = proc.communicate()
        raise AssertionError("dist_check timed out\n" + stdout[-3000:] + stderr[-6000:])
    assert proc.returncode == 0, stdout[-3000:] + stderr[-6000:]

    class out:   # noqa: N801
        pass
    out.stdout = stdout
    assert "DIST_CHECK_OK" in out.stdout
