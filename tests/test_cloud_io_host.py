"""The .pcd / .ply readers behind loadOctomapCallback's point-cloud branch (octomap_manager.cc:
313-344; vmap_cloud_io_host.inl), on files laid out the way PCL writes them: ascii and binary PCD
with extra fields, ascii / little- / big-endian PLY with colour properties and a face element."""
import ctypes as C
import os
import struct

import numpy as np
import pytest

from volumetric_mapping_b200 import _capi


def read(path):
    L = _capi.lib()
    n = C.c_size_t(0)
    st = L.vmap_host_read_pointcloud_file(os.fsencode(str(path)), None, 0, C.byref(n))
    if st not in (0, 5):
        return st, None
    out = np.zeros((n.value, 3), np.float32)
    if n.value:
        assert L.vmap_host_read_pointcloud_file(os.fsencode(str(path)), out.ctypes.data, n.value, C.byref(n)) == 0
    return 0, out


def cloud(n=257, seed=0):
    rng = np.random.default_rng(seed)
    p = rng.normal(size=(n, 3)).astype(np.float32) * 7.0
    p[3] = [np.nan, 1.0, 2.0]
    p[5] = [np.inf, -np.inf, 0.0]
    p[7] = [1e-30, -0.0, 3.4e38]
    return p


def same(a, b):
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


def pcd_header(fields, sizes, types, n, data):
    return ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS %s\nSIZE %s\nTYPE %s\nCOUNT %s\n"
            "WIDTH %d\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA %s\n" %
            (" ".join(fields), " ".join(map(str, sizes)), " ".join(types), " ".join("1" for _ in fields), n, n, data))


def test_pcd_ascii_and_binary(tmp_path):
    p = cloud()
    # ascii, xyz only; repr() round-trips float32 through strtod
    f = tmp_path / "a.pcd"
    with open(f, "w") as o:
        o.write(pcd_header("xyz", [4, 4, 4], "FFF", len(p), "ascii"))
        for r in p:
            o.write(" ".join("nan" if np.isnan(v) else repr(float(v)) for v in r) + "\n")
    st, got = read(f)
    assert st == 0 and same(got, p)
    # binary with an rgb field in front and an intensity behind: offsets must come from FIELDS / SIZE
    f = tmp_path / "b.pcd"
    rec = np.zeros(len(p), dtype=[("rgb", "<u4"), ("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("i", "<f8")])
    rec["x"], rec["y"], rec["z"] = p[:, 0], p[:, 1], p[:, 2]
    rec["rgb"] = 0xFFAA00
    with open(f, "wb") as o:
        o.write(pcd_header(["rgb", "x", "y", "z", "intensity"], [4, 4, 4, 4, 8], "UFFFF", len(p), "binary").encode())
        o.write(rec.tobytes())
    st, got = read(f)
    assert st == 0 and same(got, p)
    # double coordinates are narrowed to float
    f = tmp_path / "d.pcd"
    with open(f, "wb") as o:
        o.write(pcd_header("xyz", [8, 8, 8], "FFF", len(p), "binary").encode())
        o.write(p.astype("<f8").tobytes())
    st, got = read(f)
    assert st == 0 and same(got, p)


def ply_header(fmt, n, extra_props=(), faces=0):
    h = "ply\nformat %s 1.0\ncomment made by a test\nelement vertex %d\n" % (fmt, n)
    h += "property float x\nproperty float y\nproperty float z\n"
    for t, name in extra_props:
        h += "property %s %s\n" % (t, name)
    if faces:
        h += "element face %d\nproperty list uchar int vertex_indices\n" % faces
    return h + "end_header\n"


def test_ply_ascii_and_binary(tmp_path):
    p = cloud(seed=1)
    f = tmp_path / "a.ply"
    with open(f, "w") as o:
        o.write(ply_header("ascii", len(p), [("uchar", "red"), ("uchar", "green"), ("uchar", "blue")], faces=1))
        for r in p:
            o.write(" ".join("nan" if np.isnan(v) else repr(float(v)) for v in r) + " 255 0 7\n")
        o.write("3 0 1 2\n")
    st, got = read(f)
    assert st == 0 and same(got, p)
    for fmt, e in (("binary_little_endian", "<"), ("binary_big_endian", ">")):
        f = tmp_path / (fmt + ".ply")
        with open(f, "wb") as o:
            o.write(ply_header(fmt, len(p), [("uchar", "red"), ("double", "curvature")], faces=1).encode())
            for r in p:
                o.write(struct.pack(e + "fffBd", float(r[0]), float(r[1]), float(r[2]), 9, 0.25))
            o.write(struct.pack(e + "Biii", 3, 0, 1, 2))
        st, got = read(f)
        assert st == 0 and same(got, p), fmt


def test_errors_are_reported_not_guessed(tmp_path):
    L = _capi.lib()
    n = C.c_size_t(0)
    assert L.vmap_host_read_pointcloud_file(b"/nonexistent/x.pcd", None, 0, C.byref(n)) == 1
    f = tmp_path / "x.xyz"
    f.write_text("1 2 3\n")
    assert read(f)[0] == 1                                 # "No known file extension"
    f = tmp_path / "c.pcd"
    f.write_bytes(pcd_header("xyz", [4, 4, 4], "FFF", 4, "binary_compressed").encode() + b"\0" * 64)
    assert read(f)[0] == 1
    f = tmp_path / "t.pcd"
    f.write_bytes(pcd_header("xyz", [4, 4, 4], "FFF", 100, "binary").encode() + b"\0" * 64)
    assert read(f)[0] == 1                                 # truncated
    f = tmp_path / "n.pcd"
    f.write_text(pcd_header(["x", "y", "intensity"], [4, 4, 4], "FFF", 1, "ascii") + "1 2 3\n")
    assert read(f)[0] == 1                                 # no z
    f = tmp_path / "l.ply"
    f.write_text("ply\nformat ascii 1.0\nelement vertex 1\nproperty float x\nproperty float y\nproperty float z\n"
                 "property list uchar int neighbours\nend_header\n0 0 0 0\n")
    assert read(f)[0] == 1
    f = tmp_path / "e.pcd"
    f.write_text(pcd_header("xyz", [4, 4, 4], "FFF", 0, "ascii"))
    st, got = read(f)
    assert st == 0 and got.shape == (0, 3)
