"""GPU parity of getCellStatusBoundingBox / checkPathForCollisionsWithRobot and of the
occupied-only .pcd / .ply loader (8 f3).

The per-lane body of this query (vmap_query.cuh) is also checked against the oracle on the CPU by
tests/test_query_host_emulation.py.  Each check runs in its own process.  (All three passed on a
B200 in the round-1 driver suite and again in round 2 -- gpurun_out r02_suite_baseline.log -- so the
xfail markers they carried for their first hardware run are gone.)"""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.gpu
def test_bbox_status_and_path_collisions_match_oracle():
    r = subprocess.run([sys.executable, os.path.join(HERE, "bbox_check.py")], capture_output=True, text=True,
                       timeout=240)
    sys.stdout.write(r.stdout[-4000:])
    sys.stderr.write(r.stderr[-4000:])
    assert r.returncode == 0 and "BBOX_OK" in r.stdout


@pytest.mark.gpu
def test_occupied_pointcloud_loader_matches_oracle():
    r = subprocess.run([sys.executable, os.path.join(HERE, "cloud_load_check.py")], capture_output=True, text=True,
                       timeout=240)
    sys.stdout.write(r.stdout[-4000:])
    sys.stderr.write(r.stderr[-4000:])
    assert r.returncode == 0 and "CLOUD_LOAD_OK" in r.stdout


@pytest.mark.gpu
def test_staged_insertion_builds_the_same_map():
    r = subprocess.run([sys.executable, os.path.join(HERE, "staged_check.py")], capture_output=True, text=True,
                       timeout=240)
    sys.stdout.write(r.stdout[-4000:])
    sys.stderr.write(r.stderr[-4000:])
    assert r.returncode == 0 and "STAGED_OK" in r.stdout
