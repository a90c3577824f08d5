"""C++ host code over the C ABI: builds examples/demo_octomap_world (g++ only, links the in-tree
libvmap_b200.so) and runs it on the GPU; the program checks hand-derived known answers through
the OctomapWorld / WorldBase mirror classes and exits non-zero on any mismatch."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_demo_builds_on_cpu():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "examples")], stdout=subprocess.DEVNULL)
    assert os.path.exists(os.path.join(ROOT, "examples", "demo_octomap_world"))


@pytest.mark.gpu
def test_cpp_demo_known_answers_on_gpu():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "examples")], stdout=subprocess.DEVNULL)
    out = subprocess.run([os.path.join(ROOT, "examples", "demo_octomap_world")], capture_output=True,
                         text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "demo ok" in out.stdout
