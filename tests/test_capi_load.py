"""CPU-side checks of the drop-in boundary: libvmap_b200.so loads, exports every symbol that
include/vmap/vmap.h declares, and fails loudly (no CPU fallback) without a CUDA device."""
import os
import re

import pytest

import volumetric_mapping_b200 as vm
from volumetric_mapping_b200 import _capi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "vmap", "vmap.h")).read()
    return sorted(set(re.findall(r"VMAP_API\s+[\w\s\*]+?\b(vmap_\w+)\s*\(", src)))


def test_library_is_built_in_tree():
    build.build()
    assert os.path.exists(_capi.lib_path())
    assert os.path.dirname(_capi.lib_path()) == os.path.join(ROOT, "volumetric_mapping_b200")


def test_every_declared_symbol_is_exported_and_bound():
    L = _capi.lib()
    names = header_symbols()
    assert len(names) >= 71
    for n in names:
        assert hasattr(L, n), n
    assert sorted(_capi.SYMBOLS) == names


def test_struct_layouts_match_header():
    import ctypes as C
    assert C.sizeof(_capi.Params) == 13 * 8
    assert C.sizeof(_capi.Config) == 8 + 5 * 8
    assert C.sizeof(_capi.ScanStats) == 20 * 8
    p = _capi.default_params()
    assert (p.resolution, p.probability_hit, p.probability_miss, p.threshold_min, p.threshold_max,
            p.threshold_occupancy, p.sensor_max_range, p.max_free_space,
            p.treat_unknown_as_occupied) == (0.15, 0.65, 0.4, 0.12, 0.97, 0.7, 5.0, 0.0, 1)


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(vm.VmapError) as e:
        vm.OctomapWorld()
    assert e.value.status == 4 and "no CPU fallback" in str(e.value)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "volumetric_mapping_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inl", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "liboracle" not in src, f
