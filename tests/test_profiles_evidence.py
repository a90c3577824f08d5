"""The bench lines committed under profiles/ (the evidence DESIGN.md and README.md quote) carry every
key of the bench contract and are self-consistent: value = traversals per step / time per step,
roofline.frac = achieved / peak, the N > 1 lines passed their parity check.  CPU only."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROFILES = os.path.join(ROOT, "profiles")


def bench_line(name):
    with open(os.path.join(PROFILES, name)) as f:
        lines = [ln for ln in f if ln.startswith("{")]
    assert len(lines) == 1, name
    return json.loads(lines[0])


def check_common(d, n_gpus):
    assert d["metric"] == "ray_voxel_traversals_per_s" and d["unit"] == "traversals/s"
    assert d["n_gpus"] == n_gpus and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and d["warmup"] >= 3
    assert "workload" in d["config"] and "l2" in d["config"] and "model" not in d["config"]
    assert d["gpu_launches"] > 0
    per_step = d["config"]["traversals_per_step"]
    assert d["value"] == pytest.approx(per_step / (d["ms_per_step"] * 1e-3), rel=1e-6)
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 133312 * 12 * n_gpus and e["d2h_bytes_per_step"] > 0
    assert 0.3 * d["value"] < e["value"] < 1.05 * d["value"]       # a copy in the timed region cannot make it faster
    c = d["clocks"]
    assert c["sm_mhz"] and c["sm_mhz"] >= 0.9 * c["sm_max_mhz"]
    assert not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


@pytest.mark.parametrize("name", ["r02_bench_n1_final.json", "r02_bench_n1_driver_style.json"])
def test_single_gpu_lines(name):
    d = bench_line(name)
    check_common(d, 1)
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] == "GB/s" and r["peak"] > 1000
    assert r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-9)
    assert r["achieved"] == pytest.approx(r["algorithmic_bytes_per_launch"] / (r["avg_launch_ms"] * 1e-3) / 1e9, rel=1e-6)
    assert r["traffic"] and 0 < r["issue_frac"] < 1
    cpu = d["cpu_baseline"]
    assert cpu["kind"] == "port" and cpu["cores"] == 1 and cpu["value"] > 1e6 and "sample" in cpu
    assert d["value"] > 1000 * cpu["value"]
    q = d["queries"]
    for leg in ("uniform", "hard"):
        assert q[leg]["probes_per_segment"] > 1 and 0 < q[leg]["roofline"]["frac"] < 1
    assert d["e2e_dropin"]["same_map_as_lean"] is True
    assert d["map_build"]["scans"] == 32 and d["map_build"]["replayed"] == 0


def test_small_configs_leg_agrees_with_the_oracle_counts():
    s = bench_line("r02_bench_n1_driver_style.json")["small_configs"]
    assert s["cfg1"]["traversals"] == 362225 and s["cfg1"]["rays_cast"] == 9638      # SURVEY 8(d), [PROBED]
    for cfg in ("cfg1", "cfg2"):
        assert s[cfg]["same_traversals_as_cpu"] is True and s[cfg]["ms_per_insert"] < s[cfg]["cpu_ms"]


@pytest.mark.parametrize("name,n", [("r02_bench_n2_final.log", 2), ("r02_bench_n4_final.log", 4),
                                    ("r02_bench_n8_final.log", 8)])
def test_multi_gpu_lines_passed_their_parity_check(name, n):
    d = bench_line(name)
    check_common(d, n)
    p = d["parity"]
    assert p["checked"] is True and p["equal"] is True
    assert p["leaves"] == p["single_gpu_leaves"] and p["digest"] == p["single_gpu_digest"]
    q = d["queries_sharded"]
    assert q["replica_equals_shards"] is True and q["replica_leaves"] == p["leaves"]
    single = bench_line("r02_bench_n1_driver_style.json")
    assert d["value"] > 0.8 * n * single["value"]                  # scaling the docs quote: >= 0.8 of linear
