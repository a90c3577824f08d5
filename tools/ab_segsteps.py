"""A/B of the ray-segment length (VMAP_SEG_STEPS, a compile-time constant of vmap_walk.cuh: DDA steps per
work slot of the segment walker).  Torch-free.

  python tools/ab_segsteps.py build 32 48 96 128     here (nvcc): volumetric_mapping_b200/_ab/libvmap_segN.so
  python tools/ab_segsteps.py run [n_scans]          on a B200: one subprocess per library (the default build
                                                     first), synchronous stage times + pipelined ms per scan with
                                                     and without the in-stream L2 flush; every variant must build
                                                     the default's map (leaf digest + per-scan statistics)
"""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
AB = os.path.join(ROOT, "volumetric_mapping_b200", "_ab")


def build(values):
    sys.path.insert(0, ROOT)
    from volumetric_mapping_b200 import build as b
    os.makedirs(AB, exist_ok=True)
    procs = []
    for v in values:
        out = os.path.join(AB, "libvmap_seg%d.so" % v)
        cmd = [b.nvcc_path()] + b.NVCC_FLAGS + ["-DVMAP_SEG_STEPS=%d" % v, "-o", out] + \
              [os.path.join(b.CSRC, s) for s in b.SOURCES] + b.LINK_FLAGS
        procs.append((out, subprocess.Popen(cmd, cwd=b.HERE)))
    for out, p in procs:
        assert p.wait() == 0, out
        print(out)


def worker(n_scans):
    src = open(os.path.join(ROOT, "tools", "ab_helpers.py")).read()
    sys.argv = [sys.argv[0], str(n_scans)]
    ns = {"__name__": "ab_helpers_setup"}
    os.chdir(ROOT)
    exec(compile(src[: src.index("# ---- main")], "ab_helpers.py", "exec"), ns)
    summary, counts, dg = ns["run_sync"](1, 123, None, reps=2, flush=256 << 20)
    a, dg2, _ = ns["run_pipe"](1, 123, None, 3, 0, reps=3)
    b, dg3, _ = ns["run_pipe"](1, 123, None, 3, 256 << 20, reps=3)
    print("ROW " + json.dumps({"sync": summary, "pipelined_ms": a, "pipelined_flushed_ms": b,
                                "digest": [list(dg), list(dg2), list(dg3)], "counts": counts}), flush=True)


def run(n_scans):
    libs = [("default (64)", None)] + [(os.path.basename(p), p) for p in sorted(glob.glob(os.path.join(AB, "libvmap_seg*.so")))]
    rows, ref = [], None
    for rnd in range(2):
        for name, path in libs:
            env = dict(os.environ)
            if path:
                env["VMAP_LIB"] = path
            else:
                env.pop("VMAP_LIB", None)
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "worker", str(n_scans)], env=env,
                               capture_output=True, text=True, timeout=240)
            line = [ln for ln in r.stdout.splitlines() if ln.startswith("ROW ")]
            if r.returncode != 0 or not line:
                print(json.dumps({"lib": name, "error": (r.stdout + r.stderr)[-600:]}), flush=True)
                continue
            d = json.loads(line[0][4:])
            key = (d["digest"], d["counts"])
            ref = ref or key
            row = {"round": rnd, "lib": name, **d["sync"], "pipelined_ms": d["pipelined_ms"],
                   "pipelined_flushed_ms": d["pipelined_flushed_ms"], "identical": key == ref}
            rows.append(row)
            print(json.dumps(row), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "ab_segsteps.json"), "w"), indent=1)
    print("AB_OK" if rows and all(r["identical"] for r in rows) else "AB_MISMATCH")


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "run"
    if mode == "build":
        build([int(v) for v in sys.argv[2:]] or [32, 48, 96, 128])
    elif mode == "worker":
        worker(int(sys.argv[2]))
    else:
        run(int(sys.argv[2]) if len(sys.argv) > 2 else 24)
