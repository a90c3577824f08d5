"""Pipelined throughput (three scan contexts) against the walker's CTAs per SM and the K4 variant: how the
issue-bound walk of one scan and the HBM-bound update of the scan before it share the SMs.  Torch-free.
   usage: tools/ab_cosched.py [n_scans]"""
import json
import os
import sys

src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ab_helpers.py")).read()
ns = {"__name__": "ab_helpers_setup"}
exec(compile(src[: src.index("# ---- main")], "ab_helpers.py", "exec"), ns)     # scans on the device, make / run_sync / run_pipe
run_pipe, EXTRA = ns["run_pipe"], ns["EXTRA"]
rows = []
ref = None
for rnd in range(2):
    for ctas in (3, 4, 5):
        for upd in (0, 2):
            EXTRA["VMAP_WALK_CTAS"] = str(ctas)
            EXTRA["VMAP_UPDATE_VARIANT"] = str(upd)
            a, dg, _ = run_pipe(1, 123, None, 3, 0, reps=2)
            b, dg2, _ = run_pipe(1, 123, None, 3, 256 << 20, reps=2)
            ref = ref or dg
            row = {"round": rnd, "walk_ctas": ctas, "update_variant": upd, "pipelined_ms": a, "pipelined_flushed_ms": b,
                   "identical": dg == ref and dg2 == ref}
            rows.append(row)
            print(json.dumps(row), flush=True)
json.dump(rows, open("gpurun_out/ab_cosched.json", "w"), indent=1)
print("AB_OK" if all(r["identical"] for r in rows) else "AB_MISMATCH")
