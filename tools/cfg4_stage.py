import sys, os, json
sys.path.insert(0, ".")
src = open("tools/ab_helpers.py").read()
sys.argv = ["ab_helpers.py", "10", "--res", "0.05"]
ns = {"__name__": "setup"}
exec(compile(src[: src.index("# ---- main")], "ab_helpers.py", "exec"), ns)
s, counts, dg = ns["run_sync"](1, 123, None, reps=2)
print(json.dumps(s))
s, counts, dg = ns["run_sync"](1, 123, None, reps=2, flush=256 << 20)
print(json.dumps(s))
for ctx in (2, 3):
    ms, dg, rep = ns["run_pipe"](1, 123, None, ctx, 0, warm=3, reps=2)
    print(json.dumps({"ctx": ctx, "pipelined_ms": ms, "replayed": rep}))
