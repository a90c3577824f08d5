#!/usr/bin/env python
"""Print the SASS of one kernel of the built library, normalised for diffing (no addresses, no
encodings), plus an instruction histogram.   usage: tools/sass_fn.py <substring of the mangled name> [--hist]"""
import re
import subprocess
import sys
from collections import Counter


def functions(lib="volumetric_mapping_b200/libvmap_b200.so"):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    cur, fns = None, {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            fns[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?)\s*;?\s*/\*", line)
        if cur and m:
            fns[cur].append((int(m.group(1), 16), m.group(2).rstrip(" ;")))
    return fns


if __name__ == "__main__":
    want = sys.argv[1]
    for name, ins in functions().items():
        if want not in name:
            continue
        print(f"== {name}: {len(ins)} instructions")
        if "--hist" in sys.argv:
            c = Counter(i.split()[1] if i.startswith("@") else i.split()[0] for _, i in ins)
            print(", ".join(f"{k}:{v}" for k, v in c.most_common(25)))
        else:
            for a, i in ins:
                print(f"{a:05x}  {i}")
