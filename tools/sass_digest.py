"""SASS digest of the shipped library: per kernel, how many of the instructions that prove the design are present
(UBLKCP = TMA bulk copies, SYNCS = mbarrier operations, IMAD.WIDE.U32 = the 32x32->64 multiply of the lazy Goldilocks MAC,
256-bit stores, ...), registers and shared memory.  No GPU needed.   python tools/sass_digest.py profiles/r02_sass_digest.json"""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "latticeum_b200", "lib", "liblattice_ajtai.so")
WATCH = ["UBLKCP", "SYNCS", "IMAD.WIDE.U32", "IADD3.X", "IMAD.X", "SHF", "SHFL", "LDS", "STS", "LDG", "STG", "ATOMG", "RED", "HMMA", "UTC"]

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()  # noqa: E731
usage = {}
cur = None
for line in res.splitlines():
    m = re.match(r"\s*Function (\S+):", line)
    if m:
        cur = m.group(1)
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", line)
    if m and cur:
        usage[cur] = {"registers": int(m.group(1)), "stack": int(m.group(2)), "static_smem": int(m.group(3))}
out = {"library": os.path.relpath(LIB, ROOT), "arch": sorted(set(re.findall(r"arch = (sm_\w+)", sass))), "kernels": {}}
cur = None
counts = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts = out["kernels"].setdefault(cur, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and counts is not None:
        op = m.group(1)
        counts["instructions"] += 1
        for w in WATCH:
            if op == w or op.startswith(w + ".") or (w in ("UBLKCP", "SYNCS", "UTC") and op.startswith(w)):
                counts[w] += 1
        if op.startswith("STG") and ".256" in op:
            counts["STG.256"] += 1
summary = {}
for k, c in out["kernels"].items():
    name = re.sub(r"\(.*", "", demangle(k)).replace("void ", "")
    d = dict(c)
    d.update(usage.get(k, {}))
    summary[name] = d
out["kernels"] = dict(sorted(summary.items()))
out["totals"] = {w: sum(v.get(w, 0) for v in summary.values()) for w in WATCH + ["STG.256", "instructions"]}
out["tensor_core_instructions"] = out["totals"]["HMMA"] + out["totals"]["UTC"]
dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "sass_digest.json")
json.dump(out, open(dst, "w"), indent=1)
print(dst, out["arch"], out["totals"])
