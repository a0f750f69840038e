"""Per-kernel, per-CUDA-source-line warp-stall samples of EVERY capture in an .ncu-rep taken with --import-source on
(one block per capture, repeated kernels of the same name keep only the first two):  ncu_stalls_all.py report.ncu-rep [top_n]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
caps, cur, hdr, fname = [], None, None, None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":        # precedes the block's "Function Name"
        fname = r[1].split("/")[-1]
        continue
    if len(r) == 2 and r[0] == "Function Name":
        # one block per (capture, source file): consecutive blocks of one kernel belong to the same capture until a file repeats
        if cur is None or cur["name"] != r[1] or fname in cur["files"]:
            cur = {"name": r[1], "rows": [], "files": set()}
            caps.append(cur)
        cur["files"].add(fname)
        continue
    if len(r) > 2 and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or cur is None or len(r) < 10 or not r[0] or r[2] != "-":
        continue
    k = int(r[6]) if r[6].isdigit() else 0
    st = {h[6:]: int(r[i]) for i, h in enumerate(hdr)
          if h.startswith("stall_") and "Not Issued" not in h and i < len(r) and r[i].isdigit() and int(r[i]) > 0}
    cur["rows"].append((k, fname, int(r[0]), r[1].strip()[:100], sorted(st.items(), key=lambda kv: -kv[1])[:2]))
seen = collections.Counter()
for c in caps:
    tot = sum(x[0] for x in c["rows"])
    if tot == 0:
        continue
    seen[c["name"]] += 1
    if seen[c["name"]] > 2:
        continue
    print(f"=== {c['name'][:110]}   ({tot} samples)")
    for x in sorted(c["rows"], key=lambda x: -x[0])[:n]:
        print(f"  {x[0]:6d} {100 * x[0] / tot:5.1f}%  {x[1]}:{x[2]:<5d} {x[3]}  {x[4]}")
