"""Turn the raw outputs of tools/gpu_evidence_{a,b}.sh (gpurun_out/<tag>_*) into the tracked summaries under profiles/
(reads .ncu-rep files here with `ncu -i`, no GPU needed).  Usage: python tools/collect_profiles.py <tag> [round_prefix]"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ncu_summary  # noqa: E402

HBM_PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbps_sustained", 6536.7) \
    if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6536.7


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        return []
    hdr = rows[0]
    return [dict(zip(hdr, r)) for r in rows[2:]]


def fnum(d, k):
    try:
        return float(d.get(k, "0").replace(",", ""))
    except ValueError:
        return 0.0


def dram_bytes(d):
    tot = 0.0
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        v = fnum(d, k)
        # ncu prints a unit row separately; raw page values for *_bytes are in the unit of the second header row, which
        # ncu_summary keeps: recompute from the text when needed
        tot += v
    return tot


def main(tag, pre="r02"):
    go, pr = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
    cp = {f"{tag}_bench.json": f"{pre}_bench_1gpu.json", f"{tag}_bench.err": f"{pre}_bench_1gpu.err",
          f"{tag}_bench_ref.json": f"{pre}_bench_reference_arm.json", f"{tag}_tests.log": f"{pre}_gpu_tests.log",
          f"{tag}_smoke.log": f"{pre}_smoke.log",
          f"{tag}_memcheck.log": f"{pre}_sanitizer_memcheck.log", f"{tag}_racecheck.log": f"{pre}_sanitizer_racecheck.log",
          f"{tag}_synccheck.log": f"{pre}_sanitizer_synccheck.log"}
    for c in ("waymo_l", "lc", "c_r50"):
        cp[f"{tag}_bench_{c}.json"] = f"{pre}_bench_{c}.json"
    for a, b in cp.items():
        if os.path.exists(os.path.join(go, a)):
            shutil.copyfile(os.path.join(go, a), os.path.join(pr, b))
            print("copied", b)
    # ---- launch list
    lc = os.path.join(go, f"{tag}_launches.csv")
    if os.path.exists(lc):
        rows = list(csv.reader(open(lc)))
        hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
        h = rows[hi]
        ki, vi, gi, bi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size"), h.index("Block Size")
        seq = [(r[ki], float(r[vi].replace(",", "")) / 1e3, r[gi], r[bi]) for r in rows[hi + 2:] if len(r) > vi]
        # one forward = from one vox_hash_kernel to the next
        starts = [i for i, s in enumerate(seq) if "vox_hash_kernel" in s[0]]
        if len(starts) >= 2:
            seq = seq[starts[0]:starts[1]]
        tot = sum(s[1] for s in seq)
        agg = collections.OrderedDict()
        for k, us, _, _ in seq:
            name = k.split("(")[0].replace("void ", "").replace("ff3d::", "")[:70]
            a = agg.setdefault(name, [0, 0.0])
            a[0] += 1
            a[1] += us
        with open(os.path.join(pr, f"{pre}_launches.txt"), "w") as f:
            f.write(f"# one FocalFormer3D_L forward, bs = 4 x 300k points: {len(seq)} launches, {tot / 1e3:.3f} ms summed kernel time\n"
                    "# (ncu --metrics gpu__time_duration.sum --clock-control none: serialised, cold-cache per-launch times -- the SHARES\n"
                    "#  are what compares with the CUDA-event step time)\n\n# by kernel\n")
            for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                f.write(f"{us:10.1f} us  {100 * us / tot:5.1f}%  x{n:<4d} {name}\n")
            f.write("\n# in launch order\n")
            for k, us, g, b in seq:
                f.write(f"{us:9.1f} us  grid={g:14s} block={b:14s} {k.split('(')[0].replace('void ', '')[:90]}\n")
        print("wrote", f"{pre}_launches.txt", len(seq), "launches")
    # ---- ncu --set full reports
    for kind in ("gemm", "nongemm"):
        rep = os.path.join(go, f"{tag}_{kind}.ncu-rep")
        if not os.path.exists(rep):
            continue
        res = ncu_summary.summarize(rep)
        seen, lines = collections.Counter(), []
        for text, d in res:
            name = d.get("Kernel Name", "?")
            seen[name] += 1
            if kind == "nongemm" and seen[name] > 2:
                continue
            lines.append(text)
        with open(os.path.join(pr, f"{pre}_ncu_{kind}.txt"), "w") as f:
            f.write(f"# ncu --set full --clock-control none, one capture per line block; source: gpurun_out/{tag}_{kind}.ncu-rep\n")
            f.write("\n".join(lines) + "\n")
        print("wrote", f"{pre}_ncu_{kind}.txt", len(lines), "captures")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "r02")
