"""SASS evidence per kernel of libff3d.so (runs here: `cuobjdump -sass`, no GPU): for every kernel that contains one of the
Blackwell data-path instructions, the count of each and one sample line.  Output -> profiles/r02_sass_excerpt.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["UTCHMMA", "UTMALDG", "UBLKCP", "LDGSTS", "ARRIVES.LDGSTSBAR", "HMMA.16816", "LDSM", "LDTM", "UTCBAR", "UTCATOMSWS",
        "SYNCS.ARRIVE", "SYNCS.PHASECHK", "REDG", "ATOMG", "REDUX", "MUFU.EX2"]


def main(out_path):
    so = os.path.join(ROOT, "focalformer3d_b200", "libff3d.so")
    txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    fn, per = None, collections.OrderedDict()
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            per[fn] = collections.OrderedDict()
            continue
        if fn is None:
            continue
        for w in WANT:
            if w in line:
                c = per[fn].setdefault(w, [0, None])
                c[0] += 1
                if c[1] is None:
                    txt1 = re.sub(r"\s+", " ", line.split("*/")[1] if "*/" in line else line).strip()
                    c[1] = re.sub(r"\s*;\s*/\*.*$", "", txt1)[:110]
    lines = ["# SASS excerpt of focalformer3d_b200/libff3d.so (cuobjdump -sass, sm_100a): instruction counts per kernel + one sample",
             "# UTCHMMA = tcgen05.mma, UTMALDG = cp.async.bulk.tensor (TMA tile / gather4 loads), UBLKCP = cp.async.bulk,",
             "# LDGSTS = cp.async, HMMA.16816 = mma.sync m16n8k16, LDSM = ldmatrix, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit", ""]
    tot = collections.Counter()
    for fn, d in per.items():
        if not d:
            continue
        name = re.sub(r"\(.*", "", fn)
        lines.append(name)
        for w, (n, sample) in d.items():
            tot[w] += n
            lines.append(f"    {w:18s} x{n:<5d} {sample}")
    lines += ["", "# totals: " + ", ".join(f"{w} {n}" for w, n in tot.items())]
    open(out_path, "w").write("\n".join(lines) + "\n")
    print(out_path, "kernels:", sum(1 for d in per.values() if d), dict(tot))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_excerpt.txt"))
