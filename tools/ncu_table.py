"""Compact table from a tools/ncu_summary.py text dump: one line per capture with duration, DRAM bytes, achieved GB/s as a
fraction of the measured HBM peak, L2 hit rate, tensor-pipe activity.   ncu_table.py summary.txt [max_per_kernel]"""
import collections
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TUNIT = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def main(path, cap=2):
    blocks, cur = [], None
    for line in open(path):
        if line.startswith("kernel:"):
            cur = {"name": line[len("kernel:"):].strip(), "m": {}}
            blocks.append(cur)
        elif cur is not None and line.startswith("  "):
            p = line.split()
            if len(p) >= 2:
                cur["m"][p[0]] = (p[1] if len(p) >= 3 else "", p[-1])
    seen = collections.Counter()
    print(f"# one line per ncu --set full capture (first {cap} launches of each kernel); HBM peak = {PEAK} GB/s (MEASURED_PEAKS.json)")
    print(f"# {'us':>8s} {'DRAM MB':>9s} {'GB/s':>8s} {'%HBM':>6s} {'L2hit%':>7s} {'lts%':>6s} {'tensor%':>8s} {'warps%':>7s}  kernel  grid")
    for b in blocks:
        m = b["m"]
        short = re.sub(r"\(.*?\)\s*grid", " grid", b["name"])
        key = short.split(" grid")[0]
        seen[key] += 1
        if seen[key] > cap:
            continue

        def val(k, table=None):
            if k not in m:
                return 0.0
            u, v = m[k]
            f = float(v.replace(",", ""))
            return f * table.get(u, 1.0) if table else f
        us = val("gpu__time_duration.sum", TUNIT)
        by = val("dram__bytes_read.sum", UNIT) + val("dram__bytes_write.sum", UNIT)
        gbs = by / (us * 1e-6) / 1e9 if us > 0 else 0.0
        print(f"  {us:8.1f} {by / 1e6:9.1f} {gbs:8.0f} {100 * gbs / PEAK:6.1f} {val('lts__t_sector_hit_rate.pct'):7.1f} "
              f"{val('lts__throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} "
              f"{val('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):8.1f} "
              f"{val('sm__warps_active.avg.pct_of_peak_sustained_active'):7.1f}  {short}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 2)
