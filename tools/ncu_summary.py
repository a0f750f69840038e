"""Headline metrics of .ncu-rep files (read here with `ncu -i`, no GPU needed) -> text summaries under profiles/."""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.per_cycle_active",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__cycles_elapsed.max",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "smsp__cycles_active.avg", "l1tex__lsu_writeback_active_mem_lg.sum", "smsp__inst_executed_op_shared_st.sum"]


def summarize(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        lines = [f"=== {path}", f"kernel: {d.get('Kernel Name')}  grid={d.get('Grid Size')} block={d.get('Block Size')}"]
        for i, h in enumerate(hdr):
            if h in WANT:
                lines.append(f"  {h:100s} {units[i]:16s} {vals[i]}")
        res.append(("\n".join(lines), d))
    return res


if __name__ == "__main__":
    for p in sys.argv[1:]:
        for text, _ in summarize(p):
            print(text)
