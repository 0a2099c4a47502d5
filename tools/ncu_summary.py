"""Summarise an .ncu-rep (one `ncu --set full` capture) into the handful of metrics DESIGN.md / bench.py quote.

  python tools/ncu_summary.py gpurun_out/x.ncu-rep [more-regex ...] > profiles/rNN_x.txt
"""
import csv
import re
import subprocess
import sys

KEYS = [r"^gpu__time_duration\.sum$", r"^dram__bytes_(read|write)\.sum$", r"^dram__throughput\.avg\.pct", r"^gpu__dram_throughput",
        r"^lts__t_bytes\.sum$", r"^lts__t_sector_hit_rate\.pct$", r"^launch__(registers_per_thread|grid_size|block_size|occupancy_limit|waves)",
        r"^launch__shared_mem_per_block_dynamic", r"^sm__throughput\.avg\.pct", r"^sm__warps_active\.avg\.pct",
        r"^sm__inst_executed_pipe_(xu|fma|alu|fmaheavy|fmalite|uniform|lsu|tc|tmem)[a-z_]*\.(avg|sum)\.pct",
        r"^sm__pipe_tensor.*cycles_active.*pct", r"^sm__pipe_(fma|alu|xu)[a-z_]*cycles_active.*pct", r"^smsp__issue_active\.avg\.pct",
        r"^smsp__inst_executed\.sum$", r"^sm__cycles_elapsed\.max$", r"^smsp__average_warps_issue_stalled_.*_per_issue_active",
        r"^smsp__average_warp_latency_issue_stalled", r"^sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off\.(sum|avg)(\.pct_of_peak_sustained_elapsed|\.per_cycle_elapsed)?$"]


def main():
    rep = sys.argv[1]
    extra = sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    pats = [re.compile(k) for k in KEYS + extra]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"== {d.get('Kernel Name', '?')[:100]}  grid {d.get('Grid Size')} block {d.get('Block Size')}")
        for name, unit, val in zip(hdr, units, r):
            if any(p.search(name) for p in pats) and val not in ("", "0"):
                print(f"{name:95s} {val:>18s} {unit}")


if __name__ == "__main__":
    main()
