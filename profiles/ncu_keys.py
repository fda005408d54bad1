"""Print the metrics we track from `ncu -i X.ncu-rep --page raw --csv` output.
usage: ncu -i prof.ncu-rep --page raw --csv > raw.csv; python profiles/ncu_keys.py raw.csv"""
import csv
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "launch__waves_per_multiprocessor", "launch__grid_size", "launch__block_size",
    "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_lsu.sum",
    "sm__inst_executed_pipe_xu.sum", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
]
STALL = "smsp__average_warps_issue_stalled_"


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("----", r[idx["Kernel Name"]][:90])
        for w in WANT:
            if w in idx:
                print(f"  {w:70s} {r[idx[w]]:>18s} {units[idx[w]]}")
        st = []
        for h in hdr:
            if h.startswith(STALL) and h.endswith("_per_issue_active.ratio"):
                try:
                    st.append((float(r[idx[h]]), h[len(STALL):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        st.sort(reverse=True)
        print("  stalls (warps per issue):", ", ".join(f"{n}={v:.2f}" for v, n in st[:7]))


if __name__ == "__main__":
    main(sys.argv[1])
