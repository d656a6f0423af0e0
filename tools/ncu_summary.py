"""Key metrics of one-kernel `ncu --set full` reports as a markdown table.
    python tools/ncu_summary.py a.ncu-rep b.ncu-rep ..."""
import csv, subprocess, sys, io
KEYS = [("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs/thread"), ("launch__cluster_size", "cluster size"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
        ("lts__t_bytes.sum", "L2 bytes"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
        ("sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active", "tensor pipe (inst) % active"),
        ("sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe cycles active %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("smsp__inst_executed.sum", "warp instructions"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard / issue"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts")]
cols = []
for path in sys.argv[1:]:
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(txt)))
    hdr, units, row = rd[0], rd[1], rd[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, row)}
    cols.append((path.split("/")[-1], d))
print("| metric | " + " | ".join(c[0] for c in cols) + " |")
print("|---|" + "---|" * len(cols))
print("| kernel | " + " | ".join("`" + c[1]["Kernel Name"][0][:60] + "`" for c in cols) + " |")
for k, label in KEYS:
    vals = []
    for _, d in cols:
        v, u = d.get(k, ("n/a", ""))
        vals.append(f"{v} {u}".strip())
    print(f"| {label} (`{k}`) | " + " | ".join(vals) + " |")
