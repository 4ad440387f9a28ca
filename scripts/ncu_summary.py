"""Per-launch summary of an ncu report: `python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [--md]`.

Reads `ncu -i REP --page raw --csv` and prints, per kernel launch, the metrics the profiles/ notes quote (duration, warp
instructions, IPC, occupancy, DRAM bytes, L1/L2 hit rates, the main stall reasons per issued instruction)."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
md = "--md" in sys.argv
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}


def g(r, name, scale=1.0):
    i = col.get(name)
    if i is None or r[i] == "":
        return float("nan")
    try:
        v = float(r[i].replace(",", ""))
    except ValueError:
        return float("nan")
    u = units[i]
    if name.startswith("gpu__time_duration"):
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}.get(u, 1.0)
    if "bytes" in name:
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    return v * scale


FIELDS = [
    ("us", "gpu__time_duration.sum", 1.0),
    ("winst_M", "smsp__inst_executed.sum", 1e-6),
    ("ipc", "sm__inst_executed.avg.per_cycle_active", 1.0),
    ("warps%", "sm__warps_active.avg.pct_of_peak_sustained_active", 1.0),
    ("regs", "launch__registers_per_thread", 1.0),
    ("dramR_MB", "dram__bytes_read.sum", 1e-6),
    ("dramW_MB", "dram__bytes_write.sum", 1e-6),
    ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    ("l1hit%", "l1tex__t_sector_hit_rate.pct", 1.0),
    ("l2hit%", "lts__t_sector_hit_rate.pct", 1.0),
    ("st_long", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", 1.0),
    ("st_short", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", 1.0),
    ("st_barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", 1.0),
    ("st_wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", 1.0),
    ("st_mio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", 1.0),
    ("st_lg", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", 1.0),
    ("st_math", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", 1.0),
    ("st_notsel", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", 1.0),
    ("st_branch", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", 1.0),
    ("st_nosinst", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", 1.0),
    ("st_membar", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", 1.0),
    ("st_sleep", "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", 1.0),
    ("smem_conf", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", 1.0),
    ("elig", "smsp__warps_eligible.avg.per_cycle_active", 1.0),
]
names = [f[0] for f in FIELDS]
out = []
for r in data:
    kn = r[col["Kernel Name"]].split("(")[0]
    out.append([kn] + [g(r, m, sc) for _, m, sc in FIELDS])
if md:
    print("| kernel | " + " | ".join(names) + " |")
    print("|---|" + "---|" * len(names))
    for o in out:
        print("| `" + o[0] + "` | " + " | ".join(f"{v:.2f}" for v in o[1:]) + " |")
else:
    for o in out:
        print(o[0])
        print("   " + "  ".join(f"{n}={v:.2f}" for n, v in zip(names, o[1:])))
