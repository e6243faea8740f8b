"""Summarise `ncu --set full` reports (read with `ncu -i X.ncu-rep --page raw --csv`) into a markdown
table for profiles/.  Usage: python tools/ncu_summary.py out.md rep1.ncu-rep [rep2.ncu-rep ...]
With --json out.json: also write the DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum of every
captured launch and their mean) that bench.py reports as roofline.traffic."""
import csv
import io
import subprocess
import sys

COLS = [
    ("gpu__time_duration.sum", "time us", lambda v: f"{float(v) / (1 if float(v) < 1e4 else 1e3):.1f}"),
    ("launch__grid_size", "CTAs", lambda v: str(int(float(v)))),
    ("launch__registers_per_thread", "regs", lambda v: str(int(float(v)))),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %", lambda v: f"{float(v):.1f}"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe %", lambda v: f"{float(v):.1f}"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %", lambda v: f"{float(v):.1f}"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes/inst", lambda v: f"{float(v):.1f}"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait", lambda v: f"{float(v):.2f}"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb", lambda v: f"{float(v):.2f}"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math", lambda v: f"{float(v):.2f}"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_inst", lambda v: f"{float(v):.2f}"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch", lambda v: f"{float(v):.2f}"),
    ("sass__inst_executed_local_loads", "local ld inst", lambda v: str(int(float(v)))),
    ("sass__inst_executed_local_stores", "local st inst", lambda v: str(int(float(v)))),
    ("smsp__inst_executed.sum", "warp inst", lambda v: f"{float(v) / 1e6:.2f}M"),
    ("dram__bytes_read.sum", "dram rd", lambda v: v),
    ("dram__bytes_write.sum", "dram wr", lambda v: v),
]


def _bytes(value, unit):
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    return float(value.replace(",", "")) * scale


def main(out, reps, json_out=None, pairs_per_launch=209715):
    traffic = []
    lines = ["| kernel | " + " | ".join(c[1] for c in COLS) + " |", "|---|" + "---|" * len(COLS)]
    for rep in reps:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        ci = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            name = r[ci["Kernel Name"]].replace("void ", "").replace("(GroupArgs<T1, T2>)", "")
            if "dram__bytes_read.sum" in ci and r[ci["dram__bytes_read.sum"]] != "":
                traffic.append({"kernel": name, "dram_read_bytes": _bytes(r[ci["dram__bytes_read.sum"]], units[ci["dram__bytes_read.sum"]]),
                                "dram_write_bytes": _bytes(r[ci["dram__bytes_write.sum"]], units[ci["dram__bytes_write.sum"]])})
            cells = []
            for key, _, fmt in COLS:
                if key not in ci or r[ci[key]] == "":
                    cells.append("n/a")
                    continue
                v = r[ci[key]].replace(",", "")
                u = units[ci[key]]
                try:
                    cells.append(fmt(v) + (f" {u}" if key.startswith("dram") else ""))
                except ValueError:
                    cells.append(v)
            lines.append(f"| `{name}` | " + " | ".join(cells) + " |")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))
    if json_out and traffic:
        import json
        mean = sum(t["dram_read_bytes"] + t["dram_write_bytes"] for t in traffic) / len(traffic)
        json.dump({"mean_dram_bytes_per_launch": mean, "pairs_per_launch": pairs_per_launch, "launches": len(traffic),
                   "per_launch": traffic, "source": ", ".join(reps) + " (ncu --set full --clock-control none)"},
                  open(json_out, "w"), indent=1)


if __name__ == "__main__":
    args = sys.argv[1:]
    jo = None
    if "--json" in args:
        i = args.index("--json")
        jo = args[i + 1]
        del args[i:i + 2]
    main(args[0], args[1:], jo)
