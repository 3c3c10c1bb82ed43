#!/usr/bin/env python
"""Per-launch pipe / issue utilisation table of an .ncu-rep: ncu_pipes.py rep"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "ms"), ("smsp__inst_executed.sum", "Minst"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
        ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "fmaH%"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lsuWave%"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"),
        ("sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active", "adu%"),
        ("sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "uni%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("launch__registers_per_thread", "regs"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "st_notsel"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
        ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "st_disp"),
        ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "st_br"),
        ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "st_noinst")]
idx = [(h.index(a) if a in h else None, b) for a, b in want]
print(" ".join("%9s" % b[:9] for _, b in idx))
for r in rows[2:]:
    out = []
    for i, b in idx:
        v = r[i] if i is not None else "-"
        if b == "kernel":
            v = v.split("::")[-1][:9]
        else:
            try:
                x = float(v.replace(",", "")); v = "%.2f" % (x / 1e6 if b == "Minst" else x)
            except ValueError:
                pass
        out.append("%9s" % v[:9])
    print(" ".join(out))
