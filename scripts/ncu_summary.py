#!/usr/bin/env python3
"""One-screen summary of an `ncu --set full` report: python scripts/ncu_summary.py report.ncu-rep [more.ncu-rep ...]"""
import csv, io, subprocess, sys
WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__cycles_active.avg", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
        "smsp__inst_executed_op_global_ld.sum", "sm__sass_inst_executed_op_local.sum", "l1tex__t_bytes_pipe_lsu_mem_local_op_ld.sum",
        "l1tex__t_bytes_pipe_lsu_mem_local_op_st.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum"]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("==", rep, "|", d.get("Kernel Name", "")[:90])
        for k in WANT:
            if k in d:
                print("  %-70s %16s %s" % (k, d[k], units[hdr.index(k)]))
        st = []
        for k in hdr:
            if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and "not_issued" not in k:
                try:
                    st.append((float(d[k]), k.split("issue_stalled_")[1].split("_per_issue")[0]))
                except ValueError:
                    pass
        print("  stalls per issue:", ", ".join("%s %.2f" % (n, v) for v, n in sorted(st, reverse=True)[:8]))
