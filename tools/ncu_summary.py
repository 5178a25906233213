#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into the handful of metrics DESIGN.md / profiles/ quote.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/rN_ncu_full_summary.txt"""
import csv
import io
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__shared_mem_per_block_static", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed.sum", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "sm__maximum_warps_per_active_cycle_pct", "smsp__warps_eligible.avg.per_cycle_active"]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rep = sys.argv[1]
    json_out = sys.argv[2] if len(sys.argv) > 2 else None
    traffic = {}
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    seen = set()
    for r in rows[2:]:
        name = r[ki]
        if json_out:
            def val(metric):
                v = float(r[hdr.index(metric)].replace(",", "")) if metric in hdr else 0.0
                u = units[hdr.index(metric)] if metric in hdr else ""
                return v * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3}.get(u, 1.0)
            short = name.split("(")[0].split("::")[-1].split("<")[0].strip()
            t = traffic.setdefault(short, {"launches": 0, "dram_bytes": 0.0, "time_us": 0.0, "alu_pipe_pct": 0.0, "issue_active_pct": 0.0,
                                           "warps_active_pct": 0.0})
            t["launches"] += 1
            # pipe / issue utilisation: mean over the captured launches of the kernel (the ALU roofline of VERDICT r1 item 5)
            for key, metric in (("alu_pipe_pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                                ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                                ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active")):
                t[key] += (val(metric) - t[key]) / t["launches"]
            t["dram_bytes"] += val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
            t["time_us"] += val("gpu__time_duration.sum")
        if name in seen:
            continue
        seen.add(name)
        print("==", name[:90])
        for i, h in enumerate(hdr):
            if h in KEEP or (h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and float(r[i] or 0) >= 0.2):
                print("  %-90s %s %s" % (h, r[i], units[i]))


    if json_out:
        import json
        json.dump(traffic, open(json_out, "w"), indent=1)


if __name__ == "__main__":
    main()
