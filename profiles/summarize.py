#!/usr/bin/env python
"""Condenses an .ncu-rep (read here with `ncu -i`, no GPU needed) into the few numbers DESIGN.md and bench.py cite:
duration, DRAM bytes (traffic), issue-slot utilisation, pipe utilisation, occupancy, top stall reasons, registers.
usage: python profiles/summarize.py gpurun_out/x.ncu-rep profiles/x_summary.json [kernel-substring]"""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_active", "sm__inst_executed.sum",
    "smsp__inst_executed.sum", "sm__instruction_throughput.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
    "sm__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    pat = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    result = {"report": rep.split("/")[-1], "kernels": []}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        if pat and pat not in name:
            continue
        k = {"kernel": name[:120], "grid": r[idx["Grid Size"]], "block": r[idx["Block Size"]], "metrics": {}}
        for key in hdr:
            if key in KEEP or key.startswith("smsp__average_warps_issue_stalled") and key.endswith("_per_issue_active.ratio") \
                    or key.startswith("smsp__average_warp_latency_issue_stalled") or key.startswith("smsp__pcsamp_warps_issue_stalled"):
                v = r[idx[key]]
                if v in ("", "0", "n/a"):
                    continue
                k["metrics"][key] = f"{v} {units[idx[key]]}".strip()
        m = k["metrics"]

        def num(key):
            try:
                return float(m[key].split()[0].replace(",", ""))
            except Exception:
                return None

        rd, wr = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
        if rd is not None and wr is not None:
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            ru = m["dram__bytes_read.sum"].split()[1]
            wu = m["dram__bytes_write.sum"].split()[1]
            k["dram_bytes_per_launch"] = rd * scale[ru] + wr * scale[wu]
        result["kernels"].append(k)
    json.dump(result, open(out, "w"), indent=1)
    print(f"{len(result['kernels'])} kernel(s) -> {out}")


if __name__ == "__main__":
    main()
