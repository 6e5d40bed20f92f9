#!/usr/bin/env python
"""Turn an .ncu-rep (brought back from the GPU box under gpurun_out/) into the small text
summaries that are committed under profiles/.

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r1_c2_fixpoint

writes <out>_details.csv (duration, DRAM/L2 throughput, occupancy, registers, smem),
<out>_raw.csv (dram bytes, instruction counts, stall sample counts per launch) and
<out>_hot_lines.txt (warp-stall samples aggregated per CUDA source line, needs -lineinfo).
"""
import csv
import io
import subprocess
import sys

DETAILS = ("Duration", "DRAM Throughput", "Memory Throughput", "Registers Per Thread", "Executed Ipc Active",
           "Achieved Occupancy", "L2 Cache Throughput", "SM Busy", "Dynamic Shared Memory Per Block", "Grid Size",
           "Block Size", "Elapsed Cycles", "SM Frequency", "Issue Slots Busy", "Theoretical Occupancy")
RAW = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
       "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "launch__registers_per_thread",
       "smsp__issue_active.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
       "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sass__inst_executed_local_loads")


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "details", "--csv"))))
    with open(out + "_details.csv", "w") as f:
        f.write("launch_id,kernel,metric,unit,value\n")
        for r in rows[1:]:
            if len(r) > 14 and r[12] in DETAILS:
                f.write(",".join([r[0], r[4][:48].replace(",", ";"), r[12], r[13], r[14]]) + "\n")
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    if rows:
        hdr, units = rows[0], rows[1]
        with open(out + "_raw.csv", "w") as f:
            f.write("metric,unit," + ",".join(f"launch{i}" for i in range(len(rows) - 2)) + "\n")
            for i, h in enumerate(hdr):
                if h in RAW or ("pcsamp_warps_issue_stalled" in h and "not_issued" not in h):
                    f.write(",".join([h, units[i]] + [r[i] for r in rows[2:]]) + "\n")
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv", "--print-source", "sass,cuda"))))
    agg, src, fn = {}, {}, ""
    for r in rows:
        if len(r) >= 2 and r[0] == "Function Name":
            fn = r[1]
        elif len(r) > 8 and r[2] == "-" and r[0].isdigit() and r[6].isdigit() and int(r[6]) > 0:
            k = (fn[:60], int(r[0]))
            agg[k] = agg.get(k, 0) + int(r[6])
            src[k] = r[1].strip()
    tot = sum(agg.values()) or 1
    with open(out + "_hot_lines.txt", "w") as f:
        f.write(f"warp-stall samples per source line (all captured launches), total {tot}\n")
        for k, v in sorted(agg.items(), key=lambda x: -x[1])[:40]:
            f.write(f"{v:6d} {100.0 * v / tot:5.1f}%  {k[0]}:{k[1]}  {src[k][:110]}\n")
    print("wrote", out + "_{details.csv,raw.csv,hot_lines.txt}")


if __name__ == "__main__":
    main()
