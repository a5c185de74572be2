"""
Text summary of an ncu report (.ncu-rep): the metrics the design notes cite, per kernel launch, plus the SASS opcode
histogram of the profiled kernel (DMMA / UBLKCP / LDS / LDL / STL ...).  Runs wherever `ncu` is installed - on the GPU box
right after the capture, so that only this summary (a few KB) has to travel back.

    python tools/ncu_summary.py gpurun_out/prof_n50_r02a.ncu-rep [--traffic-key n50] > profiles/r02_ncu_n50_mma2.txt
"""
import argparse
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed_pipe_fp64_op_dmma.sum", "sm__inst_executed_pipe_fp64_op_dmma.sum",
    "smsp__pipe_fp64_op_dmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_op_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__sass_inst_executed_op_local_ld.sum",
    "smsp__sass_inst_executed_op_local_st.sum", "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_shared_st.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
]
STALLS = re.compile(r"smsp__average_warps?_issue_stalled_(\w+)_per_issue_active\.ratio|smsp__average_warp_latency_issue_stalled_(\w+)\.ratio")


def ncu(*args):
    return subprocess.run(["ncu", *args], capture_output=True, text=True).stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--traffic-key", default=None, help="also record dram bytes per launch under this key in profiles/ncu_traffic.json")
    ap.add_argument("--source", default=None, help="name stored as the source of the traffic figure")
    a = ap.parse_args()
    raw = ncu("-i", a.report, "--page", "raw", "--csv")
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        print("no data in", a.report)
        sys.exit(1)
    head, units = rows[0], rows[1]
    print(f"# ncu summary of {os.path.basename(a.report)}  (ncu --set full --clock-control none; per-launch values, cold-cache and serialised)")
    for r in rows[2:]:
        rec = dict(zip(head, r))
        unit = dict(zip(head, units))
        print(f"\nkernel: {rec.get('Kernel Name', '?')}   grid {rec.get('Grid Size', '?')} block {rec.get('Block Size', '?')}")
        for m in METRICS:
            if m in rec and rec[m] != "":
                print(f"  {m:75s} {rec[m]:>18s} {unit.get(m, '')}")
        stalls = []
        for k, v in rec.items():
            mt = STALLS.match(k)
            if mt and v not in ("", "0"):
                try:
                    stalls.append((float(v.replace(",", "")), k))
                except ValueError:
                    pass
        for v, k in sorted(stalls, reverse=True)[:8]:
            print(f"  {k:75s} {v:18.3f}")
        if a.traffic_key:
            try:
                rd = float(rec["dram__bytes_read.sum"].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit["dram__bytes_read.sum"], 1)
                wr = float(rec["dram__bytes_write.sum"].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit["dram__bytes_write.sum"], 1)
                path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
                tr = json.load(open(path)) if os.path.exists(path) else {}
                tr[a.traffic_key] = {"bytes": int(rd + wr), "source": a.source or os.path.basename(a.report), "kernel": rec.get("Kernel Name", "?")}
                json.dump(tr, open(path, "w"), indent=1)
                print(f"  dram bytes per launch (read + write): {int(rd + wr)}")
            except Exception as err:  # noqa: BLE001
                print("  traffic not recorded:", err)
    # SASS opcode histogram of the first profiled kernel
    src = ncu("-i", a.report, "--page", "source", "--csv", "--print-source", "sass")
    ops = {}
    for r in csv.reader(io.StringIO(src)):
        if len(r) > 2:
            for cell in r[:4]:
                mt = re.match(r"\s*(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_]+)*)\s", cell + " ")
                if mt and not cell.startswith("Address") and any(ch.isupper() for ch in mt.group(1)[:2]) and ("R" in cell or ";" in cell or "[" in cell):
                    ops[mt.group(1).split(".")[0]] = ops.get(mt.group(1).split(".")[0], 0) + 1
                    break
    if ops:
        keep = sorted(ops.items(), key=lambda kv: -kv[1])
        print("\nSASS opcode histogram (static instruction count of the profiled kernel):")
        print("  " + "  ".join(f"{k}:{v}" for k, v in keep[:40]))


if __name__ == "__main__":
    main()
