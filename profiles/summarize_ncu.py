#!/usr/bin/env python
"""Summarise an ncu report (.ncu-rep) of one kernel into the markdown kept under profiles/.

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep [rows_per_launch] [algorithmic_bytes_per_row]
"""
import csv
import io
import json
import subprocess
import sys
from collections import Counter


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    rows_per_launch = int(sys.argv[2]) if len(sys.argv) > 2 else None
    bpr = float(sys.argv[3]) if len(sys.argv) > 3 else None
    raw = ncu_csv(rep, "raw")
    hdr, units, vals = raw[0], raw[1], raw[2]
    m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    keys = ["Kernel Name", "Grid Size", "Block Size", "launch__registers_per_thread", "gpu__time_duration.sum",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
            "launch__shared_mem_per_block_dynamic", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
    print(f"# ncu summary: {rep.split('/')[-1]}\n")
    print("| metric | value | unit |\n|---|---|---|")
    for k in keys:
        if k in m:
            print(f"| {k} | {m[k][0]} | {m[k][1]} |")
    rd = float(m["dram__bytes_read.sum"][0]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[m["dram__bytes_read.sum"][1]]
    wr = float(m["dram__bytes_write.sum"][0]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[m["dram__bytes_write.sum"][1]]
    print(f"\nDRAM traffic per launch: {rd + wr:.4g} B (read {rd:.4g} + write {wr:.4g})")
    if rows_per_launch and bpr:
        alg = rows_per_launch * bpr
        print(f"algorithmic bytes per launch: {alg:.4g} B ({rows_per_launch} rows x {bpr} B) -> traffic / algorithmic = {(rd + wr) / alg:.3f}")
        print("TRAFFIC_JSON " + json.dumps({"dram_bytes_per_launch": rd + wr, "rows_per_launch": rows_per_launch,
                                            "source": rep.split('/')[-1] + " (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)"}))
    print("\nwarp stall reasons (per issue-active cycle):\n")
    for h in hdr:
        if "issue_stalled" in h and "per_issue_active" in h and float(m[h][0] or 0) > 0.1:
            print(f"- {h.split('issue_stalled_')[1].split('_per_issue')[0]}: {float(m[h][0]):.2f}")
    src = ncu_csv(rep, "source")
    h2 = src[1]
    isrc, ie = h2.index("Source"), h2.index("Instructions Executed")
    c = Counter()
    for r in src[2:]:
        if len(r) > ie and r[ie] not in ("", "0"):
            op = r[isrc].split()[0] if not r[isrc].startswith("@") else r[isrc].split()[1]
            c[op.split(".")[0]] += int(r[ie])
    tot = sum(c.values())
    print(f"\nSASS opcode mix ({tot} warp instructions): " + ", ".join(f"{o} {100 * n / tot:.1f}%" for o, n in c.most_common(14)))
    tma = [o for o in c if o.startswith(("UBLKCP", "UTMA", "SYNCS"))]
    print("TMA / mbarrier SASS present: " + ", ".join(f"{o} x{c[o]}" for o in tma))


if __name__ == "__main__":
    main()
