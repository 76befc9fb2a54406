#!/usr/bin/env python
"""Summarise an `ncu --set full` report (read here, on the CPU box) into the files that are committed
under profiles/:

    python tools/ncu_summary.py gpurun_out/prof_r1.ncu-rep profiles/r01_seq_r1   [--traffic-key name=regex ...]

writes <out>.md (a table of the metrics the roofline argument needs, per profiled launch), <out>.csv (the
same, machine readable) and, with --traffic-key, updates profiles/traffic.json: DRAM bytes (read + write) per
launch of the kernel whose name matches the regex -- bench.py copies that into `roofline.traffic`.
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu (MUFU) pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu pipe %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor instructions"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs), CTAs/SM"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "global load sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "global load requests"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "global store sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "global store requests"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
]
STALL = re.compile(r"smsp__average_warps?_issue_stalled_(\w+)_per_issue_active\.ratio|smsp__average_warp_latency_issue_stalled_(\w+)\.ratio")


def to_bytes(v, unit):
    v = float(v)
    u = unit.lower()
    for k, m in (("gbyte", 1e9), ("mbyte", 1e6), ("kbyte", 1e3), ("byte", 1.0)):
        if u.startswith(k):
            return v * m
    return v


def main():
    rep, out = sys.argv[1], sys.argv[2]
    tkeys = {}
    if "--traffic-key" in sys.argv:
        for kv in sys.argv[sys.argv.index("--traffic-key") + 1:]:
            k, rx = kv.split("=", 1)
            tkeys[k] = re.compile(rx)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    md = [f"# ncu --set full summary of `{os.path.basename(rep)}`", "",
          "Captured with `--clock-control none --import-source on`; values are per profiled launch "
          "(serialised, cold cache: compare shares and ratios, not absolute times).", ""]
    table = []
    traffic = {}
    per_name = {}
    for r in body:
        name = r[col["Kernel Name"]]
        md += [f"## `{name}`", "", "| metric | value | unit |", "|---|---|---|"]
        rec = {"kernel": name}
        for m, label in METRICS:
            if m in col and r[col[m]] != "":
                md.append(f"| {label} (`{m}`) | {r[col[m]]} | {units[col[m]]} |")
                rec[m] = r[col[m]]
        stalls = []
        for h, i in col.items():
            mm = STALL.match(h)
            if mm and r[i] not in ("", "n/a"):
                try:
                    stalls.append((float(r[i]), mm.group(1) or mm.group(2)))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        if stalls:
            md += ["", "Top warp-stall reasons (cycles stalled per issued instruction): " +
                   ", ".join(f"{n} {v:.2f}" for v, n in stalls[:6])]
        if "dram__bytes_read.sum" in col:
            rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
            wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
            rec["dram_bytes_total"] = rd + wr
            md += ["", f"DRAM traffic per launch: {rd + wr:.4g} B (read {rd:.4g} + write {wr:.4g})"]
            for k, rx in tkeys.items():           # a key may cover several kernels of one call: sum over distinct names
                if rx.search(name):
                    per_name.setdefault(k, {})[name] = rd + wr
                    traffic[k] = sum(per_name[k].values())
        md.append("")
        table.append(rec)
    with open(out + ".md", "w") as f:
        f.write("\n".join(md))
    keys = sorted({k for rec in table for k in rec})
    with open(out + ".csv", "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=keys)
        w.writeheader()
        w.writerows(table)
    if traffic:
        path = os.path.join(os.path.dirname(out) or ".", "traffic.json")
        cur = {}
        if os.path.exists(path):
            cur = json.load(open(path))
        cur.update(traffic)
        cur["_source"] = os.path.basename(out) + ".md"
        json.dump(cur, open(path, "w"), indent=1)
    print(f"wrote {out}.md / .csv" + (f" and traffic.json {traffic}" if traffic else ""))


if __name__ == "__main__":
    main()
