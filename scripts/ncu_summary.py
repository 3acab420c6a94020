#!/usr/bin/env python
"""Summarise an `ncu --set full` capture of the IBSS kernel into profiles/*.json.

usage: ncu_summary.py <report.ncu-rep> <out.json> --loci N --iters I --L L --bytes-per-ser B [--tag TEXT]

The capture is a small fixed-iteration run (ncu replays the kernel ~40 times), so what carries over to
the bench line is DRAM traffic *per single-effect update* (one traversal of the stored genotypes / LD)."""
import argparse, csv, json, subprocess, sys

ap = argparse.ArgumentParser()
ap.add_argument("report"); ap.add_argument("out")
ap.add_argument("--loci", type=int, required=True); ap.add_argument("--iters", type=int, required=True)
ap.add_argument("--L", type=int, default=10); ap.add_argument("--bytes-per-ser", type=float, required=True)
ap.add_argument("--first-traversal", type=int, default=1, help="1 if each locus has the extra X^T y traversal")
ap.add_argument("--tag", default="")
a = ap.parse_args()
raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u, r = rows[0], rows[1], rows[2]
d = dict(zip(h, r)); units = dict(zip(h, u))
def val(k, scale={"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}):
    return float(d[k].replace(",", "")) * scale.get(units[k], 1.0)
n_trav = a.loci * (a.iters * a.L + a.first_traversal)
rd, wr, ms = val("dram__bytes_read.sum"), val("dram__bytes_write.sum"), val("gpu__time_duration.sum")
out = {
    "tag": a.tag, "kernel": d.get("Kernel Name"), "report": a.report,
    "workload": f"{a.loci} loci x {a.iters} fixed iterations x L={a.L} (+{a.first_traversal} X^T y traversal per locus)",
    "gpu_time_ms_under_ncu": ms, "dram_bytes_read": rd, "dram_bytes_write": wr, "traversals": n_trav,
    "dram_bytes_per_ser": (rd + wr) / n_trav, "dram_read_bytes_per_ser": rd / n_trav,
    "algorithmic_bytes_per_ser": a.bytes_per_ser, "traffic_over_algorithmic": (rd + wr) / n_trav / a.bytes_per_ser,
    "dram_GBps_under_ncu": (rd + wr) / ms / 1e6,
    "registers_per_thread": d.get("launch__registers_per_thread"),
    "issue_active_pct": d.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "pipe_fp64_pct": d.get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
    "pipe_alu_pct": d.get("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    "pipe_lsu_pct": d.get("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
    "dram_throughput_pct": d.get("dram__throughput.avg.pct_of_peak_sustained_elapsed"),
    "warp_instructions": d.get("smsp__inst_executed.sum"),
}
json.dump(out, open(a.out, "w"), indent=1)
print(json.dumps(out))
