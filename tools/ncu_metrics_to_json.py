"""ncu report(s) -> profiles/r2_kernel_metrics.json: the figures bench.py's roofline reads (DRAM bytes, duration, steps)
for the dominant kernels, with the git revision they were captured at.

  python tools/ncu_metrics_to_json.py k_epoch=gpurun_out/r2_k_epoch.ncu-rep:12 [k_wide_gemm=...:1] > profiles/r2_kernel_metrics.json

`name=report:steps` -- steps = optimiser steps the captured launch ran (for per-step figures)."""
import csv, io, json, subprocess, sys


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def num(d, key, unit_scale=None):
    v, u = d[key]
    x = float(v.replace(",", ""))
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(u, 1.0)
    return x * scale


git = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
res = {}
for arg in sys.argv[1:]:
    name, rest = arg.split("=")
    rep, steps = rest.rsplit(":", 1)
    d = raw(rep)
    res[name] = {
        "kernel": d["Kernel Name"][0], "report": rep, "git": git, "steps": int(steps),
        "grid": d["launch__grid_size"][0], "block": d["launch__block_size"][0], "registers": d["launch__registers_per_thread"][0],
        "duration_us": num(d, "gpu__time_duration.sum"),
        "dram_bytes_read": num(d, "dram__bytes_read.sum"), "dram_bytes_write": num(d, "dram__bytes_write.sum"),
        "issue_active_pct": float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"][0]),
        "smem_wavefronts_pct_of_peak": float(d["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"][0]),
    }
print(json.dumps(res, indent=1))
