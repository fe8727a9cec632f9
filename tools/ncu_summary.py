#!/usr/bin/env python3
"""Turn the ncu artefacts a gpurun call left in gpurun_out/ into the tracked summaries under profiles/.

  python tools/ncu_summary.py <tag>      e.g. r01
reads  gpurun_out/launches_<tag>.csv, gpurun_out/prof_keccak_<tag>.ncu-rep, gpurun_out/prof_walk_<tag>.ncu-rep
writes profiles/launches_<tag>.md, profiles/keccak_<tag>.md, profiles/walk_<tag>.md, profiles/keccak_traffic.json
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
GO = os.path.join(ROOT, "gpurun_out")

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__grid_size",
    "launch__block_size", "sm__cycles_elapsed.avg", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def launches(tag):
    path = os.path.join(GO, f"launches_{tag}.csv")
    if not os.path.exists(path):
        return
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Metric Name")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3}.get(r[ui], 1e-6)
        name = r[ki].split("(")[0].replace("void ", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    ours = sum(v[1] for k, v in agg.items() if "phant::" in k and "synth" not in k)
    with open(os.path.join(OUT, f"launches_{tag}.md"), "w") as f:
        f.write(f"# ncu launch list, {tag}\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` over `python bench.py --steps 2 --warmup 3 --no-cpu --skip-extras`\n"
                "(cold-cache, serialised launches: compare SHARES, not absolutes).  `synth_c2_kernel` is the untimed setup\n"
                "that generates the witness in HBM; shares below are of the hot path (everything except synth).\n\n")
        f.write("| kernel | launches | total ms | share of hot path |\n|---|---:|---:|---:|\n")
        hot = tot - sum(v[1] for k, v in agg.items() if "synth" in k)
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            share = "setup" if "synth" in k else f"{100 * v[1] / hot:.1f}%"
            f.write(f"| `{k[:90]}` | {v[0]} | {v[1]:.3f} | {share} |\n")
        f.write(f"\nhand-written phant kernels: {100 * ours / hot:.1f}% of hot-path device time; the rest is torch fills / copies (round 2: no library sort on the hot path).\n")


def report(tag, which, title):
    rep = os.path.join(GO, f"prof_{which}_{tag}.ncu-rep")
    if not os.path.exists(rep):
        return None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    with open(os.path.join(OUT, f"{which}_{tag}.md"), "w") as f:
        f.write(f"# {title} -- `ncu --set full --clock-control none --import-source on`, {tag}\n\n")
        res = None
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            f.write(f"## launch: `{name[:100]}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            for m in METRICS:
                if m in hdr:
                    f.write(f"| {m} | {r[hdr.index(m)]} | {units[hdr.index(m)]} |\n")
            st = sorted(((float(r[hdr.index(s)] or 0), s) for s in stalls), reverse=True)[:6]
            f.write("\nTop issue-stall reasons (warps per issue-active cycle):\n\n")
            for v, s in st:
                f.write(f"- {s.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}: {v:.2f}\n")
            f.write("\n")
            rd = float(r[hdr.index("dram__bytes_read.sum")].replace(",", ""))
            wr = float(r[hdr.index("dram__bytes_write.sum")].replace(",", ""))
            mul = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            res = rd * mul[units[hdr.index("dram__bytes_read.sum")]] + wr * mul[units[hdr.index("dram__bytes_write.sum")]]
        return res


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(OUT, exist_ok=True)
    launches(tag)
    t = report(tag, "keccak", "batched Keccak kernel (keccak256_staged_kernel)")
    if t:
        json.dump({"dram_bytes_per_launch": t, "source": f"profiles/keccak_{tag}.md (dram__bytes_read.sum + dram__bytes_write.sum, one launch = 1M proofs)"},
                  open(os.path.join(OUT, "keccak_traffic.json"), "w"))
    report(tag, "walk", "proof-walk kernel (walk_kernel)")
    import glob
    for rep in sorted(glob.glob(os.path.join(GO, f"prof_*_{tag}.ncu-rep"))):
        which = os.path.basename(rep)[len("prof_"):-len(f"_{tag}.ncu-rep")]
        if which not in ("keccak", "walk"):
            report(tag, which, f"kernel capture `{which}`")
    print(os.listdir(OUT))
