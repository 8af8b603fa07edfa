#!/usr/bin/env python
"""Summarise ncu output for profiles/: a launch list (gpu__time_duration CSV) or a
`--set full` report (.ncu-rep, read with `ncu -i ... --page raw --csv`).

    python tools/ncu_summary.py launches gpurun_out/launches.csv > profiles/rN_launches.md
    python tools/ncu_summary.py full gpurun_out/prof.ncu-rep      > profiles/rN_full.md
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
]


def to_ms(v, unit):
    v = float(v.replace(",", ""))
    return {"ns": v / 1e6, "us": v / 1e3, "ms": v, "s": v * 1e3}.get(unit, v)


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    def short(name):
        # tdoa::<unnamed>::k_x(const tdoa::SigJob *) -> k_x ; torch's (setup-only) kernels are dropped
        m = re.search(r"tdoa::(?:<unnamed>::)?(k_\w+)", name)
        return m.group(1) if m else None
    seq = [(short(r["Kernel Name"]), r["Grid Size"], to_ms(r["Metric Value"], r["Metric Unit"]))
           for r in csv.DictReader(lines)]
    other = sum(ms for name, _, ms in seq if name is None)
    seq = [x for x in seq if x[0] is not None]
    agg = collections.OrderedDict()
    for name, grid, ms in seq:
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    total = sum(a[1] for a in agg.values())
    print(f"# ncu launch list: {path}\n\n{len(seq)} launches, {total:.3f} ms of kernel time "
          "(cold-cache, serialised: compare shares); "
          f"torch kernels of the synthetic-capture setup, outside the timed region: {other:.3f} ms, not listed\n")
    print("| kernel | launches | total ms | share | avg ms |\n|---|---:|---:|---:|---:|")
    for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {name} | {n} | {ms:.3f} | {100 * ms / total:.1f}% | {ms / n:.3f} |")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu --set full summary: {path}\n")
    for r in data:
        m = re.search(r"(k_\w+)", r[idx["Kernel Name"]])
        name = m.group(1) if m else r[idx["Kernel Name"]]
        print(f"## {name}\n")
        for k in KEYS:
            if k in idx:
                print(f"- {k}: {r[idx[k]]} {units[idx[k]]}")
        rd = float(r[idx["dram__bytes_read.sum"]].replace(",", ""))
        wr = float(r[idx["dram__bytes_write.sum"]].replace(",", ""))
        mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        tot = rd * mult[units[idx["dram__bytes_read.sum"]]] + wr * mult[units[idx["dram__bytes_write.sum"]]]
        ms = to_ms(r[idx["gpu__time_duration.sum"]], units[idx["gpu__time_duration.sum"]])
        print(f"- traffic (dram read+write): {tot / 1e9:.4f} GB -> {tot / 1e9 / (ms / 1e3):.0f} GB/s under ncu\n")


def traffic(path):
    """profiles/traffic.json: dram read+write bytes per launch of each kernel, averaged over the
    captured launches (bench.py puts it into roofline.traffic)."""
    import json
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    acc = collections.OrderedDict()
    for r in data:
        m = re.search(r"(k_\w+)", r[idx["Kernel Name"]])
        if not m:
            continue
        tot = sum(float(r[idx[k]].replace(",", "")) * mult[units[idx[k]]]
                  for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        acc.setdefault(m.group(1), []).append(tot)
    print(json.dumps({k: sum(v) / len(v) for k, v in acc.items()}, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](sys.argv[2])
