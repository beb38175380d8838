#!/usr/bin/env python
"""Turn an `ncu --set full` capture of benchmarks/ncu_kernels.py into the markdown table kept
under profiles/ (run in the build container: `ncu -i` reads the report without a GPU).

    python benchmarks/ncu_summary.py gpurun_out/prof_kernels_r01.ncu-rep gpurun_out/kernels_ms_r01.json > profiles/r01_kernels_ncu.md
"""

from __future__ import annotations

import csv
import io
import json
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread",
    "lts__t_sector_hit_rate.pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
]


def to_bytes(value: str, unit: str) -> float:
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]
    return float(value.replace(",", "")) * scale


def to_ms(value: str, unit: str) -> float:
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[unit]
    return float(value.replace(",", "")) * scale


def short(name: str) -> str:
    name = name.replace("void ", "").replace("at::", "")
    return name.split("(")[0]


def main():
    rep, timings = sys.argv[1], json.load(open(sys.argv[2]))
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", ",".join(METRICS)], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units, data = rows[0], rows[1], rows[2:]
    ix = {n: i for i, n in enumerate(head)}
    alg, ms = timings["algorithmic_bytes_per_launch"], timings["ms"]
    peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]

    def key_of(kernel: str) -> str | None:
        k = short(kernel)
        for name in alg:
            base = name.split("<")[0]
            if not k.startswith(base):
                continue
            if "<" not in name:
                return name
            want = [w.strip() for w in name.split("<")[1].rstrip(">").split(",")]
            have = [h.strip() for h in (k.split("<")[1].rstrip(">") if "<" in k else "").split(",")]
            if have[: len(want)] == want:
                return name
        return None

    print("# Round 1 — `ncu --set full` of every hot kernel at BASELINE sizes\n")
    print("Program: `benchmarks/ncu_kernels.py` (one launch per kernel under ncu; the same launches timed with CUDA")
    print("events, warm, best of 5, in a separate plain run — the `events` column).  ncu's own durations are")
    print("cold-cache and serialised.  `algorithmic` = the bytes of DESIGN.md §3 for that launch; `traffic` =")
    print(f"`dram__bytes_read.sum + dram__bytes_write.sum`; roofline fraction = algorithmic / events time / {peak} GB/s")
    print("(MEASURED_PEAKS.json).\n")
    print("| kernel | regs | ncu ms | events ms | algorithmic GB | traffic GB (read + write) | traffic / alg | alg GB/s | frac of measured HBM peak | DRAM % of ncu peak | SM % | issue-slot % | warps active % | L2 hit % |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    seen = set()
    for r in data:
        name = r[ix["Kernel Name"]]
        key = key_of(name)
        if key is None or key in seen:
            continue
        seen.add(key)
        dur = to_ms(r[ix["gpu__time_duration.sum"]], units[ix["gpu__time_duration.sum"]])
        rd = to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]])
        wr = to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
        a, t = alg[key], ms[key]

        def pct(m):
            return f"{float(r[ix[m]]):.1f}"

        print(
            f"| `{short(name)}` | {r[ix['launch__registers_per_thread']]} | {dur:.3f} | {t:.3f} | {a / 1e9:.3f} | {(rd + wr) / 1e9:.3f} ({rd / 1e9:.3f} + {wr / 1e9:.3f}) | {(rd + wr) / a:.2f} | "
            f"{a / t / 1e6:.0f} | {a / t / 1e6 / peak:.2f} | {pct('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')} | {pct('sm__throughput.avg.pct_of_peak_sustained_elapsed')} | "
            f"{pct('smsp__issue_active.avg.pct_of_peak_sustained_active')} | {pct('sm__warps_active.avg.pct_of_peak_sustained_active')} | {pct('lts__t_sector_hit_rate.pct')} |"
        )


if __name__ == "__main__":
    main()
