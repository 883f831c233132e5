#!/usr/bin/env python
"""Summarise ncu output into small tracked files under profiles/.

    python scripts/ncu_summary.py launches gpurun_out/launches.csv profiles/rNN_launches.md
    python scripts/ncu_summary.py kernel   gpurun_out/prof.ncu-rep  profiles/rNN_force.md [--id K]
    python scripts/ncu_summary.py traffic  profiles/rNN_traffic.json --agents N --head SHA rep1.ncu-rep [rep2 ...]

`launches`: the `--metrics gpu__time_duration.sum --csv --log-file` launch list -> per-kernel count,
total, average and SHARE of the listed launches (cold-cache, serialised: compare shares only).
`traffic`: the DRAM bytes (dram__bytes_read.sum, dram__bytes_write.sum) and duration of the first launch of every
kernel in the given `ncu --set full` reports -> the small JSON bench.py reads for `roofline.traffic`.
`kernel`: one `ncu --set full` report -> the counters DESIGN.md / bench.py quote (DRAM bytes, hit
rates, issue utilisation, lane efficiency, pipe utilisation, stall mix) plus the hottest source lines.
"""
from __future__ import annotations

import collections
import csv
import io
import subprocess
import sys


def read_csv_after_preamble(text: str):
    lines = text.splitlines()
    for i, ln in enumerate(lines):
        if ln.startswith('"ID"') or ln.startswith('"Kernel Name"') or ln.startswith('"Address"'):
            return lines[i:]
    raise SystemExit("no CSV header found")


def launches(src: str, dst: str) -> None:
    rows = list(csv.DictReader(io.StringIO("\n".join(read_csv_after_preamble(open(src).read())))))
    agg = collections.OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"].split("(")[0].replace("void ", "")
        ns = float(r["Metric Value"]) * {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(r["Metric Unit"], 1.0)
        a = agg.setdefault(name, [0, 0.0, r["Grid Size"], r["Block Size"]])
        a[0] += 1
        a[1] += ns
    total = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary ({src})\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none`: per-launch times are cold-cache and\n"
                "serialised, so compare the SHARE column with bench.py's CUDA-event `kernel_ms_per_step`, not absolutes.\n\n")
        f.write("| kernel | launches | total us | avg us | share | last grid | block |\n|---|---|---|---|---|---|---|\n")
        for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{name}` | {a[0]} | {a[1] / 1e3:.1f} | {a[1] / a[0] / 1e3:.1f} | {100 * a[1] / total:.1f}% | "
                    f"{a[2]} | {a[3]} |\n")
        f.write(f"\ntotal listed: {len(rows)} launches, {total / 1e6:.3f} ms\n")
    print(open(dst).read())


KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("l1tex__t_sector_hit_rate.pct", "L1 sector hit rate %"),
    ("lts__t_sector_hit_rate.pct", "L2 sector hit rate %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__occupancy_limit_registers", "occupancy limit (registers), blocks"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem), blocks"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes per warp instruction (of 32)"),
    ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue slots busy %"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / scheduler / cycle"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "FMA pipe %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "ALU pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", "XU (MUFU) pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed", "LSU pipe %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared bank conflicts"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected / issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle / issue"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch_resolving / issue"),
]


def kernel(src: str, dst: str, which: int | None) -> None:
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO("\n".join(read_csv_after_preamble(raw)))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary ({src})\n\n")
        for k, r in enumerate(data):
            if which is not None and k != which:
                continue
            f.write(f"## launch {k}: `{r[col['Kernel Name']]}`  grid {r[col['Grid Size']]} block {r[col['Block Size']]}\n\n")
            f.write("| counter | value | unit |\n|---|---|---|\n")
            for key, label in KEYS:
                if key in col:
                    f.write(f"| {label} (`{key}`) | {r[col[key]]} | {units[col[key]]} |\n")
            f.write("\n")
        # hottest source lines (needs -lineinfo + --import-source on)
        srcp = subprocess.run(["ncu", "-i", src, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                              capture_output=True, text=True)
        hot = []  # (samples, instr, lanes, file, line, text)
        fname, header, first_fn, seen_files = "?", None, None, set()
        for row in csv.reader(io.StringIO(srcp.stdout)):
            if not row:
                continue
            if row[0] == "File Path":
                fname, header = row[1].split("/")[-1], None
                if fname in seen_files:  # the next launch's tables repeat the same files
                    break
                seen_files.add(fname)
            elif row[0] == "Function Name":
                first_fn = first_fn or row[1]
                if row[1] != first_fn:
                    break
            elif row[0] == "Line No":
                header = {h: i for i, h in reversed(list(enumerate(row)))}
            elif header and row[0].strip().isdigit():
                g = lambda k: row[header[k]] if k in header and header[k] < len(row) else "0"  # noqa: E731
                samples = g("# Samples")
                hot.append((int(samples) if samples.isdigit() else 0, g("Instructions Executed"), g("Avg. Threads Executed"), fname,
                            row[0], row[1].strip()))
        tot = sum(h[0] for h in hot) or 1
        f.write(f"## hottest source lines by warp-stall samples (`{first_fn}`)\n\n")
        f.write("| samples % | warp instr | avg lanes | where | source |\n|---|---|---|---|---|\n")
        for h in sorted(hot, key=lambda h: -h[0])[:30]:
            f.write(f"| {100 * h[0] / tot:.1f} | {h[1]} | {h[2]} | {h[3]}:{h[4]} | `{h[5][:110]}` |\n")
    print(open(dst).read())


UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6,
        "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3, "second": 1e6}


def traffic(dst: str, reports, agents: int, head: str) -> None:
    import json
    kernels = {}
    for src in reports:
        raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
        rows = list(csv.reader(io.StringIO("\n".join(read_csv_after_preamble(raw)))))
        hdr, units, data = rows[0], rows[1], rows[2:]
        col = {h: i for i, h in enumerate(hdr)}
        for r in data:
            name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0].split("::")[-1]
            if name in kernels:
                continue
            val = lambda k: float(r[col[k]].replace(",", "")) * UNIT[units[col[k]]]  # noqa: E731
            kernels[name] = {"dram_read_bytes": val("dram__bytes_read.sum"), "dram_write_bytes": val("dram__bytes_write.sum"),
                             "duration_us": val("gpu__time_duration.sum"), "agents": agents,
                             "grid": r[col["Grid Size"]], "block": r[col["Block Size"]],
                             "template": r[col["Kernel Name"]], "report": src}
    doc = {"what": "per-launch DRAM traffic of the benched binary: one `ncu --set full --clock-control none` launch per "
                   "kernel, python bench.py at its default workload; agents = live pedestrians of that launch",
           "head": head, "agents_total": agents, "kernels": kernels}
    open(dst, "w").write(json.dumps(doc, indent=1) + "\n")
    print(open(dst).read())


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        a = sys.argv[2:]
        agents = int(a[a.index("--agents") + 1])
        head = a[a.index("--head") + 1] if "--head" in a else "?"
        reps = [x for k, x in enumerate(a[1:], 1) if not x.startswith("--") and a[k - 1] not in ("--agents", "--head")]
        traffic(a[0], reps, agents, head)
        sys.exit(0)
    mode, src, dst = sys.argv[1:4]
    if mode == "launches":
        launches(src, dst)
    else:
        which = int(sys.argv[sys.argv.index("--id") + 1]) if "--id" in sys.argv else None
        kernel(src, dst, which)
