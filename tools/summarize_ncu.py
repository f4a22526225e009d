#!/usr/bin/env python
"""Turn the ncu outputs of tools/gpu_round.sh into the small text summaries committed under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/<tag>/launches.csv  [--last-step-kernels N]  > profiles/<name>.md
  python tools/summarize_ncu.py full     gpurun_out/<tag>/attn_full.ncu-rep                       > profiles/<name>.md
  python tools/summarize_ncu.py traffic  gpurun_out/<tag>/attn_full.ncu-rep  (writes profiles/attn_traffic.json)
"""
import csv
import io
import json
import os
import re
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def short(name: str) -> str:
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"<.*", "", name)
    name = name.replace("void ", "")
    return name.strip()[:70]


def launches(path: str):
    rows = []
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(io.StringIO("".join(lines))):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1e3 if unit == "ns" else (v if unit == "us" else v * 1e3 if unit == "ms" else v)
        rows.append((int(r["ID"]), r["Kernel Name"], r["Grid Size"], r["Block Size"], us))
    # the last warm step = everything after the (n_steps-1)-th mask re-sampling; simpler: take the last 1/4 of the
    # attention launches' span.  bench.py ran 3 warm-up steps + 1 timed step => 4 identical steps.
    attn_ids = [i for i, (_, k, *_r) in enumerate(rows) if "csa_attn_kernel" in k]
    # one step = the last ATTN_PER_STEP attention launches (36 for the reference placement, 70 for --placement all)
    per_step = min(int(os.environ.get("ATTN_PER_STEP", "36")), len(attn_ids))
    first = attn_ids[-per_step] if per_step else 0
    # include the projections that precede the first attention launch of the step (q and k|v GEMMs)
    first = max(0, first - 2)
    step = rows[first:]
    agg = OrderedDict()
    for _, k, grid, block, us in step:
        key = short(k)
        a = agg.setdefault(key, [0, 0.0, grid, block])
        a[0] += 1
        a[1] += us
    total = sum(a[1] for a in agg.values())
    print(f"# ncu launch list — last warm step of `bench.py --steps 1 --warmup 3` ({len(step)} launches, "
          f"{total / 1e3:.3f} ms summed device time; cold-cache, serialised: compare SHARES)\n")
    print("| kernel | launches | total us | share | avg us | grid | block |")
    print("|---|---|---|---|---|---|---|")
    for k, (n, us, grid, block) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {us:.1f} | {us / total * 100:.1f}% | {us / n:.1f} | {grid} | {block} |")
    print("\nPer-launch list of the step (id, kernel, us):\n")
    print("```")
    for i, k, grid, block, us in step:
        print(f"{i:5d}  {short(k):<70s} {us:9.1f}")
    print("```")


RAW_METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_selected_per_issue_active.ratio",
]


def raw_table(rep: str):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    return hdr, units, data


def full(rep: str):
    hdr, units, data = raw_table(rep)
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu --set full --clock-control none: {os.path.basename(rep)} ({len(data)} launches)\n")
    print("| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(data))) + " |")
    print("|---|---|" + "---|" * len(data))
    print("| kernel | | " + " | ".join(short(r[col["Kernel Name"]]) for r in data) + " |")
    for m in RAW_METRICS:
        if m in col:
            i = col[m]
            print(f"| `{m}` | {units[i]} | " + " | ".join(r[i] for r in data) + " |")


def traffic(rep: str):
    hdr, units, data = raw_table(rep)
    col = {h: i for i, h in enumerate(hdr)}

    def to_bytes(v, u):
        v = float(v.replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]

    per = []
    for r in data:
        rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
        wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
        us = float(r[col["gpu__time_duration.sum"]].replace(",", ""))
        per.append({"dram_read": rd, "dram_write": wr, "duration": us, "duration_unit": units[col["gpu__time_duration.sum"]]})
    # bench step = 30 launches of the first class + 6 of the second (capture holds 2 + 2)
    half = len(per) // 2
    small = sum(p["dram_read"] + p["dram_write"] for p in per[:half]) / max(1, half)
    big = sum(p["dram_read"] + p["dram_write"] for p in per[half:]) / max(1, len(per) - half)
    res = {"source": os.path.relpath(rep, ROOT), "per_launch": per,
           "dram_bytes_32x32_layer": small, "dram_bytes_64x64_layer": big,
           "dram_bytes_per_launch": (30 * small + 6 * big) / 36,
           "note": "dram__bytes_read.sum + dram__bytes_write.sum per attention launch; weighted over the 30+6 "
                   "launches of one bench step"}
    with open(os.path.join(ROOT, "profiles", "attn_traffic.json"), "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](sys.argv[2])
