#!/usr/bin/env python
"""Markdown tables for profiles/README.md from the files tools/profile_round.sh leaves in gpurun_out/ (and copies them into
profiles/):   python tools/profiles_summary.py r02 > /tmp/summary.md"""
import csv
import json
import shutil
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
R = sys.argv[1] if len(sys.argv) > 1 else "r02"
G, P = REPO / "gpurun_out", REPO / "profiles"


def last_json(path):
    lines = [l for l in open(path) if l.startswith("{")]
    return json.loads(lines[-1]) if lines else None


def jsonl(path):
    return [json.loads(l) for l in open(path) if l.startswith("{")]


keep = [f"{R}_bench_n1.json", f"{R}_bench_reference_arm.json", f"{R}_bench_cfg1.json", f"{R}_bench_cfg3.json", f"{R}_bench_n1_serial_ema.json",
        f"{R}_microbench_cfg2.txt", f"{R}_microbench_rows3584_bank65536.txt", f"{R}_microbench_cfg2_f32.txt", f"{R}_k3_tune.jsonl",
        f"{R}_k3_fp32_storage_tc.jsonl", f"{R}_k3_fp32_storage_ffma.jsonl", f"{R}_opt_bench.json", f"{R}_sweep_cfg5.jsonl",
        f"{R}_micro_tmem_mufu_sts.jsonl", f"{R}_micro_mma_issue.jsonl", f"{R}_launches_bench_cfg2.csv"]
for k in keep:
    if (G / k).exists():
        shutil.copy(G / k, P / k)

b = last_json(G / f"{R}_bench_n1.json")
ref = last_json(G / f"{R}_bench_reference_arm.json")
print("## Headline (driver contract: `python bench.py`, N = 1)\n")
print("| | value |\n|---|---|")
print(f"| device-resident (`value`) | {b['value'] / 1e6:.2f} M samples/s, {b['ms_per_step'] * 1e3:.1f} µs/step (25 blocks: min {b['config']['timing']['min_ms_per_step'] * 1e3:.1f} / max {b['config']['timing']['max_ms_per_step'] * 1e3:.1f}) |")
print(f"| end to end (`e2e`) | {b['e2e']['value'] / 1e6:.2f} M samples/s, {b['e2e']['ms_per_step'] * 1e3:.1f} µs/step, {b['e2e']['h2d_bytes_per_step']} B H2D + {b['e2e']['d2h_bytes_per_step']} B D2H per step |")
print(f"| eager (no graph) | {b['config']['eager_ms_per_step'] * 1e3:.0f} µs/step |")
s = last_json(G / f"{R}_bench_n1_serial_ema.json")
if s:
    print(f"| `--ema-overlap 0` (EMA after the backward, in series) | {s['ms_per_step'] * 1e3:.1f} µs/step, e2e {s['e2e']['ms_per_step'] * 1e3:.1f} |")
rf = b["roofline"]
print(f"| roofline (EMA kernel) | {rf['achieved']:.0f} GB/s of {rf['peak']:.0f} = {rf['frac']:.3f}; traffic {rf['traffic'] / 1e6 if rf.get('traffic') else float('nan'):.1f} MB vs {rf['algorithmic_bytes_per_launch'] / 1e6:.1f} MB algorithmic |")
print(f"| reference arm (`--impl reference`, oracle port, {ref['cpu_baseline']['cores']} host threads) | {ref['value'] / 1e3:.1f} k samples/s ({ref['ms_per_step']:.1f} ms/step) → e2e ratio {b['e2e']['value'] / ref['value']:.0f}× |")
print(f"| `cpu_baseline` of the same run | {b['cpu_baseline']['value'] / 1e3:.1f} k samples/s |")
print(f"| cfg 4 block (K = 65536, one GPU) | {b['cfg4']['ms_per_step'] * 1e3:.1f} µs/step, e2e {b['cfg4']['e2e']['ms_per_step'] * 1e3:.1f} |")
print(f"| fp32 block (the reference's storage precision, tensor-core kernels on split operands) | {b['fp32']['ms_per_step'] * 1e3:.1f} µs/step |")
fo = b["fused_opt_ema"]
print(f"| fused optimizer + EMA (f1) | {fo['fused_graph_ms'] * 1e3:.1f} µs graph-replayed / {fo['fused_eager_ms'] * 1e3:.1f} eager vs {fo['torch_fused_adam_plus_ema_kernel_ms'] * 1e3:.1f} (torch fused Adam + EMA kernel) |")
for w in ("cfg1", "cfg3"):
    x = last_json(G / f"{R}_bench_{w}.json")
    if x:
        print(f"| `--workload {w}` | {x['value'] / 1e6:.2f} M samples/s, {x['ms_per_step'] * 1e3:.1f} µs/step, e2e {x['e2e']['ms_per_step'] * 1e3:.1f} µs; CPU {x['cpu_baseline']['value'] / 1e3:.1f} k samples/s |")
print(f"| clocks during the timed regions | {b['clocks']} |\n")

print("## Step breakdown (ncu launch list of the bench command, µs per launch)\n")
rows = list(csv.reader(open(G / f"{R}_launches_bench_cfg2.csv")))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
idx = {h: i for i, h in enumerate(rows[hdr])}
agg = {}
for r in rows[hdr + 1:]:
    if len(r) <= idx["Metric Value"] or r[idx["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("b200ssl::<unnamed>::", "")
    v = float(r[idx["Metric Value"]].replace(",", ""))
    unit = r[idx["Metric Unit"]]
    v = v / 1e3 if unit in ("ns", "nsecond") else v
    agg.setdefault(name, []).append(v)
tot = sum(sum(v) / len(v) for v in agg.values())
print("| kernel | launches | µs (mean) | share |\n|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]) / len(kv[1])):
    m = sum(v) / len(v)
    print(f"| `{k}` | {len(v)} | {m:.1f} | {100 * m / tot:.1f} % |")
print(f"| sum of the means | | {tot:.1f} | |\n")

print("## cfg 5 sweep (head forward + backward without EMA, one GPU, graph replay; CPU = oracle port on 16 host threads, same run)\n")
sw = jsonl(G / f"{R}_sweep_cfg5.jsonl")
banks = sorted({r["bank"] for r in sw})
for dt in ("bf16", "f32"):
    print(f"**{dt}** — µs/step (GPU ÷ CPU speed-up)\n")
    print("| rows \\ K | " + " | ".join(str(k) for k in banks) + " |\n|---|" + "---|" * len(banks))
    for rows_ in sorted({r["rows"] for r in sw}):
        cells = []
        for k in banks:
            c = [r for r in sw if r["dtype"] == dt and r["rows"] == rows_ and r["bank"] == k]
            cells.append(f"{c[0]['us_per_step']:.0f} ({c[0]['gpu_over_cpu']:.0f}×)" if c and c[0].get("gpu_over_cpu") else (f"{c[0]['us_per_step']:.0f}" if c else "—"))
        print(f"| {rows_} | " + " | ".join(cells) + " |")
    print()
