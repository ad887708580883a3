#!/usr/bin/env python
"""Turn the `.ncu-rep` captures that `tools/profile_round.sh` leaves under gpurun_out/ into the text summaries kept under
profiles/ (run in the build container; `ncu -i` needs no GPU):

    python tools/ncu_extract.py r02

Per capture: `<name>.txt` with the selected raw metrics and, for the kernels whose stalls matter, the 25 hottest SASS lines
with their dominant stall reasons.  Also writes profiles/ema_traffic.json (DRAM bytes of one EMA launch), which bench.py
reports as `roofline.traffic`."""
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
R = sys.argv[1] if len(sys.argv) > 1 else "r02"
METRICS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
           "launch__shared_mem_per_block_dynamic", "sm__cycles_active.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum"]


def ncu_csv(rep, page, *extra):
    out = subprocess.run(["ncu", "-i", str(rep), "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    for rep in sorted((REPO / "gpurun_out").glob(f"{R}_ncu_*.ncu-rep")):
        raw = ncu_csv(rep, "raw")
        if len(raw) < 3:
            continue
        d = {h: (v, u) for h, u, v in zip(raw[0], raw[1], raw[2])}
        lines = [f"# {rep.name}: ncu --set full --clock-control none (one launch), read with tools/ncu_extract.py",
                 f"kernel: {d.get('Kernel Name', ('?',))[0]}"]
        for m in METRICS:
            if m in d:
                lines.append(f"{m:75s} {d[m][0]} {d[m][1]}")
        src = ncu_csv(rep, "source", "--print-source", "sass")
        if len(src) > 3:
            hdr = src[1]
            idx = {h: i for i, h in enumerate(hdr)}
            stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
            data = [r for r in src[2:] if len(r) == len(hdr)]
            tot = sum(int(r[idx["# Samples"]]) for r in data) or 1
            agg = sorted(((sum(int(r[idx[s]]) for r in data), s) for s in stalls), reverse=True)[:8]
            lines.append(f"\nwarp-state samples: {tot}; by reason: " + ", ".join(f"{s[6:]} {100 * v / tot:.1f} %" for v, s in agg))
            lines.append("hottest SASS lines (samples, share, two dominant reasons):")
            for i in sorted(sorted(range(len(data)), key=lambda i: -int(data[i][idx["# Samples"]]))[:25]):
                r = data[i]
                n = int(r[idx["# Samples"]])
                top = sorted(((int(r[idx[s]]), s) for s in stalls), reverse=True)[:2]
                lines.append(f"  {r[1].strip()[:70]:70s} {n:5d} {100 * n / tot:5.1f} %  " + " ".join(f"{s[6:]}={v}" for v, s in top if v))
        out = REPO / "profiles" / (rep.stem + ".txt")
        out.write_text("\n".join(lines) + "\n")
        print("wrote", out)
        if rep.stem.endswith("_ncu_ema") and "dram__bytes_read.sum" in d:
            scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
            rd = float(d["dram__bytes_read.sum"][0]) * scale.get(d["dram__bytes_read.sum"][1], 1.0)
            wr = float(d["dram__bytes_write.sum"][0]) * scale.get(d["dram__bytes_write.sum"][1], 1.0)
            (REPO / "profiles" / "ema_traffic.json").write_text(json.dumps(
                {"kernel": "ema_multi_tensor_kernel", "dram_bytes_read": rd, "dram_bytes_write": wr, "traffic_bytes_per_launch": rd + wr,
                 "source": f"profiles/{rep.stem}.txt (ncu --set full, one launch of the ModelwEmb-ResNet-50 update, {R})"}, indent=1) + "\n")


if __name__ == "__main__":
    main()
