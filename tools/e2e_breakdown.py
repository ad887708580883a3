#!/usr/bin/env python
"""Where the end-to-end step goes (bench.py's e2e: stage -> one graph launch [H2D, head fwd+bwd || EMA, D2H] -> sync):
host-side split of `GraphedStep.replay_host` over N steps of BASELINE cfg 2 on one GPU.

    python tools/e2e_breakdown.py [N]
"""
import ctypes
import json
import statistics
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

import bench  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
ctx = bench.Ctx(0, 1, 0)
case = bench.Case(ctx, bench.WORKLOADS["cfg2"])
gs = case.capture()
plain = [{k: v.clone() for k, v in hb.items()} for hb in case.host]
stream = torch.cuda.current_stream(ctx.dev)
for i in range(20):
    gs.replay_host(plain[i % 4])
rows = []
for i in range(n):
    b = plain[i % 4]
    t0 = time.perf_counter()
    for k, dst, nbytes, dtype, shape in gs._stage_plan:
        ctypes.memmove(dst, b[k].data_ptr(), nbytes)
    t1 = time.perf_counter()
    gs.graph_host.replay()
    t2 = time.perf_counter()
    gs._note()
    stream.synchronize()
    t3 = time.perf_counter()
    v = float(gs.result_host[0])
    t4 = time.perf_counter()
    rows.append((t1 - t0, t2 - t1, t3 - t2, t4 - t3, t4 - t0))
med = [1e6 * statistics.median(r[i] for r in rows) for i in range(5)]
print(json.dumps({"steps": n, "us_stage_7_tensors_into_pinned": round(med[0], 1), "us_graph_launch_call": round(med[1], 1),
                  "us_wait_for_the_stream": round(med[2], 1), "us_read_the_loss": round(med[3], 1), "us_total": round(med[4], 1),
                  "gpu_step_us_device_resident": None}))
