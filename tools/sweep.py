#!/usr/bin/env python
"""SURVEY 8(d) cfg-5 sweep of the CoMatch head (forward + backward, no EMA) on one GPU: B_u x K x {bf16, fp32}, the
graph-replayed step time and unlabeled samples/s, with the reference's CPU path (oracle port, fp32, all host threads) timed
beside every cell in the same run.  One JSON line per cell.

    python tools/sweep.py [--diag] [--no-cpu] [--dtypes bf16,f32]
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from endoscopy_image_classification_b200 import synthetic as S  # noqa: E402
from endoscopy_image_classification_b200.comatch_head import CoMatchHead  # noqa: E402
from endoscopy_image_classification_b200.graphs import GraphedStep  # noqa: E402

ROWS = [448, 896, 1792, 3584, 7168, 14336]
BANK = [2560, 8192, 16384, 32768, 65536]
ap = argparse.ArgumentParser()
ap.add_argument("--diag", action="store_true", help="only the diagonal of the grid (+ the two corners)")
ap.add_argument("--no-cpu", action="store_true")
ap.add_argument("--dtypes", default="bf16,f32")
ap.add_argument("--cpu-budget", type=float, default=2.0, help="seconds of CPU timing per cell (at least one step)")
a = ap.parse_args()
grid = ([(448, 2560), (896, 8192), (1792, 16384), (3584, 32768), (7168, 65536), (14336, 65536), (448, 65536), (14336, 2560)]
        if a.diag else [(r, k) for r in ROWS for k in BANK])
dev = torch.device("cuda:0")
C, D = 23, 64
keys = ["logits_u_w", "logits_u_s0", "feats_u_w", "feats_u_s0", "feats_u_s1", "feats_x", "targets_x"]
protos = S.rownorm(torch.randn(C, D, generator=torch.Generator().manual_seed(99)))
one = torch.ones((), device=dev)
cores = os.cpu_count() or 1
torch.set_num_threads(cores)
cpu_cache = {}


def cpu_us(rows, K, batch):
    """The oracle's restatement of comatch.py:162-220 (+ backward) on the same fp32 inputs; best of the steps that fit the budget."""
    if (rows, K) in cpu_cache:
        return cpu_cache[(rows, K)]
    from oracle import ssl_oracle as O
    g = torch.Generator().manual_seed(1)
    state = O.CoMatchState.zeros(K, D, C)
    state.queue_feats.copy_(S.rownorm(torch.randn(K, D, generator=g)))
    state.queue_probs.copy_(torch.softmax(torch.randn(K, C, generator=g), 1))
    b = {k: (v.float() if v.is_floating_point() else v) for k, v in batch.items()}
    best, t_end = float("inf"), time.perf_counter() + a.cpu_budget
    O.comatch_head(state, **b, thr=0.9, num_classes=C, enqueue_mode="always")          # warm-up
    while True:
        t0 = time.perf_counter()
        O.comatch_head(state, **b, thr=0.9, num_classes=C, enqueue_mode="always")
        best = min(best, time.perf_counter() - t0)
        if time.perf_counter() > t_end:
            break
    cpu_cache[(rows, K)] = best * 1e6
    return cpu_cache[(rows, K)]


for dt_name in a.dtypes.split(","):
    dt = torch.bfloat16 if dt_name == "bf16" else torch.float32
    for rows, K in grid:
        B = rows // 7
        if B + rows > K:
            continue
        g = torch.Generator().manual_seed(0)
        cpu_batch = {k: v for k, v in S.comatch_step_inputs(g, B, 7, D, C, protos, dt).items() if k in keys}
        batch = {k: v.to(dev) for k, v in cpu_batch.items()}
        head = CoMatchHead(C, D, K, 0.9, enqueue_mode="always", device=dev, dtype=dt)
        head.queue_feats.copy_(S.rownorm(torch.randn(K, D, generator=g)).to(dt))
        qp = torch.softmax(torch.randn(K, C, generator=g), 1).to(dt)
        head.queue_probs.copy_(qp)
        if head.queue_probs_t is not None:
            head.queue_probs_t[:C].copy_(qp.t())

        def step(b):
            for k in ("logits_u_s0", "feats_u_s0", "feats_u_s1"):
                b[k].grad = None
                b[k].requires_grad_(True)
            total = head.total_loss(*[b[k] for k in keys], lambda_u=1.0, lambda_c=1.0)[0]
            total.backward(gradient=one)
            return total

        gs = GraphedStep(step, batch, dev, warmup=3, on_replay=lambda: head.note_graph_replay(B + rows),
                         after_capture=head.sync_ptr_from_device, capture_host_io=False)
        heavy = rows * K > 1792 * 16384
        n = (10 if dt_name == "f32" else 50) if heavy else 200
        for _ in range(3):
            gs.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            gs.replay()
        e1.record()
        torch.cuda.synchronize()
        us = 1e3 * e0.elapsed_time(e1) / n
        flop = rows * (2.0 * K * (D + C) + 430.0 * rows)               # K3 + K6 fwd+bwd, no-recompute count
        line = {"dtype": dt_name, "rows": rows, "bank": K, "us_per_step": round(us, 2), "samples_per_s": round(rows / us * 1e6),
                "algorithmic_tflops": round(flop / us * 1e-6, 2)}
        if not a.no_cpu:
            c = cpu_us(rows, K, cpu_batch)
            line.update({"cpu_us_per_step": round(c, 1), "cpu_samples_per_s": round(rows / c * 1e6), "cpu_threads": cores,
                         "gpu_over_cpu": round(c / us, 1)})
        print(json.dumps(line), flush=True)
        del gs, head
