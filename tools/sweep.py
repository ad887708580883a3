#!/usr/bin/env python
"""SURVEY 8(d) cfg-5 sweep of the CoMatch head (forward + backward, no EMA) on one GPU: graph-replayed step time and
unlabeled samples/s for B_u x K x {bf16, fp32}.  Prints one JSON line per configuration.

    python tools/sweep.py [--full]        # default: the diagonal of the grid; --full: every (B_u, K) pair
"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from endoscopy_image_classification_b200 import synthetic as S  # noqa: E402
from endoscopy_image_classification_b200.comatch_head import CoMatchHead  # noqa: E402
from endoscopy_image_classification_b200.graphs import GraphedStep  # noqa: E402

ROWS = [448, 896, 1792, 3584, 7168, 14336]
BANK = [2560, 8192, 16384, 32768, 65536]
full = "--full" in sys.argv
grid = [(r, k) for r in ROWS for k in BANK] if full else [(448, 2560), (896, 8192), (1792, 16384), (3584, 32768), (7168, 65536),
                                                          (14336, 65536), (448, 65536), (14336, 2560)]
dev = torch.device("cuda:0")
C, D = 23, 64
keys = ["logits_u_w", "logits_u_s0", "feats_u_w", "feats_u_s0", "feats_u_s1", "feats_x", "targets_x"]
protos = S.rownorm(torch.randn(C, D, generator=torch.Generator().manual_seed(99)))
one = torch.ones((), device=dev)
for dt_name, dt in (("bf16", torch.bfloat16), ("f32", torch.float32)):
    for rows, K in grid:
        if dt_name == "f32" and rows * K > 3584 * 32768:
            continue                                  # exact-fp32 FFMA path: keep the sweep short
        B = rows // 7
        g = torch.Generator().manual_seed(0)
        batch = {k: v.to(dev) for k, v in S.comatch_step_inputs(g, B, 7, D, C, protos, dt).items() if k in keys}
        if B + rows > K:
            continue
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
        n = 50 if rows * K > 1792 * 16384 else 200
        for _ in range(5):
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
        print(json.dumps({"dtype": dt_name, "rows": rows, "bank": K, "us_per_step": round(us, 2),
                          "samples_per_s": round(rows / us * 1e6), "algorithmic_tflops": round(flop / us * 1e-6, 2)}), flush=True)
        del gs, head
