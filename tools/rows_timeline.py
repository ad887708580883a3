#!/usr/bin/env python
"""In-kernel timeline of the fused rows kernel (clock64 stamps per CTA, see B200SSL_STAMP)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch  # noqa
from endoscopy_image_classification_b200 import _native as N, synthetic as S  # noqa
from endoscopy_image_classification_b200.comatch_head import CoMatchHead  # noqa
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
dt = torch.bfloat16
b = {k: v.to(dev) for k, v in S.comatch_step_inputs(g, 64, 7, 64, 23, dtype=dt).items()}
head = CoMatchHead(23, 64, 2560, 0.9, enqueue_mode="always", device=dev, dtype=dt)
lw, ls0, fw, fx, tx = b["logits_u_w"], b["logits_u_s0"], b["feats_u_w"], b["feats_x"], b["targets_x"]
rowsum, numer = head._k_smooth(fw)
for _ in range(3):
    head._k_rows_fused(lw, ls0, rowsum, numer, (0, 0), fw, fx, tx, True, False)
torch.cuda.synchronize()
buf = torch.zeros(4 * 4096 * 16, dtype=torch.int64, device=dev)   # one region per instrumented kernel
N.lib().b200ssl_debug_set_timing_buffer(buf.data_ptr())
head._k_rows_fused(lw, ls0, rowsum, numer, (0, 0), fw, fx, tx, True, False)
torch.cuda.synchronize()
N.lib().b200ssl_debug_set_timing_buffer(None)
t = buf.cpu().numpy().reshape(-1, 16); t = t[t[:, 0] != 0]
names = ["start", "phase1 colsums", "cluster sync 1", "prob_avg ready", "phase2 tiles loaded", "rows finalised", "stores+enqueue issued", "end"]
print(f"{len(t)} CTAs; cycles since CTA start (min / median / max)")
for i in range(1, 8):
    col = t[:, i] - t[:, 0]
    print(f"  {names[i]:24s} {col.min():8d} {int(np.median(col)):8d} {col.max():8d}")
