#!/usr/bin/env python
"""A few launches of the tensor-core contrastive kernels alone (for `ncu -k regex:contrast_tc`).

    python tools/contrast_only.py ROWS
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from endoscopy_image_classification_b200 import synthetic as S  # noqa: E402
from endoscopy_image_classification_b200.comatch_head import CoMatchHead  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 448
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
dt = torch.bfloat16
nf = lambda: S.rownorm(torch.randn(rows, 64, generator=g)).to(dt).to(dev)
f0, f1 = nf(), nf()
y = torch.randint(0, 23, (rows,), generator=g)
probs = torch.softmax(6.0 * torch.nn.functional.one_hot(y, 23).float() + torch.randn(rows, 23, generator=g), 1)
hi = probs.to(dt)
lo = (probs - hi.float()).to(dt)
hl = torch.zeros(rows, 64, dtype=dt)
hl[:, :23], hl[:, 32:55] = hi, lo
head = CoMatchHead(23, 64, 64, 0.9, dtype=dt, device=dev)
scal = torch.zeros(4, device=dev)
dp, dhl, one = probs.to(dev), hl.to(dev), torch.ones(1, device=dev)
for _ in range(4):
    stats, _ = head._k_contrast_fwd(f0, f1, dp, scal, probs_hl=dhl)
    g0, g1 = head._k_contrast_bwd(f0, f1, dp, stats, one, probs_hl=dhl)
torch.cuda.synchronize()
print("ok", float(scal[2]), float(g0.float().abs().sum()))
