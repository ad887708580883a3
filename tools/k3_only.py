#!/usr/bin/env python
"""A handful of launches of the tensor-core K3 alone (for `ncu -k regex:bank_smooth`).

    python tools/k3_only.py ROWS BANK [MT CLUSTER NOUTER [POLY]]     (0 = planner)
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from endoscopy_image_classification_b200 import _native as N  # noqa: E402
from endoscopy_image_classification_b200 import synthetic as S  # noqa: E402
from endoscopy_image_classification_b200.comatch_head import CoMatchHead  # noqa: E402

rows, K = int(sys.argv[1]), int(sys.argv[2])
N.lib().b200ssl_debug_set_k3(*(int(x) for x in (sys.argv[3:6] + ['0', '0', '0'])[:3]), int(sys.argv[6]) if len(sys.argv) > 6 else -1)
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
head = CoMatchHead(23, 64, K, 0.9, enqueue_mode="always", device=dev, dtype=torch.bfloat16)
head.queue_feats.copy_(S.rownorm(torch.randn(K, 64, generator=g)).to(torch.bfloat16))
qp = torch.softmax(torch.randn(K, 23, generator=g), 1).to(torch.bfloat16)
head.queue_probs.copy_(qp)
head.queue_probs_t[:23].copy_(qp.t())
fw = S.rownorm(torch.randn(rows, 64, generator=g)).to(torch.bfloat16).to(dev)
for _ in range(5):
    rowsum, numer = head._k_smooth(fw)
torch.cuda.synchronize()
print("ok", float(rowsum.sum()))
