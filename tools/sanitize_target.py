#!/usr/bin/env python
"""The launches whose correctness rests on intra-kernel synchronisation -- thread-block clusters + distributed shared memory
(fused row kernel, K3 / K6 folds), mbarrier pipelines around TMA / tcgen05 (K3, K6), epoch flags in peer memory (sharded
bank, two emulated ranks) -- once each at BASELINE cfg 2's shapes, for `compute-sanitizer --tool racecheck|synccheck`:

    compute-sanitizer --tool racecheck python tools/sanitize_target.py

Parity of the results is checked elsewhere (tests/); this script only has to exercise the code paths."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from endoscopy_image_classification_b200 import synthetic as S  # noqa: E402
from endoscopy_image_classification_b200.comatch_head import CoMatchHead, lockstep_total_loss  # noqa: E402
from endoscopy_image_classification_b200.ema import ModelEMA  # noqa: E402
from endoscopy_image_classification_b200.loss import ce_loss, consistency_loss  # noqa: E402
from endoscopy_image_classification_b200.peer import LocalArenaSet  # noqa: E402

dev = torch.device("cuda:0")
C, D, B, MU = 23, 64, 64, 7
g = torch.Generator().manual_seed(0)
protos = S.rownorm(torch.randn(C, D, generator=g))


def batch(dtype, b=B):
    d = {k: v.to(dev) for k, v in S.comatch_step_inputs(g, b, MU, D, C, protos, dtype).items()}
    lx = d.pop("logits_x")
    for k in ("logits_u_s0", "feats_u_s0", "feats_u_s1"):
        d[k].requires_grad_(True)
    return d, lx


for dtype in (torch.bfloat16, torch.float32):                     # tcgen05 kernels / exact-fp32 kernels
    head = CoMatchHead(C, D, 2560, 0.9, enqueue_mode="always", device=dev, dtype=dtype)
    for _ in range(2):
        d, lx = batch(dtype)
        total = head.total_loss(**d, lambda_u=2.0, lambda_c=2.0)[0]
        total.backward()
    torch.cuda.synchronize()
    print("head", dtype, float(total))

# K3 with an outer (ticketed) fold and with the row loop
head = CoMatchHead(C, D, 20480, 0.9, enqueue_mode="always", device=dev, dtype=torch.bfloat16)
d, _ = batch(torch.bfloat16)
print("K3 outer fold", float(head._k_smooth(d["feats_u_w"].detach())[0].sum()))

# two emulated ranks of a sharded bank (peer-memory flags, remote-style tensor maps, side-stream enqueue)
ranks = LocalArenaSet(2, dev)
heads = [CoMatchHead(C, D, 3 * 2 * 128, 0.9, enqueue_mode="always", device=dev, dtype=torch.bfloat16, exchange="direct",
                     local_ranks=(ranks, r)) for r in range(2)]
for _ in range(2):
    outs = lockstep_total_loss(heads, [batch(torch.bfloat16, 16)[0] for _ in range(2)])
    for o in outs:
        o[0].backward()
torch.cuda.synchronize()
print("emulated ranks", [float(o[0]) for o in outs], [h.peer_timeouts() for h in heads])
for h in heads:
    h.close()
ranks.close()

# FixMatch head, labeled CE, EMA
w, s = torch.randn(448, C, device=dev), torch.randn(448, C, device=dev, requires_grad=True)
lu, mm = consistency_loss(w, s, p_cutoff=0.5)
(lu + ce_loss(s[:64], torch.randint(0, C, (64,), device=dev), reduction="mean", type_loss="poly")).backward()
net = torch.nn.Sequential(torch.nn.Conv2d(3, 16, 3), torch.nn.BatchNorm2d(16)).to(dev)
ema = ModelEMA(net, 0.999, device=dev)
ema.update(net)
torch.cuda.synchronize()
print("sanitize target done")
