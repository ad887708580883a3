#!/usr/bin/env python
"""Quick A/B timer: graph-replayed CoMatch step (head fwd+bwd + EMA, BASELINE configs[1]) in us/step.
Used with B200SSL_PDL_MASK (bit per kernel, see csrc/common.cuh) to see which launches gain from
programmatic dependent launch.   python tools/step_time.py [replays]"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from endoscopy_image_classification_b200 import synthetic as S  # noqa: E402
from endoscopy_image_classification_b200.comatch_head import CoMatchHead  # noqa: E402
from endoscopy_image_classification_b200.ema import ModelEMA  # noqa: E402
from endoscopy_image_classification_b200.graphs import GraphedStep  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 500
with_ema = os.environ.get("STEP_NO_EMA") is None
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
torch.cuda.set_device(dev)
pg = None
if world > 1:                      # under torchrun: the sharded bank (STEP_EXCHANGE = auto | direct | peer | collective)
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
    pg = dist.group.WORLD
B, MU, C, D, K = 64, 7, 23, 64, 2560
g = torch.Generator().manual_seed(1 + rank)
keys = ["logits_u_w", "logits_u_s0", "feats_u_w", "feats_u_s0", "feats_u_s1", "feats_x", "targets_x"]
protos = S.rownorm(torch.randn(C, D, generator=torch.Generator().manual_seed(99)))
batch = {k: v.to(dev) for k, v in S.comatch_step_inputs(g, B, MU, D, C, protos, torch.bfloat16).items() if k in keys}
head = CoMatchHead(C, D, K * world, 0.9, enqueue_mode="always", device=dev, dtype=torch.bfloat16, process_group=pg,
                   exchange=os.environ.get("STEP_EXCHANGE", "auto"))
model = S.modelwemb_like("resnet50", C, D).to(dev)
ema = ModelEMA(model, 0.999, device=dev)
one = torch.ones((), device=dev)


def step(b):
    for k in ("logits_u_s0", "feats_u_s0", "feats_u_s1"):
        b[k].grad = None
        b[k].requires_grad_(True)
    total = head.total_loss(*[b[k] for k in keys], lambda_u=1.0, lambda_c=1.0)[0]
    total.backward(gradient=one)
    if with_ema:
        ema.update(model)
    return total


for _ in range(5):
    step(batch)
gs = GraphedStep(step, batch, dev, warmup=3, on_replay=lambda: head.note_graph_replay(B + B * MU), after_capture=head.sync_ptr_from_device)
for _ in range(20):
    gs.replay()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        gs.replay()
    e1.record()
    torch.cuda.synchronize()
    best = min(best, 1e3 * e0.elapsed_time(e1) / n)
if world > 1:
    tb = torch.tensor([best], device=dev, dtype=torch.float64)
    dist.all_reduce(tb, op=dist.ReduceOp.MAX)
    best = float(tb)
if rank == 0:
    print(f"world={world} exchange={head.exchange if world > 1 else '-'} mask={os.environ.get('B200SSL_PDL_MASK', 'default')} "
          f"ema={with_ema}: {best:.2f} us/step")
if world > 1:
    assert head.peer_timeouts() == 0
    gs.graph = gs.graph_host = None
    del gs
    import gc
    gc.collect()
    torch.cuda.synchronize()
    head.close()
    dist.barrier()
    dist.destroy_process_group()
