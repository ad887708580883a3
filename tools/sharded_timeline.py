#!/usr/bin/env python
"""In-kernel timeline of one sharded CoMatch step (directly addressed bank) on every rank:
clock64() stamps of the smoothing, row, contrastive forward and backward kernels, printed by rank 0
as microseconds since each CTA's own first stamp (clock64 is per SM).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/sharded_timeline.py [exchange]
"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from endoscopy_image_classification_b200 import _native as N  # noqa: E402
from endoscopy_image_classification_b200 import synthetic as S  # noqa: E402
from endoscopy_image_classification_b200.comatch_head import CoMatchHead  # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
torch.cuda.set_device(dev)
pg = None
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    pg = dist.group.WORLD
exchange = sys.argv[1] if len(sys.argv) > 1 else "auto"
B, MU, C, D, K = 64, 7, 23, 64, int(os.environ.get("K_GLOBAL", 2560 * world))      # K_GLOBAL=65536: BASELINE cfg 4
keys = ["logits_u_w", "logits_u_s0", "feats_u_w", "feats_u_s0", "feats_u_s1", "feats_x", "targets_x"]
g = torch.Generator().manual_seed(1 + rank)
protos = S.rownorm(torch.randn(C, D, generator=torch.Generator().manual_seed(99)))
batch = {k: v.to(dev) for k, v in S.comatch_step_inputs(g, B, MU, D, C, protos, torch.bfloat16).items() if k in keys}
head = CoMatchHead(C, D, K, 0.9, enqueue_mode="always", device=dev, dtype=torch.bfloat16, process_group=pg, exchange=exchange)
one = torch.ones((), device=dev)


def step():
    for k in ("logits_u_s0", "feats_u_s0", "feats_u_s1"):
        batch[k].grad = None
        batch[k].requires_grad_(True)
    total = head.total_loss(*[batch[k] for k in keys], lambda_u=1.0, lambda_c=1.0)[0]
    total.backward(gradient=one)


for _ in range(10):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
REGION = 4096 * 16
buf = torch.zeros(6 * REGION, dtype=torch.int64, device=dev)
N.lib().b200ssl_debug_set_timing_buffer(buf.data_ptr())
step()
torch.cuda.synchronize()
N.lib().b200ssl_debug_set_timing_buffer(None)
t = buf.cpu().numpy().reshape(6, -1, 16)
if rank == 0:
    # clock64() is per SM: only differences inside one CTA are meaningful
    print(f"exchange={head.exchange} world={world}; us since the CTA's first stamp (min / median / max over CTAs), 1.965 GHz")
    for tag, name in enumerate(["smooth", "rows", "contrast fwd", "contrast bwd"]):
        r = t[tag]
        r = r[(r != 0).any(axis=1)]
        if not len(r):
            continue
        clk = r[:, :10]                                        # slots 0..9: clock64; 10..12 (smooth): %globaltimer ns
        first = np.where(clk != 0, clk, np.iinfo(np.int64).max).min(axis=1)
        print(f"  {name} ({len(r)} CTAs)")
        if tag == 0 and (r[:, 10] != 0).any():
            end = np.maximum(r[:, 11], r[:, 12])
            print(f"    wall clock: CTA starts spread {(r[:, 10].max() - r[:, 10].min()) / 1e3:.2f} us, first start -> last CTA done "
                  f"{(end.max() - r[:, 10].min()) / 1e3:.2f} us")
        for slot in range(10):
            ok = r[:, slot] != 0
            if ok.any():
                d = (r[ok, slot] - first[ok]) / 1965.0
                print(f"    stamp {slot:2d}: {d.min():7.2f} {np.median(d):7.2f} {d.max():7.2f}   (n={int(ok.sum())})")
if world > 1:
    print(f"rank {rank}: peer timeouts {head.peer_timeouts()}")
    head.close()
    dist.barrier()
    dist.destroy_process_group()
