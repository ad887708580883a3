"""Probe: can torch.distributed NCCL collectives be captured into a CUDA graph here?
torchrun --nproc-per-node 2 tools/nccl_graph_probe.py <capture_error_mode>"""
import os
import sys
import time

import torch
import torch.distributed as dist

mode = sys.argv[1] if len(sys.argv) > 1 else "global"
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.cuda.set_device(dev)
print(f"rank {rank}: init...", flush=True)
dist.init_process_group("nccl", device_id=dev)
print(f"rank {rank}: init done", flush=True)
x = torch.full((512, 64), float(rank + 1), device=dev)
g_out = torch.empty(world * 512, 64, device=dev)
p = torch.ones(world * 448, 24, device=dev)
rs = torch.empty(448, 24, device=dev)


def step():
    dist.all_gather_into_tensor(g_out, x)
    dist.reduce_scatter_tensor(rs, p, op=dist.ReduceOp.SUM)
    return g_out.sum() + rs.sum()


for _ in range(3):
    step()
torch.cuda.synchronize()
print(f"rank {rank}: eager collectives ok", flush=True)
if mode == "eager":
    dist.destroy_process_group()
    sys.exit(0)
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        step()
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
print(f"rank {rank}: side-stream warmup ok", flush=True)
t0 = time.time()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, capture_error_mode=mode):
    out = step()
torch.cuda.synchronize()
print(f"rank {rank}: captured ({mode}) in {time.time()-t0:.2f}s", flush=True)
for _ in range(20):
    g.replay()
torch.cuda.synchronize()
print(f"rank {rank}: replay ok value {float(out):.1f}", flush=True)
t = torch.ones(1, device=dev)
dist.all_reduce(t)
torch.cuda.synchronize()
print(f"rank {rank}: eager all_reduce after replay ok {float(t)}", flush=True)
for _ in range(5):
    g.replay()
torch.cuda.synchronize()
print(f"rank {rank}: second replay batch ok", flush=True)
dist.barrier()
torch.cuda.synchronize()
print(f"rank {rank}: barrier ok", flush=True)
del g
dist.destroy_process_group()
print(f"rank {rank}: destroyed", flush=True)
