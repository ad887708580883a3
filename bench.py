#!/usr/bin/env python
"""Benchmark of the SSL head + EMA hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload cfg2|cfg1]

Workload (default ``cfg2`` = BASELINE config 2, the one the metric is quoted on):
CoMatch ResNet-50, embedding dim 64, memory bank K=2560, B=64 labeled + mu=7
unlabeled (B_u=448 rows/step/GPU), bf16 logits / embeddings / bank, fp32 weights.
One step = the CoMatch unlabeled head (DA, memory smoothing, pseudo-label + mask,
enqueue, graph-contrastive and focal soft-CE losses, forward AND backward) followed
by ``ModelEMA.update`` over the ModelwEmb-ResNet-50 state (658 entries / 331 storages /
25.0 M unique elements).  The backbone forward/backward is stock PyTorch and not
part of the metric.

One JSON line is printed by rank 0; see the task contract for the keys.  ``value`` is
measured with inputs resident in HBM, ``e2e`` through the same public API with
host<->device copies from pinned memory and a loss read-back in the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

import torch  # noqa: E402

METRIC = "unlabeled samples/sec through SSL head+EMA"
UNIT = "samples/s"
WORKLOADS = {
    # name: (kind, arch, B, MU, K, dtype)
    "cfg2": dict(kind="comatch", arch="resnet50", B=64, MU=7, K=2560, D=64, C=23, dtype="bf16", thr=0.9,
                 lambda_u=2.0, lambda_c=2.0, decay=0.999,
                 desc="CoMatch ResNet-50, emb dim 64, queue K=2560, B=64 mu=7 bf16 (BASELINE configs[1])"),
    "cfg1": dict(kind="fixmatch", arch="resnet18", B=16, MU=7, K=0, D=0, C=23, dtype="f32", thr=0.95,
                 lambda_u=1.0, lambda_c=0.0, decay=0.999,
                 desc="FixMatch head, ResNet-18 EMA, 23 classes, B=16 mu=7 fp32 (BASELINE configs[0])"),
}


def peaks():
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 6:
                continue
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except ValueError:
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------ reference arm
def cpu_model_name() -> str:
    """Host CPU model for the CPU-baseline report (core count and CPU model beside the number)."""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.lower().startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run_reference(args, wl, rank, world):
    """The reference's own CPU path for the same step (oracle port of code/comatch.py:162-220
    or code/loss.py:126-164, + code/ema.py:51-59), all host threads, rank 0 only."""
    if rank != 0:
        return
    from endoscopy_image_classification_b200 import synthetic as S
    from oracle import ssl_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(0)
    model = S.modelwemb_like(wl["arch"], wl["C"], 64) if wl["kind"] == "comatch" else __import__("torchvision").models.__dict__[wl["arch"]](num_classes=wl["C"])
    from copy import deepcopy
    ema_model = deepcopy(model)
    Bu = wl["B"] * wl["MU"]
    if wl["kind"] == "comatch":
        state = O.CoMatchState.zeros(wl["K"], wl["D"], wl["C"])
        protos = S.rownorm(torch.randn(wl["C"], wl["D"], generator=g))
        batches = [S.comatch_step_inputs(g, wl["B"], wl["MU"], wl["D"], wl["C"], protos) for _ in range(4)]
    else:
        batches = [S.fixmatch_step_inputs(g, wl["B"], wl["MU"], wl["C"]) for _ in range(4)]

    def step(i):
        b = batches[i % len(batches)]
        if wl["kind"] == "comatch":
            c = {k: v for k, v in b.items() if k != "logits_x"}
            o = O.comatch_head(state, **c, thr=wl["thr"], num_classes=wl["C"], enqueue_mode="always")
            val = float(wl["lambda_u"] * o["loss_u"] + wl["lambda_c"] * o["loss_contrast"])
        else:
            val = float(O.fixmatch_head_details(b["logits_u_w"], b["logits_u_s"], wl["thr"])["loss"])
        O.ema_update_(list(ema_model.state_dict().values()), list(model.state_dict().values()), wl["decay"])
        return val

    for i in range(args.warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(i)
    dt = time.perf_counter() - t0
    value = Bu * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "global_batch_unlabeled": Bu,
                       "note": "reference CPU path (oracle port of the reference's PyTorch code; the reference is a Python "
                               "repo and /root/reference does not travel to the GPU box), fp32, rank 0 only"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "cpu_model": cpu_model_name(), "kind": "port",
                             "sample": f"{args.steps} full steps of the workload (head fwd+bwd + EMA) after {args.warmup} warm-up"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ B200 arm
def cpu_baseline_sample(wl, budget_s=10.0):
    from copy import deepcopy

    from endoscopy_image_classification_b200 import synthetic as S
    from oracle import ssl_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(0)
    model = S.modelwemb_like(wl["arch"], wl["C"], 64) if wl["kind"] == "comatch" else __import__("torchvision").models.__dict__[wl["arch"]](num_classes=wl["C"])
    ema_model = deepcopy(model)
    Bu = wl["B"] * wl["MU"]
    if wl["kind"] == "comatch":
        state = O.CoMatchState.zeros(wl["K"], wl["D"], wl["C"])
        protos = S.rownorm(torch.randn(wl["C"], wl["D"], generator=g))
        b = S.comatch_step_inputs(g, wl["B"], wl["MU"], wl["D"], wl["C"], protos)
        b.pop("logits_x")
    else:
        b = S.fixmatch_step_inputs(g, wl["B"], wl["MU"], wl["C"])

    def step():
        if wl["kind"] == "comatch":
            O.comatch_head(state, **b, thr=wl["thr"], num_classes=wl["C"], enqueue_mode="always")
        else:
            O.fixmatch_head_details(b["logits_u_w"], b["logits_u_s"], wl["thr"])
        O.ema_update_(list(ema_model.state_dict().values()), list(model.state_dict().values()), wl["decay"])

    for _ in range(3):
        step()
    n, t0 = 0, time.perf_counter()
    while True:
        step()
        n += 1
        dt = time.perf_counter() - t0
        if dt > budget_s and n >= 5:
            break
    return {"value": Bu * n / dt, "unit": UNIT, "cores": cores, "cpu_model": cpu_model_name(), "kind": "port",
            "sample": f"{n} full steps (oracle head fwd+bwd + EMA loop, fp32) in {dt:.1f} s on {cores} host threads"}


def run_b200(args, wl, rank, world, local_rank):
    import torch.distributed as dist

    from endoscopy_image_classification_b200 import _native as N
    from endoscopy_image_classification_b200 import synthetic as S
    from endoscopy_image_classification_b200.comatch_head import CoMatchHead
    from endoscopy_image_classification_b200.ema import ModelEMA
    from endoscopy_image_classification_b200.loss import consistency_loss

    N.lib()                                   # fail loudly when the extension is missing
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD
    dtype = torch.bfloat16 if wl["dtype"] == "bf16" else torch.float32
    B, MU, C, D = wl["B"], wl["MU"], wl["C"], wl["D"]
    Bu = B * MU
    g = torch.Generator().manual_seed(1234 + rank)

    if wl["kind"] == "comatch":
        model = S.modelwemb_like(wl["arch"], C, D).to(dev)
        # comatch.py:91: queue_size = queue_batch*(MU+1)*BATCH_SIZE.  Data parallel, BATCH_SIZE is the global batch, so the
        # global bank grows with the rank count and every rank keeps a shard of the single-GPU size (weak scaling).
        K_global = wl["K"] * world
        head = CoMatchHead(C, D, K_global, wl["thr"], enqueue_mode="always", device=dev, dtype=dtype, process_group=pg,
                           exchange=args.exchange)
        protos = S.rownorm(torch.randn(C, D, generator=torch.Generator().manual_seed(99)))
        host = [S.comatch_step_inputs(g, B, MU, D, C, protos, dtype) for _ in range(4)]
        keys = ["logits_u_w", "logits_u_s0", "feats_u_w", "feats_u_s0", "feats_u_s1", "feats_x", "targets_x"]
    else:
        import torchvision
        model = torchvision.models.__dict__[wl["arch"]](num_classes=C).to(dev)
        head = None
        host = [S.fixmatch_step_inputs(g, B, MU, C, dtype=dtype) for _ in range(4)]
        keys = ["logits_u_w", "logits_u_s"]
    host = [{k: b[k].pin_memory() for k in keys} for b in host]
    resident = [{k: v.to(dev) for k, v in b.items()} for b in host]
    ema = ModelEMA(model, decay=wl["decay"], device=dev)
    gdev = torch.Generator(device=dev).manual_seed(5)
    S.perturb_(model, gdev)                    # m != e, like after an optimizer step
    grad_keys = ("logits_u_s0", "feats_u_s0", "feats_u_s1") if wl["kind"] == "comatch" else ("logits_u_s",)
    # N=1: smooth, rows (DA+finalize+enqueue), contrast fwd, contrast bwd (+scale), ema.  N>1: + the side-stream enqueue
    # with a peer-memory resident bank; + enqueue and three exchange launches with the peer-memory exchanges; + enqueue
    # with NCCL (its kernels are not counted).
    launches_per_step = (5 if wl["kind"] == "comatch" else 3)
    if world > 1 and head is not None:
        launches_per_step += {"replicated": 1, "direct": 1, "peer": 4, "collective": 1}[head.exchange]
    one = torch.ones((), dtype=torch.float32, device=dev)

    def step(batch):
        for k in grad_keys:
            batch[k].grad = None
            batch[k].requires_grad_(True)
        if wl["kind"] == "comatch":
            total = head.total_loss(*[batch[k] for k in keys], lambda_u=wl["lambda_u"],
                                    lambda_c=wl["lambda_c"])[0]                # comatch.py:222 (unlabeled part)
        else:
            lu, _ = consistency_loss(batch["logits_u_w"], batch["logits_u_s"], T=1.0, p_cutoff=wl["thr"])
            total = wl["lambda_u"] * lu                                        # fixmatch.py:118
        total.backward(gradient=one)                                          # cached ones scalar: no fill kernel per step
        ema.update(model)
        return total

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    from endoscopy_image_classification_b200.graphs import GraphedStep

    # ---------------- eager (un-graphed) reference numbers: same public API, Python launch bound
    for i in range(args.warmup):
        step(resident[i % len(resident)])
    barrier()
    n_eager = min(args.steps, 200)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(n_eager):
        step(resident[i % len(resident)])
    t1.record()
    barrier()
    eager_ms = t0.elapsed_time(t1) / n_eager

    # ---------------- CUDA-graph capture of the whole step (head fwd + bwd + EMA)
    n_rows = B + Bu
    graphed = GraphedStep(lambda b: step(b), resident[0], dev, warmup=3,
                          on_replay=(lambda: head.note_graph_replay(n_rows)) if head is not None else None,
                          after_capture=(lambda: head.sync_ptr_from_device()) if head is not None else None)

    # ---------------- device-resident arm -------------------------------------------------
    for i in range(args.warmup):
        graphed.replay()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        time.sleep(1.0)                        # let nvidia-smi start before the timed regions
    barrier()
    t_start, t_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.profiler.start()                # ncu --profile-from-start off: only the timed graph replays
    t_start.record()
    for i in range(args.steps):
        graphed.replay()
    t_stop.record()
    barrier()
    torch.cuda.profiler.stop()
    ms = t_start.elapsed_time(t_stop)

    # dominant kernel timed on its own stream position: back-to-back EMA launches (300 MB each > L2)
    for _ in range(5):
        ema.update(model)
    barrier()
    n_ema = max(50, min(args.steps, 500))
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    for _ in range(n_ema):
        ema.update(model)
    a1.record()
    barrier()
    ema_ms = a0.elapsed_time(a1) / n_ema

    # ---------------- end-to-end arm: pinned host inputs, H2D in the timed region, loss read back
    h2d = graphed.h2d_bytes
    plain = [{k: v.clone() for k, v in hb.items()} for hb in host]        # ordinary (pageable) host tensors
    for i in range(max(3, args.warmup // 2)):
        graphed.replay_host(plain[i % len(plain)])
    barrier()
    e_start, e_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_start.record()
    for i in range(args.steps):
        graphed.replay_host(plain[i % len(plain)])     # stage -> H2D -> head fwd+bwd -> EMA -> D2H loss -> sync
    e_stop.record()
    barrier()
    e2e_ms = e_start.elapsed_time(e_stop)

    # eager end-to-end (no graph), for the record
    def e2e_eager(i):
        hb = host[i % len(host)]
        batch = {k: hb[k].to(dev, non_blocking=True) for k in keys}
        total = step(batch)
        return float(total.detach())           # the trainer's losses.item() (comatch.py:234)

    for i in range(5):
        e2e_eager(i)
    barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for i in range(n_eager):
        e2e_eager(i)
    g1.record()
    barrier()
    e2e_eager_ms = g0.elapsed_time(g1) / n_eager
    clocks = sampler.stop() if sampler else None

    if world > 1 and head is not None and head.peer_timeouts():
        raise RuntimeError(f"rank {rank}: peer-memory waits timed out; no valid measurement")
    times = torch.tensor([ms, e2e_ms, ema_ms, eager_ms, e2e_eager_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms, e2e_ms, ema_ms, eager_ms, e2e_eager_ms = (float(x) for x in times.tolist())
    if rank == 0:
        peak, peak_src = peaks()
        plan = ema.plan
        achieved = plan.bytes_per_update / (ema_ms * 1e-3) / 1e9
        cpu = cpu_baseline_sample(wl) if world == 1 and not args.no_cpu_baseline else None
        line = {"metric": METRIC, "value": world * Bu * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": wl["dtype"], "data": "synthetic",
                "config": {"workload": wl["desc"], "global_batch_unlabeled": world * Bu, "per_gpu_unlabeled": Bu,
                           "bank_rows_global": wl["K"] * world, "bank_rows_per_rank": wl["K"], "bank_sharded_over": world, "parallelism": f"dp{world}",
                           "bank_exchange": (None if world == 1 or head is None else
                                             {"replicated": "every rank keeps the whole ring in NVLink peer memory; the enqueue block is written "
                                                            "through into every copy by the row kernel (the only exchange of a step); "
                                                            "two epoch flags per step, one extra (side-stream) enqueue launch, no collective",
                                              "direct": "shards in NVLink peer memory: K3 reads every shard in place (TMA over NVLink), "
                                                        "the enqueue stores into the owning shard; two epoch flags per step, one extra "
                                                        "(side-stream) enqueue launch, no collective",
                                              "peer": "own kernels over NVLink peer memory (csrc/peer.cu): all-gather, reduce-scatter, "
                                                      "all-gather per step",
                                              "collective": "NCCL collectives"}[head.exchange]),
                           "ema_state": {"entries": plan.n_entries, "unique_storages": plan.n_unique,
                                         "unique_elems": plan.unique_elems, "blocks": plan.n_blocks},
                           "execution": "whole step (head fwd+bwd + EMA) captured once in a CUDA graph and replayed; "
                                        "eager_ms_per_step / e2e_eager_ms_per_step are the same API without the graph",
                           "eager_ms_per_step": eager_ms, "e2e_eager_ms_per_step": e2e_eager_ms,
                           "l2": "no explicit flush: the EMA kernel streams 300 MB/step (> 126 MB L2); head inputs (~0.3 MB) "
                                 "come straight from the backbone in training, i.e. L2-resident there too"},
                "roofline": {"kernel": "ema_multi_tensor_kernel", "bound": "hbm", "achieved": achieved, "peak": peak,
                             "unit": "GB/s", "frac": achieved / peak,
                             # dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full
                             # (profiles/r01_ncu_ema_full_raw.csv; ModelwEmb-R50 state only)
                             "traffic": 252.96e6 if plan.bytes_per_update == 300359460 else None, "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": plan.bytes_per_update, "avg_launch_ms": ema_ms,
                             "timed": f"{n_ema} back-to-back launches between two CUDA events on the launching stream"},
                "e2e": {"value": world * Bu * args.steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms / args.steps},
                "gpu_launches": launches_per_step * args.steps,
                "clocks": clocks}
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        # captured graphs hold NCCL work: they must be released before the communicator is torn down
        # (destroy_process_group() blocks forever otherwise, seen on 2xB200 with NCCL 2.28.9)
        import gc
        graphed.graph = graphed.graph_host = None
        del graphed
        gc.collect()
        torch.cuda.synchronize(dev)
        if head is not None:
            head.close()
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="auto", choices=["auto", "replicated", "direct", "peer", "collective"],
                    help="sharded bank at N>1: directly addressed shards in NVLink peer memory (auto), own peer-memory "
                         "exchange kernels, or NCCL collectives")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl, rank, world)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", __file__] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_b200(args, wl, rank, world, local_rank)


if __name__ == "__main__":
    main()
