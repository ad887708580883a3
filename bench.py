#!/usr/bin/env python
"""Benchmark of the SSL head + EMA hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload cfg2|cfg1|cfg3|cfg4]

Workload (default ``cfg2`` = BASELINE config 2, the one the metric is quoted on):
CoMatch ResNet-50, embedding dim 64, memory bank K=2560, B=64 labeled + mu=7
unlabeled (B_u=448 rows/step/GPU), bf16 logits / embeddings / bank, fp32 weights.
One step = the CoMatch unlabeled head (DA, memory smoothing, pseudo-label + mask,
enqueue, graph-contrastive and focal soft-CE losses, forward AND backward) followed
by ``ModelEMA.update`` over the ModelwEmb-ResNet-50 state (658 entries / 331 storages /
25.0 M unique elements).  The backbone forward/backward is stock PyTorch and not
part of the metric.

Timing: the K timed steps are replays of one CUDA graph, bracketed by a barrier +
``torch.cuda.synchronize()`` on both sides, with the W warm-up replays queued right
after the opening barrier (so the ranks' streams are in lock step when the first
event is recorded -- the ranks of a multi-GPU run are coupled through the bank's
epoch flags and a 2 ms window does not survive a late rank).  That bracket is
repeated ``blocks`` times; each block is timed by its own pair of CUDA events, the
MAX over ranks is taken per block, and the line reports the MEDIAN block (min / max
in ``config.timing``).

One JSON line is printed by rank 0; see the task contract for the keys.  ``value`` is
measured with inputs resident in HBM, ``e2e`` through the same public API with
host<->device copies from pinned memory and a loss read-back in the timed region.
Extra keyed blocks in the same line: ``cfg4`` (BASELINE configs[3]: the 65536-row bank,
sharded over the ranks when N > 1), ``fp32`` (configs[1] with the reference's own fp32
storage), ``fused_opt_ema`` and, for N > 1, ``parity`` (every rank's
step against the single-process oracle before anything is timed).
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
    "cfg2": dict(kind="comatch", arch="resnet50", B=64, MU=7, K=2560, D=64, C=23, dtype="bf16", thr=0.9,
                 lambda_u=2.0, lambda_c=2.0, decay=0.999, k_scales_with_ranks=True,
                 desc="CoMatch ResNet-50, emb dim 64, queue K=2560, B=64 mu=7 bf16 (BASELINE configs[1])"),
    "cfg4": dict(kind="comatch", arch="resnet50", B=64, MU=7, K=65536, D=64, C=23, dtype="bf16", thr=0.9,
                 lambda_u=2.0, lambda_c=2.0, decay=0.999, k_scales_with_ranks=False,
                 desc="CoMatch ResNet-50, emb dim 64, 65536-entry memory bank (sharded over the ranks), B=64 mu=7 per GPU, "
                      "bf16 (BASELINE configs[3])"),
    "cfg1": dict(kind="fixmatch", arch="resnet18", B=16, MU=7, K=0, D=0, C=23, dtype="f32", thr=0.95,
                 lambda_u=1.0, lambda_c=0.0, decay=0.999, k_scales_with_ranks=False,
                 desc="FixMatch head, ResNet-18 EMA, 23 classes, B=16 mu=7 fp32 (BASELINE configs[0])"),
    "cfg3": dict(kind="semiformer", arch="vit_s_16", B=64, MU=7, K=0, D=0, C=23, dtype="bf16", thr=0.95,
                 lambda_u=1.0, lambda_c=0.0, decay=0.999, k_scales_with_ranks=False,
                 desc="SemiFormer dual-head FixMatch (threshold 0.95), ViT-S/16-sized two-head state, EMA 0.999, B=64 mu=7 per GPU, "
                      "bf16 logits (BASELINE configs[2])"),
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
    """nvidia-smi clocks / throttle reasons sampled during the timed regions (every rank samples its own GPU, so the
    host-side load is the same on every rank)."""
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


# ------------------------------------------------------------------ models / inputs shared by both arms
def vit_s16_two_head(C):
    """ViT-S/16-sized state with the two classifier heads of the reference's Conformer contract (semiformer.py:122-131:
    ``model(x) -> (out_conv, out_trans)``): 12 layers, width 384, 6 heads, MLP 1536 (22 M parameters)."""
    import torchvision
    vit = torchvision.models.vision_transformer.VisionTransformer(image_size=224, patch_size=16, num_layers=12, num_heads=6,
                                                                  hidden_dim=384, mlp_dim=1536, num_classes=C)
    vit.conv_cls_head = torch.nn.Linear(384, C)
    return vit


def build_model(wl):
    from endoscopy_image_classification_b200 import synthetic as S
    if wl["kind"] == "comatch":
        return S.modelwemb_like(wl["arch"], wl["C"], wl["D"])
    if wl["kind"] == "semiformer":
        return vit_s16_two_head(wl["C"])
    import torchvision
    return torchvision.models.__dict__[wl["arch"]](num_classes=wl["C"])


def make_batches(wl, g, n, dtype=torch.float32):
    from endoscopy_image_classification_b200 import synthetic as S
    if wl["kind"] == "comatch":
        protos = S.rownorm(torch.randn(wl["C"], wl["D"], generator=torch.Generator().manual_seed(99)))
        out = [S.comatch_step_inputs(g, wl["B"], wl["MU"], wl["D"], wl["C"], protos, dtype) for _ in range(n)]
        keys = ["logits_u_w", "logits_u_s0", "feats_u_w", "feats_u_s0", "feats_u_s1", "feats_x", "targets_x"]
    elif wl["kind"] == "semiformer":
        out = []
        for _ in range(n):
            b = S.fixmatch_step_inputs(g, wl["B"], wl["MU"], wl["C"], dtype=dtype)
            b["logits_u_s_trans"] = S.fixmatch_step_inputs(g, wl["B"], wl["MU"], wl["C"], dtype=dtype)["logits_u_s"]
            out.append(b)
        keys = ["logits_u_w", "logits_u_s", "logits_u_s_trans"]
    else:
        out = [S.fixmatch_step_inputs(g, wl["B"], wl["MU"], wl["C"], dtype=dtype) for _ in range(n)]
        keys = ["logits_u_w", "logits_u_s"]
    return [{k: b[k] for k in keys} for b in out], keys


# ------------------------------------------------------------------ reference arm / CPU baseline
class CpuPath:
    """The reference's own CPU path for one step of the workload (oracle port of code/comatch.py:162-220 or
    code/loss.py:126-164, + code/ema.py:51-59), fp32, all host threads."""

    def __init__(self, wl):
        from copy import deepcopy

        from oracle import ssl_oracle as O
        self.O, self.wl = O, wl
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        g = torch.Generator().manual_seed(0)
        self.model = build_model(wl)
        self.ema_model = deepcopy(self.model)
        self.batches, _ = make_batches(wl, g, 4)
        self.state = O.CoMatchState.zeros(wl["K"], wl["D"], wl["C"]) if wl["kind"] == "comatch" else None

    def step(self, i):
        O, wl, b = self.O, self.wl, self.batches[i % len(self.batches)]
        if wl["kind"] == "comatch":
            o = O.comatch_head(self.state, **b, thr=wl["thr"], num_classes=wl["C"], enqueue_mode="always")
            val = float(wl["lambda_u"] * o["loss_u"] + wl["lambda_c"] * o["loss_contrast"])
        elif wl["kind"] == "semiformer":
            a = O.fixmatch_head_details(b["logits_u_w"], b["logits_u_s"], wl["thr"])["loss"]
            c = O.fixmatch_head_details(b["logits_u_w"], b["logits_u_s_trans"], wl["thr"])["loss"]
            val = float(a + c)
        else:
            val = float(O.fixmatch_head_details(b["logits_u_w"], b["logits_u_s"], wl["thr"])["loss"])
        O.ema_update_(list(self.ema_model.state_dict().values()), list(self.model.state_dict().values()), wl["decay"])
        return val


def run_reference(args, wl, rank, world):
    if rank != 0:
        return
    cpu = CpuPath(wl)
    Bu = wl["B"] * wl["MU"]
    for i in range(args.warmup):
        cpu.step(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        cpu.step(i)
    dt = time.perf_counter() - t0
    value = Bu * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "global_batch_unlabeled": Bu,
                       "note": "reference CPU path (oracle port of the reference's PyTorch code; the reference is a Python "
                               "repo and /root/reference does not travel to the GPU box), fp32 (the reference's own arithmetic; "
                               "the B200 arm stores logits / embeddings / bank in bf16 as BASELINE configs[1] asks), rank 0 only"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cpu.cores, "cpu_model": cpu_model_name(), "kind": "port",
                             "sample": f"{args.steps} full steps of the workload (head fwd+bwd + EMA) after {args.warmup} warm-up"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline_sample(wl, budget_s=10.0):
    cpu = CpuPath(wl)
    for i in range(3):
        cpu.step(i)
    n, t0 = 0, time.perf_counter()
    while True:
        cpu.step(n)
        n += 1
        dt = time.perf_counter() - t0
        if dt > budget_s and n >= 5:
            break
    return {"value": wl["B"] * wl["MU"] * n / dt, "unit": UNIT, "cores": cpu.cores, "cpu_model": cpu_model_name(), "kind": "port",
            "sample": f"{n} full steps (oracle head fwd+bwd + EMA loop, fp32) in {dt:.1f} s on {cpu.cores} host threads"}


# ------------------------------------------------------------------ B200 arm
class Ctx:
    def __init__(self, rank, world, local_rank):
        import torch.distributed as dist
        self.rank, self.world, self.dist = rank, world, dist
        self.dev = torch.device("cuda", local_rank)
        torch.cuda.set_device(self.dev)
        self.pg = None
        if world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
            self.pg = dist.group.WORLD

    def barrier(self):
        torch.cuda.synchronize(self.dev)
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, values):
        t = torch.tensor(list(values), dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]


def timed_blocks(ctx, run_once, steps, warmup, blocks):
    """``blocks`` x [barrier+sync | W untimed | event | K timed | event] ... barrier+sync; per-block ms, MAX over ranks."""
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(blocks)]
    i = 0
    for a, b in evs:
        ctx.barrier()
        for _ in range(warmup):
            run_once(i)
            i += 1
        a.record()
        for _ in range(steps):
            run_once(i)
            i += 1
        b.record()
    ctx.barrier()
    return ctx.max_over_ranks(a.elapsed_time(b) for a, b in evs)


def block_stats(ms_blocks, steps):
    per = sorted(x / steps for x in ms_blocks)
    return {"blocks": len(per), "median_ms_per_step": statistics.median(per), "min_ms_per_step": per[0], "max_ms_per_step": per[-1],
            "blocks_slower_than_1p25_median": sum(1 for x in per if x > 1.25 * statistics.median(per))}


class Case:
    """One workload on this rank: head (+ bank), model / EMA state, synthetic batches, the step, its CUDA graph."""

    def __init__(self, ctx, wl, exchange="auto", dtype_name=None, seed=1234, ema_overlap=True):
        from endoscopy_image_classification_b200 import synthetic as S
        from endoscopy_image_classification_b200.comatch_head import CoMatchHead
        from endoscopy_image_classification_b200.ema import ModelEMA
        from endoscopy_image_classification_b200.graphs import GraphedStep
        from endoscopy_image_classification_b200.loss import consistency_loss, consistency_loss_dual
        self.ctx, self.wl = ctx, wl
        dev, world, rank = ctx.dev, ctx.world, ctx.rank
        self.dtype_name = dtype_name or wl["dtype"]
        dtype = torch.bfloat16 if self.dtype_name == "bf16" else torch.float32
        B, MU, C, D = wl["B"], wl["MU"], wl["C"], wl["D"]
        self.Bu = B * MU
        g = torch.Generator().manual_seed(seed + rank)
        self.model = build_model(wl).to(dev)
        self.head = None
        self.K_global = 0
        if wl["kind"] == "comatch":
            # comatch.py:91: queue_size = queue_batch*(MU+1)*BATCH_SIZE.  cfg 2, data parallel: BATCH_SIZE is the global batch,
            # so the ring grows with the rank count (weak scaling).  cfg 4: the 65536-row ring is fixed and split over the ranks.
            self.K_global = wl["K"] * (world if wl["k_scales_with_ranks"] else 1)
            self.head = CoMatchHead(C, D, self.K_global, wl["thr"], enqueue_mode="always", device=dev, dtype=dtype,
                                    process_group=ctx.pg, exchange=exchange)
        host, self.keys = make_batches(wl, g, 4, dtype)
        self.host = [{k: v.pin_memory() for k, v in b.items()} for b in host]
        self.resident = [{k: v.to(dev) for k, v in b.items()} for b in self.host]
        # ema_overlap: the EMA launch is a parallel branch of the step (side stream forked at the start of the step, joined at
        # its end) -- in the trainer EMA(t) runs next to the head of step t+1 and is joined ahead of optimizer.step(t+1)
        self.ema_overlap = bool(ema_overlap)
        self.ema = ModelEMA(self.model, decay=wl["decay"], device=dev, overlap=self.ema_overlap,
                            overlap_ctas=int(os.environ.get("B200SSL_EMA_CTAS", "0")) or None)
        S.perturb_(self.model, torch.Generator(device=dev).manual_seed(5))        # m != e, like after an optimizer step
        self.grad_keys = {"comatch": ("logits_u_s0", "feats_u_s0", "feats_u_s1"), "fixmatch": ("logits_u_s",),
                          "semiformer": ("logits_u_s", "logits_u_s_trans")}[wl["kind"]]
        # launches of libb200ssl.so per step.  CoMatch, one GPU: smooth, rows (DA + finalize + enqueue), contrast fwd, contrast
        # bwd (+ scale), ema.  N > 1: + the side-stream enqueue (peer-memory bank), + enqueue and three exchange launches
        # (own peer exchanges), + enqueue (NCCL, its kernels are not counted).  FixMatch / SemiFormer: head, scale(s), ema.
        self.launches_per_step = {"comatch": 5, "fixmatch": 3, "semiformer": 4}[wl["kind"]]
        if world > 1 and self.head is not None:
            self.launches_per_step += {"replicated": 1, "direct": 1, "peer": 4, "collective": 1}[self.head.exchange]
        if self.ema_overlap and self.ema.overlap_mode == "capped" and self.ema.overlap_delay_ns > 0:
            self.launches_per_step += 1              # the one-thread head-start kernel ahead of the overlapped update
        one = torch.ones((), dtype=torch.float32, device=dev)
        head, ema, model, keys = self.head, self.ema, self.model, self.keys

        # masked form of the overlapped update: its placement is independent of the head's, so it is forked ahead of the inputs
        # (end-to-end graph: under the H2D copy); capped form: forked together with the head's first kernel
        prefork = self.ema_overlap and self.ema.overlap_mode == "masked"

        def pre():
            if prefork:
                ema.update(model)

        def main_part(batch):
            if self.ema_overlap and not prefork:
                ema.update(model)                                                 # side stream; joined below
            for k in self.grad_keys:
                batch[k].grad = None
                batch[k].requires_grad_(True)
            if wl["kind"] == "comatch":
                total = head.total_loss(*[batch[k] for k in keys], lambda_u=wl["lambda_u"],
                                        lambda_c=wl["lambda_c"])[0]                # comatch.py:222 (unlabeled part)
            elif wl["kind"] == "semiformer":
                la, lb, _ = consistency_loss_dual(batch["logits_u_w"], batch["logits_u_s"], batch["logits_u_s_trans"], T=1.0,
                                                  p_cutoff=wl["thr"])              # semiformer.py:129-131
                total = wl["lambda_u"] * (la + lb)
            else:
                lu, _ = consistency_loss(batch["logits_u_w"], batch["logits_u_s"], T=1.0, p_cutoff=wl["thr"])
                total = wl["lambda_u"] * lu                                        # fixmatch.py:118
            total.backward(gradient=one)                                          # cached ones scalar: no fill kernel per step
            if self.ema_overlap:
                ema.join()
            else:
                ema.update(model)
            return total

        def step(batch):                                                          # the eager step
            pre()
            return main_part(batch)

        self.step, self.pre, self.main_part = step, pre, main_part
        self._GraphedStep = GraphedStep
        self.graphed = None

    def capture(self):
        n_rows = self.wl["B"] + self.Bu
        head = self.head
        self.graphed = self._GraphedStep(lambda b: self.main_part(b), self.resident[0], self.ctx.dev, warmup=3, pre_fn=self.pre,
                                         on_replay=(lambda: head.note_graph_replay(n_rows)) if head is not None else None,
                                         after_capture=(lambda: head.sync_ptr_from_device()) if head is not None else None,
                                         high_priority=self.ema_overlap)
        return self.graphed

    def release(self):
        """Captured graphs hold peer-memory / NCCL work: drop them before arenas and the communicator go away."""
        import gc
        if self.graphed is not None:
            self.graphed.graph = self.graphed.graph_host = None
            self.graphed = None
        gc.collect()
        torch.cuda.synchronize(self.ctx.dev)
        if self.head is not None and self.ctx.world > 1:
            if self.head.peer_timeouts():
                raise RuntimeError(f"rank {self.ctx.rank}: peer-memory waits timed out; no valid measurement")
            self.head.close()


def multirank_parity(ctx, wl, exchange, steps=3):
    """Before anything is timed at N > 1: ``steps`` eager steps of every rank on a pre-filled bank against the single-process
    oracle for the concatenated batch with the whole bank (oracle.comatch_head_sharded), checked on rank 0:
    ring pointer and copied embedding rows bit-exact, probability rows / smoothed probabilities / loss within the bf16
    tolerance (1e-2).  Raises on a mismatch -- a fast wrong step is not a measurement."""
    from endoscopy_image_classification_b200.comatch_head import CoMatchHead
    dist, dev, R, rank = ctx.dist, ctx.dev, ctx.world, ctx.rank
    C, D, B, MU = wl["C"], wl["D"], wl["B"], wl["MU"]
    K = wl["K"] * (R if wl["k_scales_with_ranks"] else 1)
    dtype = torch.bfloat16 if wl["dtype"] == "bf16" else torch.float32
    head = CoMatchHead(C, D, K, wl["thr"], enqueue_mode="always", device=dev, dtype=dtype, process_group=ctx.pg, exchange=exchange)
    g0 = torch.Generator().manual_seed(4242)                       # the same bank content on every rank
    qf = torch.nn.functional.normalize(torch.randn(K, D, generator=g0), dim=1).to(dtype)
    qp = torch.softmax(2.0 * torch.randn(K, C, generator=g0), 1).to(dtype)
    lo, hi = (0, K) if head.exchange == "replicated" else (head.geom.shard_begin, head.geom.shard_begin + head.geom.shard_rows)
    head.queue_feats.copy_(qf[lo:hi])
    head.queue_probs.copy_(qp[lo:hi])
    if head.queue_probs_t is not None:
        head.queue_probs_t[:C].copy_(qp[lo:hi].t())
    ctx.barrier()
    batches, keys = make_batches(wl, torch.Generator().manual_seed(777 + rank), steps, dtype)
    mine = []
    for b in batches:
        d = {k: v.to(dev) for k, v in b.items()}
        for k in ("logits_u_s0", "feats_u_s0", "feats_u_s1"):
            d[k].requires_grad_(True)
        total = head.total_loss(*[d[k] for k in keys], lambda_u=wl["lambda_u"], lambda_c=wl["lambda_c"])[0]
        total.backward()
        mine.append({"total": float(total), "probs": head.last["probs"].float().cpu(), "mask": head.last["mask"].cpu(),
                     "ptr": int(head.ptr_state[0]), "host_ptr": head.queue_ptr})
    torch.cuda.synchronize(dev)
    payload = {"batches": batches, "outs": mine, "qf": head.queue_feats.float().cpu(), "qp": head.queue_probs.float().cpu(),
               "timeouts": head.peer_timeouts(), "exchange": head.exchange}
    gathered = [None] * R
    dist.gather_object(payload, gathered if rank == 0 else None, dst=0)
    head.close()
    verdict = [None]
    if rank == 0:
        from oracle import ssl_oracle as O
        state = O.CoMatchState.zeros(K, D, C)
        state.queue_feats.copy_(qf.float())
        state.queue_probs.copy_(qp.float())
        hist = [[] for _ in range(R)]
        worst, ok, why = 0.0, True, ""
        n = B + B * MU
        for s in range(steps):
            inputs = [{k: (v.float() if v.is_floating_point() else v) for k, v in gathered[r]["batches"][s].items()} for r in range(R)]
            state.queue_probs = state.queue_probs.to(dtype).float()            # the device bank stores bf16 rows
            ref = O.comatch_head_sharded(state, hist, inputs, thr=wl["thr"], num_classes=C)
            for r in range(R):
                got = gathered[r]["outs"][s]
                err = float((got["probs"].double() - ref[r]["probs"].double()).abs().max() / ref[r]["probs"].abs().max())
                worst = max(worst, err)
                if bool((got["mask"] == ref[r]["mask"]).all()):
                    want = float(wl["lambda_u"] * ref[r]["loss_u"] + wl["lambda_c"] * ref[r]["loss_contrast"])
                    worst = max(worst, abs(got["total"] - want) / max(abs(want), 1e-6))
                if not (got["ptr"] == got["host_ptr"] == state.queue_ptr == ((s + 1) * R * n) % K):
                    ok, why = False, f"ring pointer of rank {r} at step {s}: {got['ptr']} / {got['host_ptr']} vs {state.queue_ptr}"
        if gathered[0]["exchange"] == "replicated":
            banks = [(gathered[r]["qf"], gathered[r]["qp"]) for r in range(R)]
        else:
            banks = [(torch.cat([gathered[r]["qf"] for r in range(R)]), torch.cat([gathered[r]["qp"] for r in range(R)]))]
        rows_exact = all(torch.equal(f, state.queue_feats) for f, _ in banks)
        probs_err = max(float((p - state.queue_probs).abs().max()) for _, p in banks)
        timeouts = sum(gathered[r]["timeouts"] for r in range(R))
        ok = ok and rows_exact and probs_err < 1e-2 and worst < 1e-2 and timeouts == 0
        verdict[0] = {"ok": ok, "oracle": "oracle.comatch_head_sharded (single process, concatenated batch, whole bank)",
                      "ranks": R, "steps": steps, "exchange": gathered[0]["exchange"], "bank_rows": K,
                      "bank_feature_rows_bit_exact": rows_exact, "ring_pointer_exact": why == "", "bank_prob_rows_max_abs_err": probs_err,
                      "probs_and_loss_max_rel_err": worst, "tolerance": 1e-2, "peer_timeouts": timeouts, "why": why}
    dist.broadcast_object_list(verdict, src=0)
    if not verdict[0]["ok"]:
        raise RuntimeError(f"multi-rank parity failed: {verdict[0]}")
    return verdict[0]


def measure_case(ctx, case, args, blocks, with_e2e=True, with_eager=True):
    """Eager numbers for the record, then the graph: device-resident blocks and end-to-end blocks."""
    out = {}
    dev = ctx.dev
    if with_eager:
        n_eager = min(args.steps, 100)
        ms = timed_blocks(ctx, lambda i: case.step(case.resident[i % len(case.resident)]), n_eager, args.warmup, 1)
        out["eager_ms_per_step"] = ms[0] / n_eager
    graphed = case.capture()
    torch.cuda.profiler.start()      # ncu --profile-from-start off: everything after the capture; before the barriers of the timed blocks
    ms_blocks = timed_blocks(ctx, lambda i: graphed.replay(), args.steps, args.warmup, blocks)
    out["blocks"] = block_stats(ms_blocks, args.steps)
    out["ms_per_step"] = out["blocks"]["median_ms_per_step"]
    if with_e2e:
        plain = [{k: v.clone() for k, v in hb.items()} for hb in case.host]      # ordinary (pageable) host tensors
        ms_e = timed_blocks(ctx, lambda i: graphed.replay_host(plain[i % len(plain)]), args.steps, args.warmup, blocks)
        out["e2e_blocks"] = block_stats(ms_e, args.steps)       # stage -> H2D -> head fwd+bwd -> EMA -> D2H loss -> sync
        out["e2e_ms_per_step"] = out["e2e_blocks"]["median_ms_per_step"]
        out["h2d_bytes_per_step"] = graphed.h2d_bytes
    torch.cuda.profiler.stop()
    torch.cuda.synchronize(dev)
    return out


def ema_roofline(ctx, case, args):
    """The dominant kernel alone: back-to-back EMA launches (300 MB each > L2) between two CUDA events."""
    n_ema = max(50, min(args.steps, 500))
    overlap, case.ema.overlap = case.ema.overlap, False          # on the timed (current) stream
    ms = timed_blocks(ctx, lambda i: case.ema.update(case.model), n_ema, 5, 3)
    case.ema.overlap = overlap
    return statistics.median(ms) / n_ema, n_ema


def fused_opt_ema_block(ctx, args):
    """optimizer.step() + ema.update() as ONE multi-tensor launch, replayed from a CUDA graph (device-resident
    group scalars), against the eager pair torch fused Adam + the EMA kernel.  ModelwEmb-R50, Adam."""
    from copy import deepcopy

    from endoscopy_image_classification_b200 import synthetic as S
    from endoscopy_image_classification_b200.ema import ModelEMA
    from endoscopy_image_classification_b200.fused_step import FusedOptimizerEMA
    dev = ctx.dev
    model = S.modelwemb_like("resnet50", 23, 64).to(dev)
    gen = torch.Generator(device=dev).manual_seed(3)
    for p in model.parameters():
        p.grad = 1e-3 * torch.randn(p.shape, generator=gen, device=dev)
    ref_model = deepcopy(model)
    for p, q in zip(ref_model.parameters(), model.parameters()):
        p.grad = q.grad.clone()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    ema = ModelEMA(model, decay=0.999, device=dev)
    fused = FusedOptimizerEMA(opt, ema, model)
    opt_ref = torch.optim.Adam(ref_model.parameters(), lr=1e-3, fused=True)
    ema_ref = ModelEMA(ref_model, decay=0.999, device=dev)
    out = {"model": "ModelwEmb-ResNet-50, Adam", "reference": "code/fixmatch.py:123,127 (optimizer.step(); ema_model.update(model))"}
    if not hasattr(fused, "capture"):
        n = 50
        out["fused_eager_ms"] = statistics.median(timed_blocks(ctx, lambda i: fused.step(), n, 5, 3)) / n
    else:
        graph = fused.capture()
        n = 50
        out["fused_graph_ms"] = statistics.median(timed_blocks(ctx, lambda i: graph.replay(), n, 5, 3)) / n
        out["fused_eager_ms"] = statistics.median(timed_blocks(ctx, lambda i: fused.step(), n, 5, 3)) / n

    def pair(i):
        opt_ref.step()
        ema_ref.update(ref_model)
    out["torch_fused_adam_plus_ema_kernel_ms"] = statistics.median(timed_blocks(ctx, pair, 50, 5, 3)) / 50
    out["bytes_per_step"] = fused.bytes_per_step if hasattr(fused, "bytes_per_step") else None
    return out


def run_b200(args, wl, rank, world, local_rank):
    from endoscopy_image_classification_b200 import _native as N
    N.lib()                                   # fail loudly when the extension is missing
    ctx = Ctx(rank, world, local_rank)
    blocks = args.blocks or (25 if args.steps <= 200 else 9)
    sampler = ClockSampler(local_rank)
    time.sleep(1.0)                           # let nvidia-smi start before the timed regions
    line_extra = {}

    # ---------------- multi-rank parity (before timing) ------------------------------------
    parity = None
    if world > 1 and wl["kind"] == "comatch" and not args.no_parity:
        parity = multirank_parity(ctx, wl, args.exchange)

    # ---------------- the headline workload -------------------------------------------------
    case = Case(ctx, wl, exchange=args.exchange, ema_overlap=args.ema_overlap)
    m = measure_case(ctx, case, args, blocks)
    ema_ms, n_ema = ema_roofline(ctx, case, args)
    ema_ms = ctx.max_over_ranks([ema_ms])[0]
    plan = case.ema.plan
    exchange = case.head.exchange if (case.head is not None and world > 1) else None
    K_global = case.K_global
    launches = case.launches_per_step
    case.release()
    del case

    # ---------------- extra keyed blocks (same run, same clocks) ----------------------------
    if args.workload == "cfg2" and not args.no_extras:
        # cfg 4: the 65536-row bank; N > 1: sharded over the ranks in NVLink peer memory, K3 reads every shard in place
        wl4 = WORKLOADS["cfg4"]
        ex4 = "auto" if world == 1 else "direct"
        par4 = multirank_parity(ctx, wl4, ex4) if (world > 1 and not args.no_parity) else None
        c4 = Case(ctx, wl4, exchange=ex4, ema_overlap=args.ema_overlap)
        m4 = measure_case(ctx, c4, args, blocks, with_e2e=True, with_eager=False)
        K4, R = wl4["K"], world
        line_extra["cfg4"] = {
            "workload": wl4["desc"], "n_gpus": world, "value": world * c4.Bu / (m4["ms_per_step"] * 1e-3), "unit": UNIT,
            "ms_per_step": m4["ms_per_step"], "timing": m4["blocks"],
            "e2e": {"value": world * c4.Bu / (m4["e2e_ms_per_step"] * 1e-3), "unit": UNIT, "ms_per_step": m4["e2e_ms_per_step"]},
            "bank_rows_global": K4, "bank_rows_per_rank": K4 // R, "bank_layout": "sharded" if R > 1 else "local",
            "bank_exchange": c4.head.exchange if R > 1 else None,
            # bytes a rank pulls over NVLink per step: every remote key tile once (K3 loops over the row tiles inside the CTA)
            "nvlink_read_bytes_per_rank_step": {"algorithmic_K(D+C)s(R-1)/R": K4 * (wl4["D"] + wl4["C"]) * 2 * (R - 1) // R,
                                                "moved_K(D+32)s(R-1)/R": K4 * (wl4["D"] + 32) * 2 * (R - 1) // R},
            "nvlink_write_bytes_per_rank_step": (wl4["B"] + c4.Bu) * (wl4["D"] + wl4["C"] + 32) * 2 * (R - 1) // R if R > 1 else 0,
            "parity": par4, "gpu_launches_per_step": c4.launches_per_step}
        c4.release()
        del c4
        if world == 1:
            # the reference's own precision: fp32 logits / embeddings / bank (exact-fp32 similarity kernels)
            c32 = Case(ctx, wl, dtype_name="f32", ema_overlap=args.ema_overlap)
            m32 = measure_case(ctx, c32, args, blocks, with_e2e=False, with_eager=False)
            line_extra["fp32"] = {"workload": wl["desc"].replace("bf16", "fp32 storage"), "dtype": "f32",
                                  "value": c32.Bu / (m32["ms_per_step"] * 1e-3), "unit": UNIT, "ms_per_step": m32["ms_per_step"],
                                  "timing": m32["blocks"], "tolerance": "1e-5 relative vs the reference (fp32 goldens)"}
            c32.release()
            del c32
            try:
                line_extra["fused_opt_ema"] = fused_opt_ema_block(ctx, args)
            except Exception as e:  # the extra block must not cost the headline
                line_extra["fused_opt_ema"] = {"error": repr(e)[:300]}
    clocks = sampler.stop()

    if rank == 0:
        peak, peak_src = peaks()
        achieved = plan.bytes_per_update / (ema_ms * 1e-3) / 1e9
        traffic, traffic_src = None, None
        tfile = REPO / "profiles" / "ema_traffic.json"
        if tfile.exists() and wl["kind"] == "comatch":          # the capture is of the ModelwEmb-ResNet-50 update
            try:
                tj = json.loads(tfile.read_text())
                traffic, traffic_src = float(tj["traffic_bytes_per_launch"]), tj["source"]
            except Exception:
                pass
        Bu = wl["B"] * wl["MU"]
        cpu = cpu_baseline_sample(wl) if world == 1 and not args.no_cpu_baseline else None
        ms, e2e_ms = m["ms_per_step"], m["e2e_ms_per_step"]
        layout = None if exchange is None else ("replicated" if exchange == "replicated" else "sharded")
        line = {"metric": METRIC, "value": world * Bu / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": wl["dtype"], "data": "synthetic",
                "config": {"workload": wl["desc"], "global_batch_unlabeled": world * Bu, "per_gpu_unlabeled": Bu,
                           "bank_rows_global": K_global, "bank_layout": layout, "parallelism": f"dp{world}",
                           "bank_exchange": (None if exchange is None else
                                             {"replicated": "every rank keeps the whole ring in NVLink peer memory; the enqueue block is written "
                                                            "through into every copy (the only exchange of a step); two epoch flags per step, one "
                                                            "extra (side-stream) enqueue launch, no collective",
                                              "direct": "one shard per rank in NVLink peer memory: K3 reads every shard in place (TMA over NVLink, "
                                                        "each remote key tile once), the enqueue stores into the owning shard; two epoch flags "
                                                        "per step, one extra (side-stream) enqueue launch, no collective",
                                              "peer": "own kernels over NVLink peer memory (csrc/peer.cu): all-gather, reduce-scatter, "
                                                      "all-gather per step",
                                              "collective": "NCCL collectives"}[exchange]),
                           "ema_state": {"entries": plan.n_entries, "unique_storages": plan.n_unique,
                                         "unique_elems": plan.unique_elems, "blocks": plan.n_blocks},
                           "execution": "whole step (head fwd+bwd + EMA) captured once in a CUDA graph and replayed; "
                                        "eager_ms_per_step is the same API without the graph"
                                        + ("; the EMA launch is a parallel branch of the step graph (ModelEMA(overlap=True): side stream "
                                           "forked at the start of the step and joined at its end, the head's stream at high priority, the "
                                           "update's grid capped at 400 persistent CTAs so that the SMs the head's first kernel holds stay "
                                           "free for the rest of the head) -- EMA(t) only has to finish before optimizer.step(t+1), so in "
                                           "the trainer it runs next to the following step's head" if args.ema_overlap
                                           else "; EMA after the backward, in series"),
                           "ema_overlap": bool(args.ema_overlap),
                           "timing": {"method": f"{m['blocks']['blocks']} blocks of [barrier+sync, {args.warmup} untimed replays, event, "
                                                f"{args.steps} timed replays, event]; per block MAX over ranks; the line reports the MEDIAN block",
                                      **m["blocks"], "e2e": m["e2e_blocks"]},
                           "eager_ms_per_step": m.get("eager_ms_per_step"),
                           "l2": "no explicit flush: the EMA kernel streams 300 MB/step (> 126 MB L2); head inputs (~0.3 MB) "
                                 "come straight from the backbone in training, i.e. L2-resident there too"},
                "roofline": {"kernel": "ema_multi_tensor_kernel", "bound": "hbm", "achieved": achieved, "peak": peak,
                             "unit": "GB/s", "frac": achieved / peak,
                             # DRAM bytes of one launch of the same kernel on the same state, from the committed ncu --set full
                             # capture (profiles/ema_traffic.json, written by tools/ncu_extract.py); null when absent
                             "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": plan.bytes_per_update, "avg_launch_ms": ema_ms,
                             "timed": f"{n_ema} back-to-back launches between two CUDA events on the launching stream (median of 3)"},
                "e2e": {"value": world * Bu / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": m["h2d_bytes_per_step"],
                        "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms},
                "gpu_launches": launches * args.steps,
                "clocks": clocks}
        if parity is not None:
            line["parity"] = parity
        line.update(line_extra)
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        ctx.barrier()
        ctx.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--blocks", type=int, default=0, help="timed blocks (0 = 25 for short runs, 9 for long ones)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--ema-overlap", type=int, default=1, choices=[0, 1],
                    help="1: the EMA launch runs as a parallel branch of the step (default); 0: after the backward, in series")
    ap.add_argument("--bank-rows", type=int, default=0, help="override the workload's bank rows per rank (tuning aid)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the cfg4 / fp32 / fused-optimizer blocks")
    ap.add_argument("--no-parity", action="store_true", help="skip the multi-rank parity check before timing")
    ap.add_argument("--exchange", default="auto", choices=["auto", "replicated", "direct", "peer", "collective"],
                    help="multi-rank bank: write-through copies or directly addressed shards in NVLink peer memory (auto), "
                         "own peer-memory exchange kernels, or NCCL collectives")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    wl = WORKLOADS[args.workload]
    if args.bank_rows:                         # tuning aid: e.g. cfg 2's N = 8 ring (20480 rows) on one GPU
        wl = dict(wl, K=args.bank_rows, desc=wl["desc"] + f" [bank rows overridden: {args.bank_rows}]")
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
