#!/usr/bin/env python
"""Per-kernel GPU time of the head's launches, measured without host launch overhead:
each kernel is captured N times into one CUDA graph (back-to-back dependent nodes) and the
graph is replayed between two CUDA events.  Prints microseconds per launch.

    python tools/microbench.py [--rows 448] [--bank 2560] [--dtype bf16|f32] [--reps 50]
"""
import argparse
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from endoscopy_image_classification_b200 import synthetic as S  # noqa: E402
from endoscopy_image_classification_b200.comatch_head import CoMatchHead  # noqa: E402
from endoscopy_image_classification_b200.ema import ModelEMA  # noqa: E402
from endoscopy_image_classification_b200.loss import fixmatch_head  # noqa: E402


def graph_time(fn, reps, replays=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / (reps * replays)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=448)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--bank", type=int, default=2560)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--reps", type=int, default=50)
    ap.add_argument("--ema-arch", default="resnet50")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    dt = torch.bfloat16 if a.dtype == "bf16" else torch.float32
    C, D = 23, 64
    mu = a.rows // a.batch
    g = torch.Generator().manual_seed(0)
    b = {k: v.to(dev) for k, v in S.comatch_step_inputs(g, a.batch, mu, D, C, dtype=dt).items()}
    head = CoMatchHead(C, D, a.bank, 0.9, enqueue_mode="always", device=dev, dtype=dt)
    # fill the bank
    head.queue_feats.copy_(S.rownorm(torch.randn(a.bank, D, generator=g)).to(dt))
    qp = torch.softmax(torch.randn(a.bank, C, generator=g), 1).to(dt)
    head.queue_probs.copy_(qp)
    if head.queue_probs_t is not None:
        head.queue_probs_t[:C].copy_(qp.t())
    lw, ls0, fw, fs0, fs1, fx, tx = (b[k] for k in ("logits_u_w", "logits_u_s0", "feats_u_w", "feats_u_s0", "feats_u_s1",
                                                    "feats_x", "targets_x"))
    out = {}
    head._k_da(lw)
    rowsum, numer = head._k_smooth(fw)
    fin = head._k_finalize(lw, ls0, rowsum, numer)
    stats, _ = head._k_contrast_fwd(fs0, fs1, fin["probs"], fin["scalars"], probs_hl=fin["probs_hl"])
    one = torch.ones(1, device=dev)
    out["comatch_da"] = graph_time(lambda: head._k_da(lw), a.reps)
    out["bank_smooth"] = graph_time(lambda: head._k_smooth(fw), a.reps)
    out["comatch_finalize"] = graph_time(lambda: head._k_finalize(lw, ls0, rowsum, numer), a.reps)
    out["bank_enqueue"] = graph_time(lambda: head._k_enqueue(fw, fx, fin["probs_orig"], tx, 0, a.rows + a.batch), a.reps)
    hl = fin["probs_hl"]
    out["contrast_fwd"] = graph_time(lambda: head._k_contrast_fwd(fs0, fs1, fin["probs"], fin["scalars"], probs_hl=hl), a.reps)
    out["contrast_bwd"] = graph_time(lambda: head._k_contrast_bwd(fs0, fs1, fin["probs"], stats, one, probs_hl=hl), a.reps)
    out["scale_inplace"] = graph_time(lambda: head._k_scale(fin["grad_s0"], one), a.reps)
    w = b["logits_u_w"]
    out["fixmatch_head_fwd_bwd"] = graph_time(lambda: fixmatch_head(w, ls0, None, 0.95), a.reps)
    model = S.modelwemb_like(a.ema_arch, C, D).to(dev)
    ema = ModelEMA(model, 0.999, device=dev)
    ema.update(model)
    out["ema_multi_tensor(%s)" % a.ema_arch] = graph_time(lambda: ema.update(model), 10, replays=10)
    out["_config"] = vars(a)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
