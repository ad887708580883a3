#!/usr/bin/env python
"""time `optimizer.step(); ema.update(model)` (torch foreach / torch fused + the multi-tensor EMA
kernel) against `FusedOptimizerEMA.step()` on the ModelwEmb-R50 shaped state.   python tools/opt_bench.py"""
import copy
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from endoscopy_image_classification_b200 import synthetic as S  # noqa: E402
from endoscopy_image_classification_b200.ema import ModelEMA  # noqa: E402
from endoscopy_image_classification_b200.fused_step import FusedOptimizerEMA  # noqa: E402
from endoscopy_image_classification_b200.optimizer import set_weight_decay  # noqa: E402

dev = torch.device("cuda:0")
out = {}


def timed(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return round(1e3 * e0.elapsed_time(e1) / n, 1)


for kind in ("adam", "adamw", "sgd"):
    model = S.modelwemb_like("resnet50", 23, 64).to(dev)
    n_params = sum(p.numel() for p in model.parameters())
    for p in model.parameters():
        p.grad = torch.randn_like(p) * 1e-2

    def make(m, **kw):
        groups = set_weight_decay(m)
        if kind == "sgd":
            return torch.optim.SGD(groups, momentum=0.9, nesterov=True, lr=1e-3, weight_decay=0.05, **kw)
        if kind == "adamw":
            return torch.optim.AdamW(groups, lr=1e-3, weight_decay=0.05, **kw)
        return torch.optim.Adam(groups, lr=1e-3, **kw)

    res = {"params": n_params}
    for label, kw in (("foreach", {}), ("torch_fused", {"fused": True})):
        m = copy.deepcopy(model)
        for p, q in zip(m.parameters(), model.parameters()):
            p.grad = q.grad.clone()
        opt, ema = make(m, **kw), ModelEMA(m, 0.999, device=dev)
        res[f"{label}_step_plus_ema_us"] = timed(lambda: (opt.step(), ema.update(m)))
    m = copy.deepcopy(model)
    for p, q in zip(m.parameters(), model.parameters()):
        p.grad = q.grad.clone()
    fused = FusedOptimizerEMA(make(m), ModelEMA(m, 0.999, device=dev), m)
    res["b200_fused_us"] = timed(fused.step)
    bytes_per = (36 if kind != "sgd" else 28) * n_params
    res["b200_fused_algorithmic_GBps"] = round(bytes_per / res["b200_fused_us"] / 1e3, 1)
    out[kind] = res
print(json.dumps(out, indent=1))
