#!/usr/bin/env python
"""K3 with fp32 storage (the reference's own precision): the tensor-core kernel on bf16 hi + mid operands against fp64 math
and against the exact-fp32 FFMA tiles (run once more with B200SSL_K3_F32_SIMT=1 for those).  One JSON line per size.

    python tools/k3_f32.py [--sizes 448x2560,...] [--reps 20]
"""
import argparse
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from endoscopy_image_classification_b200.comatch_head import CoMatchHead  # noqa: E402
from tools.k3_tune import graph_time  # noqa: E402

C, D = 23, 64


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="448x2560,448x2563,100x777,448x65536,3584x32768,3584x65536")
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    path = "ffma (exact fp32)" if os.environ.get("B200SSL_K3_F32_SIMT") else "tcgen05, bf16 hi+mid operands"
    for size in a.sizes.split(","):
        rows, K = (int(x) for x in size.split("x"))
        g = torch.Generator().manual_seed(rows + K)
        protos = torch.nn.functional.normalize(torch.randn(C, D, generator=g), dim=1)
        lab = torch.randint(0, C, (K,), generator=g)
        qf = torch.nn.functional.normalize(protos[lab] + 0.35 * torch.randn(K, D, generator=g), dim=1)
        f = torch.nn.functional.normalize(protos[torch.randint(0, C, (rows,), generator=g)] + 0.35 * torch.randn(rows, D, generator=g), dim=1)
        qp = torch.softmax(3.0 * torch.randn(K, C, generator=g) + 4.0 * torch.nn.functional.one_hot(lab, C), 1)
        head = CoMatchHead(C, D, K, 0.9, enqueue_mode="always", dtype=torch.float32)
        head.queue_feats.copy_(qf)
        head.queue_probs.copy_(qp)
        fd = f.cuda()
        rowsum, numer = head._k_smooth(fd)
        torch.cuda.synchronize()
        A = torch.exp(fd.double() @ qf.cuda().double().t() / 0.2)
        ref = (A @ qp.cuda().double()) / A.sum(1, keepdim=True)
        got = numer.double() / rowsum.double().unsqueeze(1)
        err = float((got - ref).abs().max() / ref.abs().max())
        err_rs = float((rowsum.double() - A.sum(1)).abs().max() / A.sum(1).abs().max())
        del A
        us = graph_time(lambda: head._k_smooth(fd), a.reps if rows * K < 3584 * 32768 else 4, replays=5)
        print(json.dumps({"rows": rows, "bank": K, "path": path, "us": round(us, 2), "smoothed_probs_max_rel_err_vs_fp64": err,
                          "rowsum_max_rel_err_vs_fp64": err_rs}), flush=True)
        del head


if __name__ == "__main__":
    main()
