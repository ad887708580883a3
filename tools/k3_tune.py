#!/usr/bin/env python
"""A/B of the tensor-core K3 (csrc/bank_tc.cu) on one GPU: launch plan (legacy round-1 plan / planner / forced row loop) x
share of the exponentials on the FMA-pipe polynomial, per problem size.  Times graphs of back-to-back launches with CUDA
events and checks every variant against fp64 math on the same bf16 operands.  One JSON line per (size, variant).

    python tools/k3_tune.py [--reps 20]
"""
import argparse
import ctypes
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from endoscopy_image_classification_b200 import _native as N  # noqa: E402
from endoscopy_image_classification_b200.comatch_head import CoMatchHead  # noqa: E402

C, D = 23, 64


def graph_time(fn, reps, replays=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    for _ in range(2):
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
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--sizes", default="448x2560,448x20480,448x65536,1792x16384,3584x32768,3584x65536,14336x65536")
    a = ap.parse_args()
    lib = N.lib()
    plan = (ctypes.c_int32 * 3)()
    for size in a.sizes.split(","):
        rows, K = (int(x) for x in size.split("x"))
        g = torch.Generator().manual_seed(rows + K)
        nf = lambda n: torch.nn.functional.normalize(torch.randn(n, D, generator=g), dim=1).to(torch.bfloat16)
        f, qf = nf(rows), nf(K)
        qp = torch.softmax(2.0 * torch.randn(K, C, generator=g), 1).to(torch.bfloat16)
        head = CoMatchHead(C, D, K, 0.9, enqueue_mode="always", dtype=torch.bfloat16)
        head.queue_feats.copy_(qf)
        head.queue_probs.copy_(qp)
        head.queue_probs_t[:C].copy_(qp.t())
        fd = f.cuda()
        ref = None
        if rows * K <= 3584 * 65536:
            A = torch.exp(fd.double() @ qf.cuda().double().t() / 0.2)
            ref = (A.sum(1), A @ qp.cuda().double())
            del A
        variants = [(1, 0), (0, 0), (0, 8), (0, 12), (0, 16), (0, 20)]
        if rows <= 512:
            variants += [(4, 0), (4, 12)]
        for mt, poly in variants:
            lib.b200ssl_debug_set_k3(mt, poly)
            lib.b200ssl_debug_smooth_plan(rows, K, 0, plan)
            rowsum, numer = head._k_smooth(fd)
            torch.cuda.synchronize()
            err = None
            if ref is not None:
                err = max(float((rowsum.double() - ref[0]).abs().max() / ref[0].abs().max()),
                          float((numer.double() - ref[1]).abs().max() / ref[1].abs().max()))
            us = graph_time(lambda: head._k_smooth(fd), a.reps if rows * K < 3584 * 65536 else max(4, a.reps // 4))
            flop = 2.0 * rows * K * (D + 32)
            print(json.dumps({"rows": rows, "bank": K, "force_mt": mt, "poly_of_32": poly, "plan_mt_cluster_nouter": list(plan),
                              "us": round(us, 2), "tflops_executed": round(flop / us * 1e-6, 1), "max_rel_err_vs_fp64": err}), flush=True)
        lib.b200ssl_debug_set_k3(0, -1)
        del head


if __name__ == "__main__":
    main()
