#!/usr/bin/env python
"""A/B of the tensor-core K3 (csrc/bank_tc.cu) on one GPU: launch plans (row tiles per CTA x cluster size x clusters per
row group; 0 = the planner's choice) and the share of the exponentials on the FMA-pipe polynomial (--poly N of 32; + 100 keeps
the P tile in shared memory, the round-1 form), per problem size.
Times graphs of back-to-back launches with CUDA events and checks every variant against fp64 math on the same bf16
operands.  One JSON line per (size, variant); the first line reports the cluster residency the driver grants.

    python tools/k3_tune.py [--reps 20] [--sizes 448x2560,...] [--poly 0,8,16]
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


def plans_for(rows, K):
    """(mt, cluster, nouter) candidates; (0, 0, 0) = planner."""
    row_tiles, ktiles = (rows + 127) // 128, (K + 127) // 128
    out = [(0, 0, 0)]
    for mt in (1, 2, 4):
        if mt > row_tiles and mt > 1:
            continue
        groups = (row_tiles + mt - 1) // mt
        for cl in (1, 2, 4, 8):
            if cl > ktiles:
                continue
            cap = {1: 148, 2: 74, 4: 33, 8: 16}[cl]
            for no in sorted({1, max(1, cap // groups), max(1, cap // groups // 2), max(1, 2 * cap // groups)}):
                if cl * no <= ktiles and groups * cl * no <= 4 * 148:
                    out.append((mt, cl, no))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--sizes", default="448x2560,448x20480,448x65536,1792x16384,3584x65536")
    ap.add_argument("--poly", default="0")
    ap.add_argument("--planner-only", action="store_true")
    a = ap.parse_args()
    lib = N.lib()
    print(json.dumps({"max_active_clusters": {cl: lib.b200ssl_debug_max_active_clusters(cl) for cl in (1, 2, 4, 8)}}), flush=True)
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
        for mt, cl, no in ([(0, 0, 0)] if a.planner_only else plans_for(rows, K)):
            for poly in (int(x) for x in a.poly.split(",")):
                lib.b200ssl_debug_set_k3(mt, cl, no, poly)
                lib.b200ssl_debug_smooth_plan(rows, K, 0, plan)
                rowsum, numer = head._k_smooth(fd)
                torch.cuda.synchronize()
                err = None
                if ref is not None:
                    err = max(float((rowsum.double() - ref[0]).abs().max() / ref[0].abs().max()),
                              float((numer.double() - ref[1]).abs().max() / ref[1].abs().max()))
                us = graph_time(lambda: head._k_smooth(fd), a.reps if rows * K < 3584 * 65536 else max(4, a.reps // 4))
                flop = 2.0 * rows * K * (D + 32)
                print(json.dumps({"rows": rows, "bank": K, "forced": [mt, cl, no], "poly_of_32": poly, "plan_mt_cluster_nouter": list(plan),
                                  "us": round(us, 2), "tflops_executed": round(flop / us * 1e-6, 1), "max_rel_err_vs_fp64": err}), flush=True)
        lib.b200ssl_debug_set_k3(0, 0, 0, -1)
        del head


if __name__ == "__main__":
    main()
