#!/usr/bin/env python
"""In-kernel timeline of the tcgen05 smoothing kernel: clock64() / %globaltimer stamps per CTA (see B200SSL_STAMP in
csrc/common.cuh).  Prints cycles between stages, min / median / max over CTAs, and the wall-clock span of the launch.

    python tools/kernel_timeline.py ROWS BANK [MT CLUSTER NOUTER [POLY]]     (0 = planner)
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from endoscopy_image_classification_b200 import _native as N  # noqa: E402
from endoscopy_image_classification_b200 import synthetic as S  # noqa: E402
from endoscopy_image_classification_b200.comatch_head import CoMatchHead  # noqa: E402

rows, K = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (448, 2560)
force = [int(x) for x in (sys.argv[3:6] + ["0", "0", "0"])[:3]]
poly = int(sys.argv[6]) if len(sys.argv) > 6 else -1
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
N.lib().b200ssl_debug_set_k3(*force, poly)
head = CoMatchHead(23, 64, K, 0.9, enqueue_mode="always", device=dev, dtype=torch.bfloat16)
head.queue_feats.copy_(S.rownorm(torch.randn(K, 64, generator=g)).to(torch.bfloat16))
fw = S.rownorm(torch.randn(rows, 64, generator=g)).to(torch.bfloat16).to(dev)
for _ in range(3):
    head._k_smooth(fw)
torch.cuda.synchronize()
buf = torch.zeros(6 * 4096 * 16, dtype=torch.int64, device=dev)   # one region per instrumented kernel
N.lib().b200ssl_debug_set_timing_buffer(buf.data_ptr())
head._k_smooth(fw)
torch.cuda.synchronize()
N.lib().b200ssl_debug_set_timing_buffer(None)
t = buf.cpu().numpy().reshape(-1, 16)[:4096]
t = t[t[:, 0] != 0]
names = ["start", "setup done", "query tiles landed", "first S ready", "last exp done", "acc complete", "acc staged, tmem freed",
         "cluster fold done", "tickets taken", "outer folds done"]
import ctypes  # noqa: E402
plan = (ctypes.c_int32 * 3)()
N.lib().b200ssl_debug_smooth_plan(rows, K, 0, plan)
print(f"rows={rows} K={K} plan(mt, cluster, nouter)={list(plan)}: {len(t)} CTAs; cycles since CTA start (min / median / max), "
      f"1.965 GHz => 1000 cyc = 0.51 us")
for i in range(1, 10):
    col = t[:, i][t[:, i] != 0] - t[:, 0][t[:, i] != 0]
    if len(col):
        print(f"  {names[i]:24s} {col.min():8d} {int(np.median(col)):8d} {col.max():8d}   (n={len(col)})")
t0 = t[:, 10].min()
end = np.maximum(t[:, 11], t[:, 12])
print(f"  wall clock (ns): CTA starts spread {t[:, 10].max() - t0}, first start -> last CTA done {end.max() - t0}, "
      f"median CTA lifetime {int(np.median(end - t[:, 10]))}")
