#!/usr/bin/env python
"""In-kernel timelines of the tcgen05 contrastive kernels (clock64 stamps per CTA)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch  # noqa
from endoscopy_image_classification_b200 import _native as N, synthetic as S  # noqa
from endoscopy_image_classification_b200.comatch_head import CoMatchHead  # noqa
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 448
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
dt = torch.bfloat16
nf = lambda: S.rownorm(torch.randn(rows, 64, generator=g)).to(dt).to(dev)
f0, f1 = nf(), nf()
y = torch.randint(0, 23, (rows,), generator=g)
probs = torch.softmax(6.0 * torch.nn.functional.one_hot(y, 23).float() + torch.randn(rows, 23, generator=g), 1)
hi = probs.to(dt); lo = (probs - hi.float()).to(dt)
hl = torch.zeros(rows, 64, dtype=dt); hl[:, :23], hl[:, 32:55] = hi, lo
head = CoMatchHead(23, 64, 64, 0.9, dtype=dt, device=dev)
scal = torch.zeros(4, device=dev); dp, dhl = probs.to(dev), hl.to(dev); one = torch.ones(1, device=dev)
for _ in range(3):
    stats, _ = head._k_contrast_fwd(f0, f1, dp, scal, probs_hl=dhl)
    head._k_contrast_bwd(f0, f1, dp, stats, one, probs_hl=dhl)
torch.cuda.synchronize()
def run(fn, names):
    buf = torch.zeros(4 * 4096 * 16, dtype=torch.int64, device=dev)   # one region per instrumented kernel
    N.lib().b200ssl_debug_set_timing_buffer(buf.data_ptr())
    fn(); torch.cuda.synchronize()
    N.lib().b200ssl_debug_set_timing_buffer(None)
    t = buf.cpu().numpy().reshape(-1, 16); t = t[t[:, 0] != 0]
    print(f"  {len(t)} CTAs; cycles since CTA start (min / median / max)")
    for i, nm in enumerate(names, 1):
        ok = t[:, i] != 0
        if ok.any():
            col = t[ok, i] - t[ok, 0]
            print(f"    {nm:34s} {col.min():8d} {int(np.median(col)):8d} {col.max():8d}  (n={ok.sum()})")
print(f"contrast fwd rows={rows}")
run(lambda: head._k_contrast_fwd(f0, f1, dp, scal, probs_hl=dhl),
    ["setup", "first S/Q ready", "pass A done", "after exchange A", "pass B done", "after exchange B", "rows folded+sync+dealloc", "grid ticket",
     "B: row stats assembled", "B: chunk 0 compacted", "B: chunk 1 compacted"])
print(f"contrast bwd rows={rows}")
run(lambda: head._k_contrast_bwd(f0, f1, dp, stats, one, probs_hl=dhl),
    ["setup", "first S/Q ready", "last dZ written", "acc complete", "staged + cluster sync", "fold done"])
