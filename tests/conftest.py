import json
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

REPO = Path(__file__).resolve().parents[1]
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))
GOLDEN = REPO / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    with np.load(GOLDEN / name, allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def golden_meta(z):
    return json.loads(str(z["meta"]))


def T(a, device="cpu", dtype=None):
    t = torch.from_numpy(np.array(a, copy=True, order="C"))
    if dtype is not None:
        t = t.to(dtype)
    return t.to(device)


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def assert_labels_match(idx, idx_ref, probs_ref_sorted_gap, what="argmax", tol=1e-6):
    """Bit-exact index check, margin aware: a mismatch is tolerated only
    where the reference's top-2 gap is below `tol` (a tie graze, not a bug)."""
    bad = (idx.cpu() != idx_ref.cpu()).nonzero().flatten()
    for i in bad.tolist():
        assert float(probs_ref_sorted_gap[i]) < tol, f"{what} mismatch at row {i} with top-2 gap {float(probs_ref_sorted_gap[i])}"


def assert_mask_match(mask, mask_ref, score_ref, thr, what="mask", tol=1e-6):
    bad = (mask.cpu() != mask_ref.cpu()).nonzero().flatten()
    for i in bad.tolist():
        assert abs(float(score_ref[i]) - thr) < tol, f"{what} mismatch at row {i}: score {float(score_ref[i])} vs thr {thr}"


def top2_gap(probs):
    t = probs.double().topk(2, dim=1).values
    return (t[:, 0] - t[:, 1]).cpu()


@pytest.fixture(autouse=True)
def _reset_process_wide_knobs():
    """ModelEMA(overlap=True) sets a process-wide SM budget for the head's launch planners: tests must not leak it."""
    yield
    try:
        from endoscopy_image_classification_b200 import _native as N
        N.lib().b200ssl_set_head_sm_budget(0)
    except Exception:
        pass
