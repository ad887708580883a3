"""the evaluation head.  CPU: every metric of ``utils.calculate_metrics`` (code/utils.py:38-55) derived from
the confusion matrix, against the golden vector produced by the REAL ``FixMatch.evaluate_one`` (tests/golden/evaluate_one.npz)
and against scikit-learn on random cases with absent classes.  GPU: the device accumulator against the same golden vector."""
import numpy as np
import pytest
import torch

from conftest import golden_meta, load_golden
from oracle import ssl_oracle as O


def test_oracle_evaluate_one_matches_reference_golden():
    z = load_golden("evaluate_one.npz")
    m = golden_meta(z)
    lg, y = torch.from_numpy(z["logits"]), torch.from_numpy(z["targets"])
    bs = m["batch_size"]
    out = O.evaluate_one([lg[i:i + bs] for i in range(0, m["N"], bs)], [y[i:i + bs] for i in range(0, m["N"], bs)], m["C"], bs)
    assert abs(out["loss_avg"] - float(z["loss_avg"])) < 1e-6 * float(z["loss_avg"])
    for k, v in zip(m["metric_keys"], z["metrics"]):
        assert abs(out["metric"][k] - v) < 1e-12, k
    assert np.array_equal(out["confusion"], z["confusion"])


def test_metrics_from_confusion_match_reference_and_sklearn():
    from endoscopy_image_classification_b200.evaluation import metrics_from_confusion
    z = load_golden("evaluate_one.npz")
    m = golden_meta(z)
    got = metrics_from_confusion(z["confusion"])
    for k, v in zip(m["metric_keys"], z["metrics"]):
        assert abs(got[k] - v) < 1e-12, k
    assert np.allclose(got["sen/spec"]["sensitivity"].values, z["sensitivity"], rtol=0, atol=1e-12)
    assert np.allclose(got["sen/spec"]["specificity"].values, z["specificity"], rtol=0, atol=1e-12)
    assert list(got["sen/spec"].columns) == ["class", "sensitivity", "specificity"]
    # random cases incl. classes that never occur / are never predicted (scikit-learn leaves them out of the macro means)
    from sklearn.metrics import f1_score, precision_score, recall_score
    rng = np.random.default_rng(0)
    for C, n, drop in ((23, 300, (3, 7)), (5, 40, (4,)), (2, 10, ())):
        targ = rng.integers(0, C, n)
        pred = np.where(rng.random(n) < 0.6, targ, rng.integers(0, C, n))
        for d in drop:
            targ[targ == d] = (d + 1) % C
            pred[pred == d] = (d + 1) % C
        pred[:3] = drop[0] if drop else pred[:3]          # predicted but never a target -> precision 0 for that class
        conf = np.zeros((C, C), dtype=np.int64)
        np.add.at(conf, (targ, pred), 1)
        got = metrics_from_confusion(conf)
        for avg in ("micro", "macro"):
            assert abs(got[f"{avg}/precision"] - precision_score(targ, pred, average=avg, zero_division=0)) < 1e-12
            assert abs(got[f"{avg}/recall"] - recall_score(targ, pred, average=avg, zero_division=0)) < 1e-12
            assert abs(got[f"{avg}/f1"] - f1_score(targ, pred, average=avg, zero_division=0)) < 1e-12


def test_eval_accumulator_has_no_cpu_path():
    from endoscopy_image_classification_b200.evaluation import EvalAccumulator
    with pytest.raises(RuntimeError):
        EvalAccumulator(23, "cpu")


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_eval_accumulator_matches_reference_golden(dtype):
    """One launch per batch, ONE device-to-host copy at the end: loss meter, confusion matrix and every metric equal the
    REAL reference's evaluate_one on the same scripted logits (bf16 logits: against the oracle on the rounded logits)."""
    from endoscopy_image_classification_b200.evaluation import EvalAccumulator
    z = load_golden("evaluate_one.npz")
    m = golden_meta(z)
    lg, y = torch.from_numpy(z["logits"]).to(dtype), torch.from_numpy(z["targets"])
    bs, N, C = m["batch_size"], m["N"], m["C"]
    acc = EvalAccumulator(C, "cuda", keep_predictions=True)
    for i in range(0, N, bs):
        acc.update(lg[i:i + bs].cuda(), y[i:i + bs].cuda())
    meter, metric = acc.finalize(bs)
    ref = O.evaluate_one([lg[i:i + bs].float() for i in range(0, N, bs)], [y[i:i + bs] for i in range(0, N, bs)], C, bs)
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert abs(meter.avg - ref["loss_avg"]) < tol * ref["loss_avg"] and meter.count == ref["loss_count"]
    assert abs(meter.val - ref["loss_val"]) < tol * ref["loss_val"]
    assert np.array_equal(acc.confusion, ref["confusion"])            # integer work: bit-exact
    pred, targ = acc.predictions()
    assert np.array_equal(pred, ref["pred"]) and np.array_equal(targ, ref["target"])
    for k in m["metric_keys"]:
        assert abs(metric[k] - ref["metric"][k]) < 1e-12, k
    if dtype == torch.float32:
        assert abs(meter.avg - float(z["loss_avg"])) < 1e-5 * float(z["loss_avg"])
        assert np.array_equal(acc.confusion, z["confusion"])
        for k, v in zip(m["metric_keys"], z["metrics"]):
            assert abs(metric[k] - v) < 1e-12, k
