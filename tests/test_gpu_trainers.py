"""Drop-in trainers on the GPU against the golden fixtures produced by the REAL reference
FixMatch.train_one / CoMatch.train_one (same scripted model outputs, same targets)."""
import numpy as np
import pytest
import torch

from conftest import T, golden_meta, load_golden, rel_err
from helpers import Loader, NoSched, ScriptedModel
from oracle import ssl_oracle as O

pytestmark = pytest.mark.gpu
C = 23


def _config(pkg_utils, **kw):
    A = pkg_utils.AttrDict
    base = dict(DATA=A(BATCH_SIZE=kw["B"], MU=kw["MU"], TARGET_NAME="target"),
                MODEL=A(NUM_CLASSES=C, LOW_DIM=kw.get("D", 64), NAME="scripted"),
                TRAIN=A(THRES=kw["thr"], T=1.0, USE_EMA=kw.get("ema", False), EMA_DECAY=0.999, LAMBDA_U=1.0, LAMBDA_C=1.0,
                        EVAL_STEP=kw.get("steps", 1), EVAL_STEP_SUP=kw.get("sup", 0), IS_FREEZE=False, CLS_WEIGHT=False,
                        BASE_LR=0.0, EPOCHS=1, FREQ_EVAL=1, SAVE_CP=kw.get("save", "/tmp/b200ssl_ckpt")))
    for k, v in kw.get("train_extra", {}).items():
        base["TRAIN"][k] = v
    return A(base)


def test_fixmatch_train_one_matches_reference_golden():
    from endoscopy_image_classification_b200 import utils
    from endoscopy_image_classification_b200.fixmatch import FixMatch
    z = load_golden("fixmatch_train_one.npz")
    m = golden_meta(z)
    B, Bu = m["B"], m["B"] * m["MU"]
    steps = [{"logits": T(z[f"s{i}/logits"])} for i in range(2)]
    model = ScriptedModel(steps, with_feats=False)
    tr = FixMatch(model, device="cuda")
    tr.get_dataloader((Loader([(torch.zeros(B, 1), T(z[f"s{i}/targets_x"])) for i in range(2)]),
                       Loader([((torch.zeros(Bu, 1), torch.zeros(Bu, 1)), None) for _ in range(2)])), None)
    cfg = _config(utils, B=B, MU=m["MU"], thr=m["thr"], steps=1)
    tr.get_config(cfg, optimizer=torch.optim.SGD(model.parameters(), lr=0.0), lr_scheduler=NoSched())
    for i in range(2):
        meter = tr.train_one(epoch=0)
        assert abs(meter.avg - float(z[f"s{i}/total_loss"])) < 1e-5 * abs(float(z[f"s{i}/total_loss"]))
        (leaf,) = model.seen[-1]
        assert rel_err(leaf.grad, T(z[f"s{i}/grad_logits"])) < 1e-5


@pytest.mark.parametrize("qb", [1, 5])
def test_comatch_train_one_matches_reference_golden(qb):
    from endoscopy_image_classification_b200 import utils
    from endoscopy_image_classification_b200.comatch import CoMatch
    z = load_golden(f"comatch_train_one_qb{qb}.npz")
    m = golden_meta(z)
    B, MU, D, n = m["B"], m["MU"], m["D"], m["nsteps"]
    Bu = B * MU
    steps = [{"logits": T(z[f"s{i}/logits"]), "feats": T(z[f"s{i}/feats"])} for i in range(n)]
    model = ScriptedModel(steps, with_feats=True)
    tr = CoMatch(model, device="cuda")
    tr.queue_batch = qb
    tr.get_dataloader((Loader([(torch.zeros(B, 1), T(z[f"s{i}/targets_x"])) for i in range(n)]),
                       Loader([((torch.zeros(Bu, 1),) * 3, None) for _ in range(1)])), None)
    tr.get_config(_config(utils, B=B, MU=MU, D=D, thr=m["thr"]), optimizer=torch.optim.SGD(model.parameters(), lr=0.0),
                  lr_scheduler=NoSched())
    assert tr.queue_size == qb * (MU + 1) * B and tr.queue_feats.shape == (tr.queue_size, D)
    for i in range(n):
        meter = tr.train_one(epoch=1)          # the unlabeled loader holds one batch -> one step
        assert abs(meter.avg - float(z[f"s{i}/total_loss"])) < 2e-5 * abs(float(z[f"s{i}/total_loss"]))
        lg, ft = model.seen[-1]
        assert rel_err(lg.grad, T(z[f"s{i}/grad_logits"])) < 1e-5
        assert rel_err(ft.grad, T(z[f"s{i}/grad_feats"])) < 1e-5
        assert tr.queue_ptr == int(z[f"s{i}/queue_ptr"])
        assert torch.equal(tr.queue_feats.cpu(), T(z[f"s{i}/queue_feats"]))
    assert len(tr.prob_list) == n


def test_semiformer_step_ema_and_checkpoint(tmp_path):
    """Two-head step == two reference-style consistency losses; EMA runs; checkpoint round-trips
    (incl. the CoMatch bank extension on the CoMatch trainer)."""
    from endoscopy_image_classification_b200 import utils
    from endoscopy_image_classification_b200.comatch import CoMatch
    from endoscopy_image_classification_b200.semiformer import SemiFormer
    g = torch.Generator().manual_seed(3)
    B, MU, thr = 8, 3, 0.7
    Bu = B * MU
    steps = [{"logits": 4 * torch.randn(B + 2 * Bu, C, generator=g), "logits2": 4 * torch.randn(B + 2 * Bu, C, generator=g)}]
    ty = torch.randint(0, C, (B,), generator=g)
    model = ScriptedModel(steps, with_feats=False, two_heads=True)
    tr = SemiFormer(model, device="cuda")
    tr.get_dataloader((Loader([(torch.zeros(B, 1), ty)]), Loader([((torch.zeros(Bu, 1), torch.zeros(Bu, 1)), None)])), None)
    tr.get_config(_config(utils, B=B, MU=MU, thr=thr, ema=True, sup=0, save=str(tmp_path)),
                  optimizer=torch.optim.SGD(model.parameters(), lr=0.0), lr_scheduler=NoSched())
    meter = tr.train_one(epoch=1)
    lc, lt = steps[0]["logits"], steps[0]["logits2"]
    w = lc[B:].chunk(2)[0]
    ref = (O.ce_loss(lc[:B], ty, reduction="mean") + O.ce_loss(lt[:B], ty, reduction="mean")
           + O.fixmatch_head_details(w, lc[B:].chunk(2)[1], thr)["loss"] + O.fixmatch_head_details(w, lt[B:].chunk(2)[1], thr)["loss"])
    assert abs(meter.avg - float(ref)) < 1e-5 * abs(float(ref))
    tr.epoch = 1
    path = tr.save_checkpoint(str(tmp_path))
    ck = torch.load(path, weights_only=False)
    assert {"ema_state_dict", "epoch", "best_valid_perf", "model_state_dict", "optimizer", "scheduler"} <= set(ck)
    tr.load_checkpoint(path, is_train=True)
    # CoMatch: bank + DA history are added to the checkpoint (absent in the reference, comatch.py:285-306)
    cm = CoMatch(ScriptedModel([], True), device="cuda")
    cm.get_dataloader((Loader([]), Loader([])), None)
    cm.get_config(_config(utils, B=4, MU=3, D=64, thr=0.9, save=str(tmp_path), train_extra={"ENQUEUE_MODE": "always", "QUEUE_SIZE": 64}),
                  optimizer=torch.optim.SGD(cm.model.parameters(), lr=0.0), lr_scheduler=NoSched())
    assert cm.queue_size == 64 and cm.head.enqueue_mode == "always"
    cm.head.queue_feats.normal_()
    cm.queue_ptr = 16
    cm.epoch = 2
    p2 = cm.save_checkpoint(str(tmp_path))
    kept = cm.head.queue_feats.clone()
    cm.head.queue_feats.zero_()
    cm.queue_ptr = 0
    cm.load_checkpoint(p2, is_train=True)
    assert torch.equal(cm.head.queue_feats, kept) and cm.queue_ptr == 16 and int(cm.head.ptr_state[0]) == 16


@pytest.mark.parametrize("opt_name", ["adam", "sgd"])
def test_trainer_with_fused_optimizer_ema_equals_the_eager_pair(opt_name):
    """TRAIN.FUSED_OPT_EMA: FixMatch.train_one with the fused optimizer + EMA launch ends with the same
    weights, EMA weights and optimizer state as with `optimizer.step(); ema_model.update(model)`."""
    import copy

    from endoscopy_image_classification_b200 import utils
    from endoscopy_image_classification_b200.fixmatch import FixMatch
    from endoscopy_image_classification_b200.optimizer import build_optimizer
    g = torch.Generator().manual_seed(11)
    B, MU, steps = 4, 2, 3
    Bu = B * MU
    lab = [(torch.randn(B, 3, 4, 4, generator=g), torch.randint(0, C, (B,), generator=g)) for _ in range(steps)]
    unl = [((torch.randn(Bu, 3, 4, 4, generator=g), torch.randn(Bu, 3, 4, 4, generator=g)), None) for _ in range(steps)]
    base = torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(48, 37), torch.nn.ReLU(), torch.nn.Linear(37, C))
    result = {}
    for fused in (False, True):
        model = copy.deepcopy(base)
        tr = FixMatch(model, device="cuda")
        tr.get_dataloader((Loader(lab), Loader(unl)), None)
        cfg = _config(utils, B=B, MU=MU, thr=0.0, ema=True, steps=steps, train_extra={"FUSED_OPT_EMA": fused})
        tr.get_config(cfg, optimizer=build_optimizer(model.cuda(), opt_name, lr=1e-2), lr_scheduler=NoSched())
        assert (tr._fused is not None) == fused
        tr.train_one(epoch=0)
        result[fused] = (copy.deepcopy(tr.model.state_dict()), copy.deepcopy(tr.ema_model.ema.state_dict()),
                         copy.deepcopy(tr.optimizer.state_dict()["state"]))
    # wiring test: the two runs back-propagate separately, so their gradients agree to rounding only (the op-by-op
    # comparison on identical gradients is tests/test_fused_step.py)
    for a, b in zip(result[False][:2], result[True][:2]):
        for k in a:
            torch.testing.assert_close(b[k], a[k], rtol=1e-5, atol=1e-6)
    moved = result[True][0]
    assert any(float((moved[k] - base.state_dict()[k].cuda()).abs().max()) > 1e-3 for k in moved)   # the steps did move the weights
    sa, sb = result[False][2], result[True][2]
    assert sa.keys() == sb.keys()
    for k in sa:
        assert sa[k].keys() == sb[k].keys()
        for n in sa[k]:
            torch.testing.assert_close(sb[k][n].float().cpu(), sa[k][n].float().cpu(), rtol=1e-4, atol=1e-7)


def test_trainer_with_overlapped_ema_equals_the_serial_one():
    """TRAIN.EMA_OVERLAP: the EMA update on a side stream (capped grid), joined ahead of the next optimizer step and before the
    EMA weights are read -- FixMatch.train_one ends with the same weights and EMA weights as the in-stream update, and
    evaluate_one / save paths see the finished update."""
    import copy

    from endoscopy_image_classification_b200 import utils
    from endoscopy_image_classification_b200.fixmatch import FixMatch
    from endoscopy_image_classification_b200.optimizer import build_optimizer
    g = torch.Generator().manual_seed(13)
    B, MU, steps = 4, 2, 5
    Bu = B * MU
    lab = [(torch.randn(B, 3, 4, 4, generator=g), torch.randint(0, C, (B,), generator=g)) for _ in range(steps)]
    unl = [((torch.randn(Bu, 3, 4, 4, generator=g), torch.randn(Bu, 3, 4, 4, generator=g)), None) for _ in range(steps)]
    base = torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(48, 37), torch.nn.ReLU(), torch.nn.Linear(37, C))
    result = {}
    for overlap in (False, True):
        torch.manual_seed(0)
        model = copy.deepcopy(base)
        tr = FixMatch(model, device="cuda")
        tr.get_dataloader((Loader(lab), Loader(unl)), None)
        cfg = _config(utils, B=B, MU=MU, thr=0.0, ema=True, steps=steps, train_extra={"EMA_OVERLAP": overlap})
        tr.get_config(cfg, optimizer=build_optimizer(model.cuda(), "sgd", lr=1e-2), lr_scheduler=NoSched())
        assert tr.ema_model.overlap == overlap
        tr.train_one(epoch=0)
        tr.ema_model.join()
        torch.cuda.synchronize()
        result[overlap] = (copy.deepcopy(tr.model.state_dict()), copy.deepcopy(tr.ema_model.ema.state_dict()))
    for a, b in zip(result[False], result[True]):
        for k in a:
            torch.testing.assert_close(b[k], a[k], rtol=1e-5, atol=1e-6)
    assert any(float((result[True][1][k] - base.state_dict()[k].cuda()).abs().max()) > 1e-5 for k in result[True][1])


def test_ddp_single_rank_nccl(tmp_path):
    """SURVEY 8 f3 on the GPU: the FixMatch trainer with its backbone inside DistributedDataParallel (a 1-rank NCCL group
    on this box; the 2-rank wiring is covered on the CPU by tests/test_ddp_trainer_gloo.py) runs the fused criteria and the
    multi-tensor EMA through DDP's backward hooks and ends with the weights of the plain single-process trainer."""
    import os
    import socket

    import torch.distributed as dist
    import torch.nn as nn
    from endoscopy_image_classification_b200 import utils
    from endoscopy_image_classification_b200.fixmatch import FixMatch

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.body = nn.Sequential(nn.Conv2d(3, 8, 3), nn.BatchNorm2d(8), nn.ReLU(), nn.AdaptiveAvgPool2d(1), nn.Flatten())
            self.fc = nn.Linear(8, C)

        def forward(self, x):
            return self.fc(self.body(x))

    def run(ddp):
        torch.manual_seed(0)
        net = Net()
        g = torch.Generator().manual_seed(1)
        B, MU = 4, 2
        lab = Loader([(torch.randn(B, 3, 8, 8, generator=g), torch.randint(0, C, (B,), generator=g)) for _ in range(3)])
        unl = Loader([((torch.randn(B * MU, 3, 8, 8, generator=g), torch.randn(B * MU, 3, 8, 8, generator=g)), None) for _ in range(3)])
        tr = FixMatch(net, opt_func="SGD", device="cuda")
        tr.get_dataloader((lab, unl), None)
        cfg = _config(utils, B=B, MU=MU, thr=0.05, steps=3, ema=True, save=str(tmp_path / "ck"), train_extra={"DDP": ddp})
        cfg.TRAIN.BASE_LR = 0.1
        tr.get_config(cfg, lr_scheduler=NoSched())
        tr.train_one(epoch=0)
        torch.cuda.synchronize()
        return tr, [p.detach().clone() for p in net.parameters()], [v.clone() for v in tr.ema_model.ema.state_dict().values()]

    _, w_plain, e_plain = run(False)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        tr, w_ddp, e_ddp = run(True)
        from torch.nn.parallel import DistributedDataParallel
        assert isinstance(tr.net, DistributedDataParallel) and tr.net.module is tr.model
        for a, b in zip(w_plain, w_ddp):                  # a 1-rank average is the identity (cuDNN's backward may reorder sums)
            torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-7)
        for a, b in zip(e_plain, e_ddp):
            torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-7)
        tr.epoch = 1
        assert tr.save_checkpoint(cfg_dir := str(tmp_path / "ck")) is not None and len(os.listdir(cfg_dir)) == 1
    finally:
        dist.destroy_process_group()


def test_fixmatch_evaluate_one_matches_reference_golden(capsys):
    """``FixMatch.evaluate_one`` of the drop-in (device evaluation head, one D2H copy) against the golden vector of the REAL
    ``FixMatch.evaluate_one`` on the same scripted validation logits: loss meter, metrics, sen/spec table."""
    from endoscopy_image_classification_b200 import utils
    from endoscopy_image_classification_b200.fixmatch import FixMatch
    z = load_golden("evaluate_one.npz")
    m = golden_meta(z)
    N, bs = m["N"], m["batch_size"]
    lg, y = T(z["logits"]), T(z["targets"])
    model = ScriptedModel([{"logits": lg[i:i + bs]} for i in range(0, N, bs)], with_feats=False)
    tr = FixMatch(model, device="cuda")
    tr.get_dataloader((Loader([]), Loader([])), Loader([(torch.zeros(len(y[i:i + bs]), 1), y[i:i + bs]) for i in range(0, N, bs)]))
    cfg = _config(utils, B=bs, MU=1, thr=0.95)
    tr.get_config(cfg, optimizer=torch.optim.SGD(model.parameters(), lr=0.0), lr_scheduler=NoSched())
    meter, metric = tr.evaluate_one(show_metric=True, show_report=True, show_cf_matrix=True)
    assert "Classification Report" in capsys.readouterr().out
    assert abs(meter.avg - float(z["loss_avg"])) < 1e-5 * float(z["loss_avg"])
    assert abs(meter.sum - float(z["loss_sum"])) < 1e-5 * float(z["loss_sum"]) and meter.count == float(z["loss_count"])
    for k, v in zip(m["metric_keys"], z["metrics"]):
        assert abs(metric[k] - v) < 1e-12, k
    assert np.allclose(metric["sen/spec"]["sensitivity"].values, z["sensitivity"], atol=1e-12)
    assert np.allclose(metric["sen/spec"]["specificity"].values, z["specificity"], atol=1e-12)
