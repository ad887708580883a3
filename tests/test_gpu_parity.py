"""GPU parity tests proper: the CUDA path (through the C ABI) against the oracle on
the same seeded inputs, against the committed golden fixtures (produced by the real
reference), and at BASELINE sizes through size-independent properties.

Tolerances (BASELINE.json north_star): argmax / mask / queue indices bit-exact
(margin-aware); losses, gradients and EMA weights <= 1e-5 relative in
fp32, <= 1e-2 with bf16 storage.
"""
import numpy as np
import pytest
import torch

from conftest import (T, assert_labels_match, assert_mask_match, golden_meta, load_golden, rel_err, top2_gap)
from oracle import ssl_oracle as O

pytestmark = pytest.mark.gpu

C = 23
FP32_TOL = 1e-5
BF16_TOL = 1e-2


@pytest.fixture(scope="module")
def pkg():
    import endoscopy_image_classification_b200 as eic
    from endoscopy_image_classification_b200 import _native, comatch_head, ema, loss
    _native.lib()
    return {"loss": loss, "ema": ema, "head": comatch_head, "native": _native, "pkg": eic}


def dev(t):
    return t.cuda()


# =============================================================== K1 =========
@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_fixmatch_head_golden(pkg, seed):
    z = load_golden(f"fixmatch_head_seed{seed}.npz")
    w, s, thr = T(z["logits_w"], "cuda"), T(z["logits_s"], "cuda").requires_grad_(True), float(z["thr"])
    loss, mask_mean, _, idx, mask = pkg["loss"].fixmatch_head(w, s, None, thr)
    (2.5 * loss).backward()
    probs_ref = torch.softmax(T(z["logits_w"]), -1)
    assert_labels_match(idx, T(z["idx"]), top2_gap(probs_ref))
    assert_mask_match(mask, T(z["mask"]), T(z["pmax"]), thr)
    assert rel_err(loss, T(z["loss"])) < FP32_TOL
    assert rel_err(mask_mean, T(z["mask_mean"])) < 1e-6
    assert rel_err(s.grad, 2.5 * T(z["grad_s"])) < FP32_TOL
    # public drop-in signature returns exactly (loss, mask.mean())
    out = pkg["loss"].consistency_loss(w, s.detach(), T=0.5, p_cutoff=thr, device="cuda")
    assert isinstance(out, tuple) and len(out) == 2
    assert torch.equal(out[0], loss.detach())      # quirk Q4: T ignored in hard-label mode; deterministic


@pytest.mark.parametrize("rows,classes", [(1, 2), (7, 23), (112, 23), (449, 23), (1000, 10), (333, 100), (65, 1000),
                                          (14336, 23)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_fixmatch_head_shapes(pkg, rows, classes, dtype):
    g = torch.Generator().manual_seed(rows * 131 + classes)
    w = (4.0 * torch.randn(rows, classes, generator=g)).to(dtype)
    s = (4.0 * torch.randn(rows, classes, generator=g)).to(dtype)
    thr = 0.6
    ref = O.fixmatch_head_details(w.float(), s.float(), thr)
    sd = dev(s).requires_grad_(True)
    loss, mask_mean, _, idx, mask = pkg["loss"].fixmatch_head(dev(w), sd, None, thr)
    loss.backward()
    assert_labels_match(idx, ref["idx"], top2_gap(torch.softmax(w.float(), -1)))
    assert_mask_match(mask, ref["mask"], ref["pmax"], thr)
    tol = FP32_TOL if dtype == torch.float32 else BF16_TOL
    if bool((mask.cpu() == ref["mask"]).all()):
        assert rel_err(loss, ref["loss"]) < tol
        assert rel_err(sd.grad.float(), ref["grad_s"]) < tol
    assert sd.grad.dtype == dtype


def test_fixmatch_head_views_unaligned_and_dual(pkg):
    """Trainer-style views (outputs[B:].chunk(2), fixmatch.py:110-112), a deliberately
    unaligned base pointer, and the SemiFormer two-strong-head variant."""
    g = torch.Generator().manual_seed(7)
    B, Bu = 16, 112
    out = dev(6.0 * torch.randn(B + 2 * Bu, C, generator=g))
    out2 = dev(6.0 * torch.randn(B + 2 * Bu, C, generator=g))
    w, s = out[B:].chunk(2)
    s2 = out2[B:].chunk(2)[1]
    la, lb, mm = pkg["loss"].consistency_loss_dual(w, s, s2, p_cutoff=0.95)
    ra = O.fixmatch_head_details(w.cpu(), s.cpu(), 0.95)
    rb = O.fixmatch_head_details(w.cpu(), s2.cpu(), 0.95)
    assert rel_err(la, ra["loss"]) < FP32_TOL and rel_err(lb, rb["loss"]) < FP32_TOL
    assert rel_err(mm, ra["mask_mean"]) < 1e-6
    # unaligned: skip one float at the front of a flat buffer
    flat = dev(6.0 * torch.randn(2 * Bu * C + 1, generator=g))
    wu = flat[1:1 + Bu * C].view(Bu, C)
    su = dev(6.0 * torch.randn(Bu * C + 3, generator=g))[3:].view(Bu, C).requires_grad_(True)
    l, _ = pkg["loss"].consistency_loss(wu, su, p_cutoff=0.7)
    l.backward()
    r = O.fixmatch_head_details(wu.cpu(), su.detach().cpu(), 0.7)
    assert rel_err(l, r["loss"]) < FP32_TOL and rel_err(su.grad, r["grad_s"]) < FP32_TOL


def test_fixmatch_head_soft_labels_and_errors(pkg):
    g = torch.Generator().manual_seed(3)
    w, s = 3.0 * torch.randn(200, C, generator=g), 3.0 * torch.randn(200, C, generator=g)
    ref = O.fixmatch_head_details(w, s, 0.5, T=0.5, use_hard_labels=False)
    sd = dev(s).requires_grad_(True)
    l, mm = pkg["loss"].consistency_loss(dev(w), sd, T=0.5, p_cutoff=0.5, use_hard_labels=False)
    l.backward()
    assert rel_err(l, ref["loss"]) < FP32_TOL and rel_err(sd.grad, ref["grad_s"]) < FP32_TOL
    # L2 branch returns a bare tensor (quirk Q8)
    l2 = pkg["loss"].consistency_loss(dev(w), dev(s), name="L2")
    assert torch.is_tensor(l2) and rel_err(l2, O.consistency_loss(w, s, name="L2")) < FP32_TOL
    with pytest.raises(AssertionError):
        pkg["loss"].consistency_loss(dev(w), dev(s), name="L2_mask")
    with pytest.raises(RuntimeError):          # no CPU path
        pkg["loss"].consistency_loss(w, s)


def test_fixmatch_property_full_size(pkg):
    """BASELINE cfg 5 top size: permuting rows permutes idx/mask/grad rows and leaves the
    loss unchanged up to summation order; mask==0 rows have exactly zero gradient."""
    g = torch.Generator().manual_seed(11)
    rows = 14336
    w, s = dev(6.0 * torch.randn(rows, C, generator=g)), dev(6.0 * torch.randn(rows, C, generator=g))
    perm = dev(torch.randperm(rows, generator=g))
    s1 = s.clone().requires_grad_(True)
    s2 = s[perm].clone().requires_grad_(True)
    l1, _, _, i1, m1 = pkg["loss"].fixmatch_head(w, s1, None, 0.95)
    l2, _, _, i2, m2 = pkg["loss"].fixmatch_head(w[perm], s2, None, 0.95)
    l1.backward()
    l2.backward()
    assert torch.equal(i1[perm], i2) and torch.equal(m1[perm], m2)
    assert torch.equal(s1.grad[perm], s2.grad)
    assert rel_err(l1, l2) < 1e-6
    assert float(s1.grad[m1 == 0].abs().max()) == 0.0
    assert torch.allclose(s1.grad.sum(1), torch.zeros(rows, device="cuda"), atol=1e-9)


# =============================================================== f2 =========
def test_labeled_ce_golden(pkg):
    z = load_golden("ce_loss.npz")
    x, y, cw = T(z["logits"], "cuda"), T(z["targets"], "cuda"), T(z["class_weights"], "cuda")
    for nm, kw in {"plain_mean_w": dict(class_weights=cw), "poly_mean_w": dict(class_weights=cw, type_loss="poly"),
                   "poly_mean": dict(type_loss="poly")}.items():
        xd = x.clone().requires_grad_(True)
        l = pkg["loss"].ce_loss(xd, y, reduction="mean", **kw)
        l.backward()
        assert rel_err(l, T(z[nm])) < FP32_TOL, nm
        assert rel_err(xd.grad, T(z[nm + "_grad"])) < FP32_TOL, nm
    xd = x.clone().requires_grad_(True)
    l = pkg["loss"].ce_loss(xd, y, reduction="mean")
    l.backward()
    xr = T(z["logits"]).requires_grad_(True)
    r = O.ce_loss(xr, T(z["targets"]), reduction="mean")
    r.backward()
    assert rel_err(l, r) < FP32_TOL and rel_err(xd.grad, xr.grad) < FP32_TOL
    with pytest.raises(NotImplementedError):
        pkg["loss"].ce_loss(x, y, reduction="mean", type_loss="focal")
    # the fused criterion scales its stashed gradient in place: a second backward must raise, not double-scale
    xd = x.clone().requires_grad_(True)
    l = pkg["loss"].ce_loss(xd, y, reduction="mean")
    l.backward(retain_graph=True)
    with pytest.raises(RuntimeError):
        l.backward()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("rows,classes", [(16, 23), (1, 2), (513, 23), (64, 100)])
def test_ce_loss_unreduced_sum_soft_and_ignore(pkg, rows, classes, dtype):
    """loss.py:118-124 with the reference's default reduction='none', 'sum', PolyLoss reduction='none', soft targets
    (use_hard_labels=False), F.cross_entropy's ignore_index and the out-of-range label guard -- against the oracle."""
    loss = pkg["loss"]
    tol = FP32_TOL if dtype == torch.float32 else BF16_TOL
    g = torch.Generator().manual_seed(rows * 7 + classes)
    x = (3 * torch.randn(rows, classes, generator=g)).to(dtype)
    y = torch.randint(0, classes, (rows,), generator=g)
    cw = torch.rand(classes, generator=g) + 0.5
    up = torch.rand(rows, generator=g) + 0.5                       # per-row upstream gradient
    for kw in (dict(), dict(class_weights=cw), dict(type_loss="poly"), dict(type_loss="poly", class_weights=cw)):
        for red in ("none", "sum"):
            xr = x.float().clone().requires_grad_(True)
            r = O.ce_loss(xr, y, reduction=red, **kw)
            (r * up).sum().backward() if red == "none" else r.backward()
            xd = x.clone().cuda().requires_grad_(True)
            kwd = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in kw.items()}
            l = loss.ce_loss(xd, y.cuda(), reduction=red, **kwd)
            assert l.shape == r.shape
            (l.float() * up.cuda()).sum().backward(retain_graph=True) if red == "none" else l.backward(retain_graph=True)
            assert rel_err(l.float(), r) < tol and rel_err(xd.grad.float(), xr.grad) < tol, (kw.keys(), red)
            first = xd.grad.clone()                                # out-of-place chaining: a second backward is exact
            xd.grad = None
            (l.float() * up.cuda()).sum().backward() if red == "none" else l.backward()
            assert torch.equal(first, xd.grad)
    # soft targets (loss.py:120-124; always un-reduced, class weights unused)
    t = torch.softmax(torch.randn(rows, classes, generator=g), 1)
    xr = x.float().clone().requires_grad_(True)
    r = O.ce_loss(xr, t, use_hard_labels=False)
    (r * up).sum().backward()
    xd = x.clone().cuda().requires_grad_(True)
    l = loss.ce_loss(xd, t.cuda(), use_hard_labels=False)
    (l.float() * up.cuda()).sum().backward()
    assert rel_err(l.float(), r) < tol and rel_err(xd.grad.float(), xr.grad) < tol
    # ignore_index rows and the out-of-range guard
    if rows >= 4 and dtype == torch.float32:
        yi = y.clone()
        yi[1] = -100
        for red in ("none", "mean"):
            for w in (None, cw):
                xr = x.float().clone().requires_grad_(True)
                r = torch.nn.functional.cross_entropy(xr, yi, weight=w, reduction=red)
                r.sum().backward()
                xd = x.clone().cuda().requires_grad_(True)
                l = loss.ce_loss(xd, yi.cuda(), class_weights=None if w is None else w.cuda(), reduction=red)
                l.sum().backward()
                assert rel_err(l, r) < tol and rel_err(xd.grad, xr.grad) < tol
        assert loss.bad_label_count() == 0
        yb = y.clone()
        yb[0], yb[2] = classes, -7
        l = loss.ce_loss(x.cuda(), yb.cuda(), reduction="none")
        assert float(l[0]) == 0.0 and float(l[2]) == 0.0 and bool(torch.isfinite(l).all())
        loss.ce_loss(x.cuda(), yb.cuda(), reduction="mean", type_loss="poly")
        assert loss.bad_label_count() == 4 and loss.bad_label_count() == 0


# =============================================================== K8 =========
def _tinynet(alias):
    import torch.nn as nn

    class _TinyNet(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv = nn.Conv2d(3, 8, 3, bias=False)
            self.bn = nn.BatchNorm2d(8)
            self.fc = nn.Linear(8, C)
            self.model = nn.Sequential(self.conv, self.bn)
            if alias:
                self.backbone = nn.Sequential(self.conv, self.bn)
    return _TinyNet()


@pytest.mark.parametrize("alias", [0, 1])
def test_ema_golden_bit_exact(pkg, alias):
    z = load_golden(f"ema_tinynet_alias{alias}.npz")
    keys = [str(k) for k in z["keys"]]
    model = _tinynet(alias).cuda()
    ema = pkg["ema"].ModelEMA(model, decay=float(z["decay"]), device="cuda")
    with torch.no_grad():
        for k, v in ema.ema.state_dict().items():
            v.copy_(T(z[f"e0/{k}"], "cuda"))
    for step in range(3):
        with torch.no_grad():
            for k, v in model.state_dict().items():
                v.copy_(T(z[f"m{step}/{k}"], "cuda"))
        ema.update(model)
    sd = ema.ema.state_dict()
    assert list(sd.keys()) == keys
    for k in keys:
        assert torch.equal(sd[k].cpu(), T(z[f"e3/{k}"])), k
    assert ema.plan.n_unique < ema.plan.n_entries          # aliased names de-duplicated (Q2)
    ema.set(model)
    for k, v in model.state_dict().items():
        assert torch.equal(ema.ema.state_dict()[k], v)


@pytest.mark.parametrize("arch,dtype", [("resnet18", torch.float32), ("resnet50", torch.float32),
                                        ("resnet18", torch.bfloat16)])
def test_ema_torchvision_bit_exact(pkg, arch, dtype):
    """Random-init torchvision backbones: 3 updates bit-identical to the reference loop
    (ema.py:51-56) executed on the CPU by the oracle, buffers and int64 counters included."""
    import torchvision
    torch.manual_seed(0)
    model = getattr(torchvision.models, arch)(num_classes=C).to(dtype)
    ref_e = {k: v.clone() for k, v in model.state_dict().items()}
    gm = model.cuda()
    ema = pkg["ema"].ModelEMA(gm, decay=0.999, device="cuda")
    g = torch.Generator().manual_seed(1)
    for step in range(3):
        with torch.no_grad():
            for k, v in gm.state_dict().items():
                if v.is_floating_point():
                    v.add_((1e-3 * torch.randn(v.shape, generator=g)).to(dtype).cuda())
                else:
                    v.add_(5)
        ema.update(gm)
        O.ema_update_(list(ref_e.values()), [v.cpu() for v in gm.state_dict().values()], 0.999)
    for k, v in ema.ema.state_dict().items():
        assert torch.equal(v.cpu(), ref_e[k]), k
    assert ema.plan.unique_elems == sum(v.numel() for v in ref_e.values())


@pytest.mark.parametrize("mode", ["capped", "masked"])
def test_ema_overlapped_forms_bit_exact(pkg, mode):
    """ModelEMA(overlap=True): the update on a side stream -- with a capped grid behind a short head start, or with a probed
    set of SMs left alone and chunks fetched from a device-side scheduler -- gives the bits of the in-stream update, launch
    after launch (the scheduler re-arms itself) and replayed from a CUDA graph."""
    import torchvision
    torch.manual_seed(0)
    model = torchvision.models.resnet18(num_classes=C).cuda()
    ref = pkg["ema"].ModelEMA(model, decay=0.999, device="cuda")
    ema = pkg["ema"].ModelEMA(model, decay=0.999, device="cuda", overlap=True)
    ema.overlap_mode = mode
    g = torch.Generator(device="cuda").manual_seed(1)

    def perturb():
        with torch.no_grad():
            for v in model.state_dict().values():
                if v.is_floating_point():
                    v.add_(1e-3 * torch.randn(v.shape, generator=g, device="cuda"))
                else:
                    v.add_(3)

    def same():
        torch.cuda.synchronize()
        return all(torch.equal(a, b) for a, b in zip(ema.ema.state_dict().values(), ref.ema.state_dict().values()))

    for _ in range(3):
        perturb()
        ref.update(model)
        ema.update(model)
        ema.join()
    assert (ema._masked is not None) == (mode == "masked")
    assert same()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):                           # (the capture does not execute)
        ema.update(model)
        ema.join()
    for _ in range(3):
        perturb()
        ref.update(model)
        graph.replay()
    assert same()


def test_ema_rejects_cpu_and_detects_realloc(pkg):
    import torch.nn as nn
    m = nn.Linear(8, 8)
    ema = pkg["ema"].ModelEMA(m, decay=0.9)
    with pytest.raises(RuntimeError):
        ema.update(m)
    m = nn.Linear(64, 64).cuda()
    ema = pkg["ema"].ModelEMA(m, decay=0.5, device="cuda")
    ema.update(m)
    with torch.no_grad():
        m.weight.data = torch.ones(64, 64, device="cuda")     # re-allocated parameter
        m.bias.data = torch.ones(64, device="cuda")
    before = ema.ema.weight.clone()
    ema.update(m)
    assert torch.equal(ema.ema.weight, 0.5 * before + 0.5)


# ========================================================= CoMatch head =======
def _split_step(z, i, B, Bu, device="cuda", dtype=None):
    lg, ft, tx = T(z[f"s{i}/logits"], device, dtype), T(z[f"s{i}/feats"], device, dtype), T(z[f"s{i}/targets_x"], device)
    luw, lus0, _ = torch.split(lg[B:], Bu)
    fuw, fus0, fus1 = torch.split(ft[B:], Bu)
    return lg, ft, tx, luw, lus0, fuw, fus0, fus1


@pytest.mark.parametrize("qb", [1, 5])
def test_comatch_head_golden_reference_train_one(pkg, qb):
    """Replays the inputs of the REAL reference CoMatch.train_one (4 steps) and checks
    losses, gradients, masks, pseudo labels, bank rows and pointer."""
    z = load_golden(f"comatch_train_one_qb{qb}.npz")
    m = golden_meta(z)
    B, MU, D = m["B"], m["MU"], m["D"]
    Bu = B * MU
    head = pkg["head"].CoMatchHead(C, D, qb * (MU + 1) * B, m["thr"], enqueue_mode="reference")
    for i in range(m["nsteps"]):
        lg, ft, tx, luw, lus0, fuw, fus0, fus1 = _split_step(z, i, B, Bu)
        lus0 = lus0.clone().requires_grad_(True)
        fus0 = fus0.clone().requires_grad_(True)
        fus1 = fus1.clone().requires_grad_(True)
        loss_u, loss_c, mask_mean, mask, lbs, scores, probs = head(luw, lus0, fuw, fus0, fus1, ft[:B], tx)
        (loss_u + loss_c).backward()
        pr = T(z[f"s{i}/probs"])
        assert rel_err(probs, pr) < FP32_TOL
        assert_labels_match(lbs, T(z[f"s{i}/lbs"]), top2_gap(pr), "lbs_u_guess")
        assert_mask_match(mask, T(z[f"s{i}/mask"]), pr.max(1).values, m["thr"])
        assert rel_err(loss_u, T(z[f"s{i}/loss_u"])) < FP32_TOL or float(z[f"s{i}/loss_u"]) == 0.0
        assert rel_err(loss_c, T(z[f"s{i}/loss_contrast"])) < FP32_TOL
        g, gf = T(z[f"s{i}/grad_logits"]), T(z[f"s{i}/grad_feats"])
        if float(g.abs().max()) > 0:
            assert rel_err(lus0.grad, g[B + Bu:B + 2 * Bu]) < FP32_TOL
        assert rel_err(fus0.grad, gf[B + Bu:B + 2 * Bu]) < FP32_TOL
        assert rel_err(fus1.grad, gf[B + 2 * Bu:]) < FP32_TOL
        # bank: indices bit-exact, feature rows are copies (bit-exact), prob rows to 1e-5
        assert head.queue_ptr == int(z[f"s{i}/queue_ptr"])
        assert torch.equal(head.queue_feats.cpu(), T(z[f"s{i}/queue_feats"]))
        qp = T(z[f"s{i}/queue_probs"])
        assert torch.equal((head.queue_probs.cpu() != 0).any(1), (qp != 0).any(1))
        assert rel_err(head.queue_probs, qp) < FP32_TOL or float(qp.abs().max()) == 0.0
    hist = head.prob_list
    assert len(hist) == m["nsteps"]


def _clustered(g, B, Bu, D, protos, noise=0.075):
    def feats(y):
        x = protos[y] + noise * torch.randn(len(y), D, generator=g)
        return x / x.norm(dim=1, keepdim=True)

    def logit(y, scale):
        return scale * torch.nn.functional.one_hot(y, C).float() + 2.0 * torch.randn(len(y), C, generator=g)
    y_u, y_x = torch.randint(0, C, (Bu,), generator=g), torch.randint(0, C, (B,), generator=g)
    return dict(logits_u_w=logit(y_u, 5.0), logits_u_s0=logit(y_u, 4.0), feats_u_w=feats(y_u), feats_u_s0=feats(y_u),
                feats_u_s1=feats(y_u), feats_x=feats(y_x), targets_x=y_x)


@pytest.mark.parametrize("fuse_rows", [True, False])
@pytest.mark.parametrize("B,MU,D,qbatches,dtype", [(8, 7, 64, 5, torch.float32), (5, 3, 32, 3, torch.float32),
                                                   (64, 7, 64, 5, torch.float32), (64, 7, 64, 5, torch.bfloat16),
                                                   (3, 5, 128, 4, torch.float32), (160, 7, 64, 2, torch.bfloat16)])
def test_comatch_head_always_mode_vs_oracle(pkg, B, MU, D, qbatches, dtype, fuse_rows):
    """'always' enqueue (upstream semantics) over enough steps for the ring to wrap;
    ragged sizes (rows not a multiple of 64, K not a multiple of the tile)."""
    g = torch.Generator().manual_seed(B * 1000 + D)
    Bu, thr = B * MU, 0.9
    n = B + Bu
    K = qbatches * n
    protos = torch.randn(C, D, generator=g)
    protos = protos / protos.norm(dim=1, keepdim=True)
    head = pkg["head"].CoMatchHead(C, D, K, thr, enqueue_mode="always", dtype=dtype)
    head.fuse_rows = fuse_rows          # one cluster launch for DA + finalize + enqueue, or the three kernels
    state = O.CoMatchState.zeros(K, D, C)
    tol = FP32_TOL if dtype == torch.float32 else BF16_TOL
    fired = 0
    for step in range(qbatches + 2):
        inp = _clustered(g, B, Bu, D, protos)
        inp = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in inp.items()}
        ref_in = {k: (v.float() if v.is_floating_point() else v) for k, v in inp.items()}
        if dtype != torch.float32:          # the oracle's bank must hold what a bf16 bank can hold
            state.queue_probs = state.queue_probs.to(dtype).float()
        ref = O.comatch_head(state, **ref_in, thr=thr, num_classes=C, enqueue_mode="always")
        din = {k: v.cuda() for k, v in inp.items()}
        for k in ("logits_u_s0", "feats_u_s0", "feats_u_s1"):
            din[k].requires_grad_(True)
        if step % 2 == 0:
            loss_u, loss_c, mask_mean, mask, lbs, scores, probs = head(**din)
            (loss_u + loss_c).backward()
        else:                                   # fused weighting (comatch.py:222), undone below
            total, loss_u, loss_c, mask_mean = head.total_loss(**din, lambda_u=2.0, lambda_c=0.5)
            total.backward()
            assert rel_err(total, 2.0 * loss_u + 0.5 * loss_c) < 1e-6
            din["logits_u_s0"].grad.mul_(0.5)
            din["feats_u_s0"].grad.mul_(2.0)
            din["feats_u_s1"].grad.mul_(2.0)
            probs, mask, lbs = head.last["probs"], head.last["mask"], head.last["lbs"]
        assert head.queue_ptr == state.queue_ptr
        assert torch.equal(head.queue_feats.float().cpu(), state.queue_feats)
        assert rel_err(head.queue_probs.float(), state.queue_probs) < tol
        assert rel_err(probs, ref["probs"]) < tol
        # bf16 storage: the inputs are rounded identically on both sides; what differs is K3's bf16 P (measured 3e-3 on the
        # smoothing sums, times 1 - alpha = 0.1) -- a label / mask may flip only where the reference grazes within 4e-3
        assert_labels_match(lbs, ref["lbs"], top2_gap(ref["probs"]), "lbs", tol=1e-6 if dtype == torch.float32 else 4e-3)
        assert_mask_match(mask, ref["mask"], ref["scores"], thr, tol=1e-6 if dtype == torch.float32 else 4e-3)
        assert rel_err(loss_c, ref["loss_contrast"]) < tol
        assert rel_err(din["feats_u_s0"].grad.float(), ref["grad_feats_s0"]) < tol
        assert rel_err(din["feats_u_s1"].grad.float(), ref["grad_feats_s1"]) < tol
        if bool((mask.cpu() == ref["mask"]).all()) and float(ref["mask"].sum()) > 0:
            fired += 1
            assert rel_err(loss_u, ref["loss_u"]) < tol
            assert rel_err(din["logits_u_s0"].grad.float(), ref["grad_logits_s0"]) < tol
    assert fired > 0, "mask never fired: the focal branch was not exercised"


def test_comatch_head_degenerate_zero_bank_and_no_smoothing(pkg):
    g = torch.Generator().manual_seed(5)
    B, Bu, D = 4, 28, 64
    protos = torch.nn.functional.normalize(torch.randn(C, D, generator=g), dim=1)
    inp = _clustered(g, B, Bu, D, protos)
    for smoothing in (True, False):
        head = pkg["head"].CoMatchHead(C, D, 5 * (B + Bu), 0.6, smoothing=smoothing)      # reference mode: never enqueues (Q1)
        st = O.CoMatchState.zeros(5 * (B + Bu), D, C)
        ref = O.comatch_head(st, **inp, thr=0.6, num_classes=C, smoothing=smoothing)
        out = head(**{k: v.cuda() for k, v in inp.items()})
        assert head.queue_ptr == 0 and float(head.queue_feats.abs().max()) == 0.0
        assert rel_err(out[6], ref["probs"]) < FP32_TOL
        assert rel_err(out[0], ref["loss_u"]) < FP32_TOL and rel_err(out[1], ref["loss_contrast"]) < FP32_TOL


@pytest.mark.parametrize("rows,K,D", [(448, 2560, 64), (100, 77, 8), (1, 1, 16), (130, 4097, 64), (448, 65536, 64)])
def test_bank_smooth_partial_vs_oracle(pkg, rows, K, D):
    g = torch.Generator().manual_seed(rows + K)
    nf = lambda n: torch.nn.functional.normalize(torch.randn(n, D, generator=g), dim=1)
    f, qf, qp = nf(rows), nf(K), torch.softmax(torch.randn(K, C, generator=g), 1)
    head = pkg["head"].CoMatchHead(C, D, K, 0.9, enqueue_mode="always")
    head.queue_feats.copy_(qf)
    head.queue_probs.copy_(qp)
    rowsum, numer = head._k_smooth(f.cuda())
    A = torch.exp(f.double() @ qf.double().t() / 0.2)
    assert rel_err(rowsum, A.sum(1)) < FP32_TOL
    assert rel_err(numer, A @ qp.double()) < FP32_TOL


@pytest.mark.parametrize("simt", [0, 1])
@pytest.mark.parametrize("rows,K", [(448, 2560), (300, 20003), (1000, 65536)])
def test_bank_smooth_fp32_storage_both_paths(pkg, rows, K, simt):
    """fp32 storage (the reference's precision, comatch.py:180-181): the tensor-core kernel on bf16 hi + mid operands
    (csrc/bank_tc.cu, default) and the exact-fp32 FFMA tiles (csrc/bank.cu) both meet the 1e-5 bar against fp64 math --
    ragged K (not a multiple of the key tile or of 8), a long bank (the per-unit accumulator flush), clustered embeddings."""
    from endoscopy_image_classification_b200 import _native as N
    g = torch.Generator().manual_seed(rows + K)
    protos = torch.nn.functional.normalize(torch.randn(C, 64, generator=g), dim=1)
    lab = torch.randint(0, C, (K,), generator=g)
    qf = torch.nn.functional.normalize(protos[lab] + 0.35 * torch.randn(K, 64, generator=g), dim=1)
    f = torch.nn.functional.normalize(protos[torch.randint(0, C, (rows,), generator=g)] + 0.35 * torch.randn(rows, 64, generator=g), dim=1)
    qp = torch.softmax(3.0 * torch.randn(K, C, generator=g) + 4.0 * torch.nn.functional.one_hot(lab, C), 1)
    head = pkg["head"].CoMatchHead(C, 64, K, 0.9, enqueue_mode="always")
    head.queue_feats.copy_(qf)
    head.queue_probs.copy_(qp)
    N.lib().b200ssl_debug_set_k3_f32_simt(simt)
    try:
        rowsum, numer = head._k_smooth(f.cuda())
        torch.cuda.synchronize()
    finally:
        N.lib().b200ssl_debug_set_k3_f32_simt(0)
    A = torch.exp(f.double() @ qf.double().t() / 0.2)
    assert rel_err(rowsum, A.sum(1)) < FP32_TOL
    assert rel_err(numer, A @ qp.double()) < FP32_TOL
    assert rel_err(numer.double().cpu() / rowsum.double().cpu().unsqueeze(1), (A @ qp.double()) / A.sum(1, keepdim=True)) < FP32_TOL


@pytest.mark.parametrize("rows,D,simt", [(448, 64, 0), (448, 64, 1), (1, 8, 0), (63, 16, 0), (65, 64, 0), (65, 64, 1), (1000, 128, 0),
                                         (3584, 64, 0), (3584, 64, 1)])
def test_contrast_fwd_bwd_vs_oracle(pkg, rows, D, simt):
    """fp32 storage against the fp64 oracle at 1e-5.  64-wide embeddings run the tensor-core kernels on bf16 hi + mid operands
    (``simt=0``, default) or the exact-fp32 FFMA tiles (``simt=1``); other widths always the FFMA tiles."""
    from endoscopy_image_classification_b200 import _native as N
    N.lib().b200ssl_debug_set_k3_f32_simt(simt)
    try:
        _contrast_fwd_bwd_vs_oracle(pkg, rows, D)
    finally:
        N.lib().b200ssl_debug_set_k3_f32_simt(0)


def _contrast_fwd_bwd_vs_oracle(pkg, rows, D):
    g = torch.Generator().manual_seed(rows * 7 + D)
    nf = lambda: torch.nn.functional.normalize(torch.randn(rows, D, generator=g), dim=1)
    y = torch.randint(0, C, (rows,), generator=g)
    probs = torch.softmax(6.0 * torch.nn.functional.one_hot(y, C).float() + torch.randn(rows, C, generator=g), 1)
    f0, f1 = nf().requires_grad_(True), nf().requires_grad_(True)
    ref = O.comatch_contrast(f0.double(), f1.double(), probs.double(), 0.2, 0.8)
    ref.backward()
    head = pkg["head"].CoMatchHead(C, D, 64, 0.9)
    scal = torch.zeros(4, device="cuda")
    d0, d1, dp = f0.detach().cuda(), f1.detach().cuda(), probs.cuda()
    stats, _ = head._k_contrast_fwd(d0, d1, dp, scal)
    up = torch.tensor(1.5, device="cuda")
    g0, g1 = head._k_contrast_bwd(d0, d1, dp, stats, up)
    torch.cuda.synchronize()
    assert abs(float(scal[2]) - float(ref)) < FP32_TOL * max(abs(float(ref)), 1e-2)   # rows=1: loss ~ -1e-7
    assert rel_err(g0, 1.5 * f0.grad) < FP32_TOL and rel_err(g1, 1.5 * f1.grad) < FP32_TOL


def test_enqueue_kernel_matches_segment_plan(pkg):
    """Sharded ring write: each emulated rank writes exactly the rows bank.local_segments
    predicts; together they equal the single-bank oracle (rank-major, wrapping)."""
    from endoscopy_image_classification_b200.bank import ShardGeometry, local_segments
    g = torch.Generator().manual_seed(2)
    K, D, R, n_u, n_x, ptr = 96, 16, 4, 14, 2, 80
    n = n_u + n_x
    blocks = [dict(fu=torch.randn(n_u, D, generator=g), fx=torch.randn(n_x, D, generator=g),
                   po=torch.rand(n_u, C, generator=g), tx=torch.randint(0, C, (n_x,), generator=g)) for _ in range(R)]
    st = O.CoMatchState.zeros(K, D, C)
    st.queue_ptr = ptr
    for b in blocks:
        onehot = torch.nn.functional.one_hot(b["tx"], C).float()
        O.bank_enqueue(st, torch.cat([b["fu"], b["fx"]]), torch.cat([b["po"], onehot]), "always")
    lib, N = pkg["native"].lib(), pkg["native"]
    full_f = torch.zeros(K, D)
    for rank in range(R):
        geom = ShardGeometry(K, R, rank)
        qf = torch.zeros(geom.shard_rows, D, device="cuda")
        qp = torch.zeros(geom.shard_rows, C, device="cuda")
        for r, b in enumerate(blocks):
            fu, fx, po, tx = (b[k].cuda() for k in ("fu", "fx", "po", "tx"))
            N.check(lib.b200ssl_bank_enqueue(qf.data_ptr(), qp.data_ptr(), None, fu.data_ptr(), fx.data_ptr(), po.data_ptr(),
                                             tx.data_ptr(), n_u, n_x, D, C, N.F32, ptr, None, 0, r * n, K, geom.shard_begin,
                                             geom.shard_rows, N.stream_ptr(qf.device)))
        lo = geom.shard_begin
        assert torch.equal(qf.cpu(), st.queue_feats[lo:lo + geom.shard_rows])
        assert torch.equal(qp.cpu(), st.queue_probs[lo:lo + geom.shard_rows])
        touched = torch.zeros(geom.shard_rows, dtype=torch.bool)
        for src, dst, ln in local_segments(ptr, R * n, geom):
            touched[dst:dst + ln] = True
        assert torch.equal(touched, (qf.cpu() != 0).any(1))


@pytest.mark.parametrize("overlap", [False, True])
def test_graphed_step_matches_eager(pkg, overlap):
    """CUDA-graph replay of head fwd+bwd (+EMA) == the eager calls: same losses, grads, bank,
    device write pointer and DA history after several steps.  ``overlap``: the EMA update as a parallel branch of the
    step graph (``ModelEMA(overlap=True)``: side stream forked at the start of the step, joined at its end, the head
    captured on a high-priority stream) against the serial eager step."""
    from endoscopy_image_classification_b200.graphs import GraphedStep
    g = torch.Generator().manual_seed(21)
    B, MU, D, thr = 8, 7, 64, 0.9
    Bu, n = B * MU, B * (MU + 1)
    protos = torch.nn.functional.normalize(torch.randn(C, D, generator=g), dim=1)
    batches = [_clustered(g, B, Bu, D, protos) for _ in range(5)]
    keys = list(batches[0].keys())
    net = torch.nn.Linear(32, 32).cuda()

    def make(ov=False):
        head = pkg["head"].CoMatchHead(C, D, 3 * n, thr, enqueue_mode="always")
        ema = pkg["ema"].ModelEMA(net, decay=0.9, device="cuda", overlap=ov)
        return head, ema

    def run(head, ema, batch):
        if ema.overlap:
            ema.update(net)
        for k in ("logits_u_s0", "feats_u_s0", "feats_u_s1"):
            batch[k].grad = None
            batch[k].requires_grad_(True)
        total, lu, lc, mm = head.total_loss(**batch, lambda_u=2.0, lambda_c=2.0)
        total.backward()
        if ema.overlap:
            ema.join()
        else:
            ema.update(net)
        return total

    head_e, ema_e = make()
    eager = []
    for b in batches:
        db = {k: v.cuda() for k, v in b.items()}
        t = run(head_e, ema_e, db)
        eager.append((float(t), db["feats_u_s0"].grad.clone(), db["logits_u_s0"].grad.clone()))

    head_g, ema_g = make(overlap)
    # capture runs 3 warm-up steps + captures on batch 0; rewind the state afterwards
    gs = GraphedStep(lambda sb: run(head_g, ema_g, sb), {k: v.cuda() for k, v in batches[0].items()}, "cuda", warmup=3,
                     on_replay=lambda: head_g.note_graph_replay(n), after_capture=lambda: head_g.sync_ptr_from_device(),
                     high_priority=overlap)
    fresh, ema_f = make()
    head_g.load_state_dict(fresh.state_dict())
    head_g.queue_ptr = 0
    with torch.no_grad():
        for a, b in zip(ema_g.ema.state_dict().values(), ema_f.ema.state_dict().values()):
            a.copy_(b)
    for i, b in enumerate(batches):
        if i % 2 == 0:
            for k in keys:
                gs.static[k].detach().copy_(b[k].cuda())
            val, grads = float(gs.replay()), gs.grads
        else:
            val, grads = gs.replay_host(b), gs.grads_host
        assert val == eager[i][0]
        assert torch.equal(grads["feats_u_s0"], eager[i][1])
        assert torch.equal(grads["logits_u_s0"], eager[i][2])
    torch.cuda.synchronize()
    assert head_g.queue_ptr == head_e.queue_ptr == int(head_g.ptr_state[0]) == int(head_e.ptr_state[0])
    assert torch.equal(head_g.queue_feats, head_e.queue_feats) and torch.equal(head_g.queue_probs, head_e.queue_probs)
    assert torch.equal(head_g.da_ring, head_e.da_ring) and torch.equal(head_g.da_state, head_e.da_state)
    for a, b in zip(ema_g.ema.state_dict().values(), ema_e.ema.state_dict().values()):
        assert torch.equal(a, b)


@pytest.mark.parametrize("rows,K", [(448, 2560), (100, 80), (1, 8), (130, 4104), (14336, 2560), (448, 65536), (3000, 16384)])
def test_bank_smooth_tcgen05_bf16(pkg, rows, K):
    """bf16 bank, D=64: the tcgen05/TMEM/TMA kernel (csrc/bank_tc.cu) against fp64 math on the same
    bf16-rounded operands; includes ragged rows/keys, one-tile CTAs and multi-tile pipelines."""
    D = 64
    g = torch.Generator().manual_seed(rows * 3 + K)
    nf = lambda n: torch.nn.functional.normalize(torch.randn(n, D, generator=g), dim=1).to(torch.bfloat16)
    f, qf = nf(rows), nf(K)
    qp = torch.softmax(2.0 * torch.randn(K, C, generator=g), 1).to(torch.bfloat16)
    head = pkg["head"].CoMatchHead(C, D, K, 0.9, enqueue_mode="always", dtype=torch.bfloat16)
    assert head.queue_probs_t is not None, "tensor-core bank layout not allocated"
    head.queue_feats.copy_(qf)
    head.queue_probs.copy_(qp)
    head.queue_probs_t[:C].copy_(qp.t())
    rowsum, numer = head._k_smooth(f.cuda())
    torch.cuda.synchronize()
    A = torch.exp(f.double() @ qf.double().t() / 0.2)
    assert rel_err(rowsum, A.sum(1)) < 3e-3
    assert rel_err(numer, A @ qp.double()) < 3e-3
    # and the smoothed probabilities that the head would form from them
    sm = (numer.double().cpu() / rowsum.double().cpu()[:, None])
    ref = (A @ qp.double()) / A.sum(1, keepdim=True)
    assert float((sm - ref).abs().max()) < 2e-3


@pytest.mark.parametrize("rows", [448, 1, 100, 130, 1000, 3584])
def test_contrast_tcgen05_bf16(pkg, rows):
    """bf16 embeddings, D=64: graph-contrastive fwd + bwd on tcgen05 (csrc/contrast_tc.cu) against
    fp64 autograd on the same bf16-rounded embeddings and fp32 probabilities."""
    D = 64
    g = torch.Generator().manual_seed(rows * 11 + 5)
    nf = lambda: torch.nn.functional.normalize(torch.randn(rows, D, generator=g), dim=1).to(torch.bfloat16)
    y = torch.randint(0, C, (rows,), generator=g)
    probs = torch.softmax(6.0 * torch.nn.functional.one_hot(y, C).float() + torch.randn(rows, C, generator=g), 1)
    if rows > 2000:
        # 12.8 M pairs: keep every Q = <p_i, p_j> at least 1e-2 away from the 0.8 graph threshold, otherwise a few
        # dozen pairs graze it within the ~6e-6 accuracy of the bf16 hi/lo split and flip (not a numerics bug)
        conf = 0.95 + 0.01 * (torch.rand(rows, 1, generator=g) - 0.5)
        probs = (1 - conf) / (C - 1) * torch.ones(rows, C) + (conf - (1 - conf) / (C - 1)) * torch.nn.functional.one_hot(y, C).float()
    f0b, f1b = nf(), nf()
    f0, f1 = f0b.double().requires_grad_(True), f1b.double().requires_grad_(True)
    ref = O.comatch_contrast(f0, f1, probs.double(), 0.2, 0.8)
    ref.backward()
    hi = probs.to(torch.bfloat16)
    lo = (probs - hi.float()).to(torch.bfloat16)
    hl = torch.zeros(rows, 64, dtype=torch.bfloat16)
    hl[:, :C], hl[:, 32:32 + C] = hi, lo
    head = pkg["head"].CoMatchHead(C, D, 64, 0.9, dtype=torch.bfloat16)
    scal = torch.zeros(4, device="cuda")
    d0, d1, dp, dhl = f0b.cuda(), f1b.cuda(), probs.cuda(), hl.cuda()
    stats, _ = head._k_contrast_fwd(d0, d1, dp, scal, probs_hl=dhl)
    up = torch.tensor(1.5, device="cuda")
    g0, g1 = head._k_contrast_bwd(d0, d1, dp, stats, up, 2.0, probs_hl=dhl)
    torch.cuda.synchronize()
    assert abs(float(scal[2]) - float(ref)) < 2e-3 * max(abs(float(ref)), 1e-2)
    A = torch.exp(f0b.double() @ f1b.double().t() / 0.2)
    assert rel_err(stats[0], A.sum(1)) < 1e-3
    if rows == 1:
        # one sample: P = 1, dZ = P (G - r) is identically zero in exact arithmetic.  The forward closes r in the moments of its
        # first pass, the backward evaluates G per pair: they cancel to rounding, not to the bit
        assert float(f0.grad.abs().max()) == 0.0 and float(g0.float().abs().max()) < 1e-6 and float(g1.float().abs().max()) < 1e-6
        return
    assert rel_err(g0.float(), 3.0 * f0.grad) < BF16_TOL and rel_err(g1.float(), 3.0 * f1.grad) < BF16_TOL


def test_tensor_core_path_from_a_fresh_thread(pkg):
    """autograd runs backward on its own worker thread, where a launch of ours can be the first CUDA call: the
    tensor-map encode (driver API) must bind the context itself (regression: CUresult 201 in bench.py)."""
    import threading
    head = pkg["head"].CoMatchHead(C, 64, 256, 0.9, enqueue_mode="always", dtype=torch.bfloat16)
    f = torch.nn.functional.normalize(torch.randn(64, 64), dim=1).to(torch.bfloat16).cuda()
    err = []

    def work():
        try:
            with torch.cuda.device(0):
                rowsum, _ = head._k_smooth(f)
                torch.cuda.synchronize()
                assert bool(torch.isfinite(rowsum).all())
        except Exception as e:  # pragma: no cover
            err.append(e)
    t = threading.Thread(target=work)
    t.start()
    t.join()
    assert not err, err
