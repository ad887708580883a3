"""CPU: the oracle restatement reproduces every committed golden vector (which were
produced by the real reference, see oracle/pin_against_reference.py)."""
import hashlib
import json

import numpy as np
import pytest
import torch

from conftest import GOLDEN, T, golden_meta, load_golden
from oracle import ssl_oracle as O


def test_manifest_hashes():
    man = json.loads((GOLDEN / "MANIFEST.json").read_text())
    for name, want in man["sha256"].items():
        z = load_golden(name)
        h = hashlib.sha256()
        for k in sorted(z):
            h.update(k.encode())
            h.update(np.ascontiguousarray(z[k]).tobytes())
        assert h.hexdigest() == want, name


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_fixmatch_head_golden(seed):
    z = load_golden(f"fixmatch_head_seed{seed}.npz")
    d = O.fixmatch_head_details(T(z["logits_w"]), T(z["logits_s"]), float(z["thr"]))
    assert torch.equal(d["idx"], T(z["idx"]))
    assert torch.equal(d["mask"], T(z["mask"]))
    torch.testing.assert_close(d["loss"], T(z["loss"]), rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(d["grad_s"], T(z["grad_s"]), rtol=1e-5, atol=1e-8)
    torch.testing.assert_close(d["mask_mean"], T(z["mask_mean"]), rtol=1e-6, atol=0)


def test_ce_loss_golden():
    z = load_golden("ce_loss.npz")
    x, y, cw = T(z["logits"]), T(z["targets"]), T(z["class_weights"])
    cases = {"plain_none": dict(reduction="none"), "plain_mean_w": dict(reduction="mean", class_weights=cw),
             "poly_mean_w": dict(reduction="mean", class_weights=cw, type_loss="poly"),
             "poly_mean": dict(reduction="mean", type_loss="poly"), "poly_none": dict(reduction="none", type_loss="poly")}
    for nm, kw in cases.items():
        xo = x.clone().requires_grad_(True)
        o = O.ce_loss(xo, y, **kw)
        o.sum().backward()
        torch.testing.assert_close(o.detach(), T(z[nm]), rtol=1e-6, atol=1e-7)
        torch.testing.assert_close(xo.grad, T(z[nm + "_grad"]), rtol=1e-5, atol=1e-8)
    torch.testing.assert_close(O.ce_loss(x, T(z["soft_targets"]), use_hard_labels=False), T(z["soft"]), rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("alias", [0, 1])
def test_ema_golden(alias):
    z = load_golden(f"ema_tinynet_alias{alias}.npz")
    keys = [str(k) for k in z["keys"]]
    decay = float(z["decay"])
    # rebuild the aliasing: entries with identical e0 AND identical final values under
    # aliased names share storage in the reference; emulate with a dict of unique tensors
    e = {}
    uniq = {}
    for k in keys:
        suffix = k.split(".", 1)[1] if k.split(".")[0] in ("model", "backbone") else None
        e[k] = T(z[f"e0/{k}"]).clone()
    # aliasing in _TinyNet: model.0 == conv, model.1 == bn (and backbone.* when alias=1)
    alias_of = {}
    for k in keys:
        for pre, tgt in (("model.0.", "conv."), ("model.1.", "bn."), ("backbone.0.", "conv."), ("backbone.1.", "bn.")):
            if k.startswith(pre):
                alias_of[k] = tgt + k[len(pre):]
    for k, tgt in alias_of.items():
        e[k] = e[tgt]
    for step in range(3):
        O.ema_update_([e[k] for k in keys], [T(z[f"m{step}/{k}"]) for k in keys], decay)
    for k in keys:
        assert torch.equal(e[k], T(z[f"e3/{k}"])), k
    # numpy rounding model on one fp32 tensor (repeat = multiplicity)
    rep = 1 + sum(1 for k, t in alias_of.items() if t == "conv.weight")
    v = z["e0/conv.weight"]
    for step in range(3):
        v = O.ema_update_numpy(v, z[f"m{step}/conv.weight"], decay, rep)
    assert np.array_equal(v, z["e3/conv.weight"])
    # int64 buffer: fp32 arithmetic then truncation (quirk Q3)
    assert z["e3/bn.num_batches_tracked"].dtype == np.int64


@pytest.mark.parametrize("qb", [1, 5])
def test_comatch_train_one_golden(qb):
    z = load_golden(f"comatch_train_one_qb{qb}.npz")
    m = golden_meta(z)
    B, MU, D, C = m["B"], m["MU"], m["D"], m["C"]
    Bu = B * MU
    state = O.CoMatchState.zeros(qb * (MU + 1) * B, D, C)
    for i in range(m["nsteps"]):
        lg, ft, tx = T(z[f"s{i}/logits"]), T(z[f"s{i}/feats"]), T(z[f"s{i}/targets_x"])
        luw, lus0, _ = torch.split(lg[B:], Bu)
        fuw, fus0, fus1 = torch.split(ft[B:], Bu)
        o = O.comatch_head(state, luw, lus0, fuw, fus0, fus1, ft[:B], tx, thr=m["thr"], num_classes=C,
                           enqueue_mode="reference")
        lx = O.ce_loss(lg[:B], tx, None, reduction="mean", type_loss="poly")
        total = float(lx + o["loss_u"] + o["loss_contrast"])
        assert abs(total - float(z[f"s{i}/total_loss"])) <= 2e-6 * abs(total)
        assert torch.equal(o["mask"], T(z[f"s{i}/mask"]))
        assert torch.equal(o["lbs"], T(z[f"s{i}/lbs"]))
        assert torch.equal(state.queue_feats, T(z[f"s{i}/queue_feats"]))
        assert torch.equal(state.queue_probs, T(z[f"s{i}/queue_probs"]))
        assert state.queue_ptr == int(z[f"s{i}/queue_ptr"])
        g = T(z[f"s{i}/grad_logits"])
        torch.testing.assert_close(o["grad_logits_s0"], g[B + Bu:B + 2 * Bu], rtol=1e-5, atol=1e-8)
        gf = T(z[f"s{i}/grad_feats"])
        torch.testing.assert_close(o["grad_feats_s0"], gf[B + Bu:B + 2 * Bu], rtol=1e-5, atol=1e-8)
        torch.testing.assert_close(o["grad_feats_s1"], gf[B + 2 * Bu:], rtol=1e-5, atol=1e-8)


def test_fixmatch_train_one_golden():
    z = load_golden("fixmatch_train_one.npz")
    m = golden_meta(z)
    B, Bu = m["B"], m["B"] * m["MU"]
    for i in range(2):
        lg, tx = T(z[f"s{i}/logits"]), T(z[f"s{i}/targets_x"])
        lx = O.ce_loss(lg[:B], tx, None, reduction="mean", type_loss="poly")
        w, s = lg[B:].chunk(2)
        d = O.fixmatch_head_details(w, s, m["thr"])
        assert abs(float(lx + d["loss"]) - float(z[f"s{i}/total_loss"])) < 2e-6 * float(z[f"s{i}/total_loss"])
        torch.testing.assert_close(d["grad_s"], T(z[f"s{i}/grad_logits"])[B + Bu:], rtol=1e-5, atol=1e-9)


def test_enqueue_always_wraps():
    st = O.CoMatchState.zeros(10, 4, 3)
    for step in range(4):
        f = torch.full((4, 4), float(step + 1))
        p = torch.full((4, 3), float(step + 1))
        O.bank_enqueue(st, f, p, "always")
    assert st.queue_ptr == 6
    assert st.queue_feats[:, 0].tolist() == [3, 3, 4, 4, 4, 4, 2, 2, 3, 3]


def test_sharded_oracle_equals_concat():
    """R ranks smoothing against the global bank == what each rank would compute alone
    with the full bank; enqueue order is rank-major."""
    torch.manual_seed(0)
    C, D, K, B, Bu, R = 5, 8, 48, 2, 6, 2
    st = O.CoMatchState(torch.nn.functional.normalize(torch.randn(K, D), dim=1), torch.softmax(torch.randn(K, C), 1), 40, [])
    hist = [[], []]
    inputs = []
    for r in range(R):
        nf = lambda n: torch.nn.functional.normalize(torch.randn(n, D), dim=1)
        inputs.append(dict(logits_u_w=torch.randn(Bu, C), logits_u_s0=torch.randn(Bu, C), feats_u_w=nf(Bu),
                           feats_u_s0=nf(Bu), feats_u_s1=nf(Bu), feats_x=nf(B), targets_x=torch.randint(0, C, (B,))))
    ref = st.clone()
    outs = O.comatch_head_sharded(st, hist, inputs, thr=0.3, num_classes=C)
    assert st.queue_ptr == (40 + R * (B + Bu)) % K
    rows = (40 + torch.arange(R * (B + Bu))) % K
    want_f = torch.cat([torch.cat([i["feats_u_w"], i["feats_x"]]) for i in inputs])
    assert torch.equal(st.queue_feats[rows], want_f)
    solo = O.comatch_head(ref.clone(), **inputs[1], thr=0.3, num_classes=C, do_enqueue=False)
    assert torch.equal(solo["probs"], outs[1]["probs"])
