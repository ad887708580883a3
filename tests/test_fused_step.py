"""optimizer step + EMA as one launch (`fused_step.FusedOptimizerEMA`).

CPU: the per-element arithmetic the kernel implements (restated with numpy fp32 scalars, same operation order as
csrc/opt_ema.cu) and the per-step group scalars (`group_row`) against `torch.optim` + the reference EMA update on CPU.
GPU: the real launch against `optimizer.step(); ema.update(model)` on the same device, several steps, all three
optimizers `code/optimizer.py:43-51` builds, weight-decay / no-decay groups, buffers, ragged sizes."""
import copy
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))

KINDS = {"sgd": 0, "adam": 1, "adamw": 2}


def _make_opt(kind, groups, lr):
    if kind == "sgd":
        return torch.optim.SGD(groups, momentum=0.9, nesterov=True, lr=lr, weight_decay=0.05)
    if kind == "adamw":
        return torch.optim.AdamW(groups, eps=1e-8, betas=(0.9, 0.999), lr=lr, weight_decay=0.05)
    return torch.optim.Adam(groups, lr=lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0)


def _kernel_math(kind, row, p, g, s1, s2, e, d, o, rep):
    """csrc/opt_ema.cu::update_elem on numpy fp32 arrays (fused multiply-adds evaluated in double then rounded)."""
    f = np.float32
    names = ["lr", "beta1", "beta2", "eps", "wd", "step_size", "bias2_sqrt", "momentum", "kind", "nesterov", "first", "rsv",
             "omb1", "omb2", "decay_factor", "pad"]
    h = dict(zip(names, row))
    fma = lambda a, b, c: (np.float64(a) * np.float64(b) + np.float64(c)).astype(np.float32)
    if kind == 0:
        if h["wd"]:
            g = fma(f(h["wd"]), p, g)
        if h["momentum"]:
            s1 = g.copy() if h["first"] else (s1 * f(h["momentum"])).astype(f) + g
            g = fma(f(h["momentum"]), s1, g) if h["nesterov"] else s1
        p = fma(-f(h["lr"]), g, p)
    else:
        if kind == 2:
            p = (p * f(h["decay_factor"])).astype(f)
        elif h["wd"]:
            g = fma(f(h["wd"]), p, g)
        s1 = fma(f(h["omb1"]), (g - s1).astype(f), s1)
        s2 = fma((f(h["omb2"]) * g).astype(f), g, (s2 * f(h["beta2"])).astype(f))
        denom = (np.sqrt(s2).astype(f) / f(h["bias2_sqrt"])).astype(f) + f(h["eps"])
        p = fma(-f(h["step_size"]), (s1 / denom).astype(f), p)
    for _ in range(rep):
        e = (f(d) * e).astype(f) + (f(o) * p).astype(f)
    return p, s1, s2, e


@pytest.mark.parametrize("kind", ["sgd", "adam", "adamw"])
def test_kernel_arithmetic_and_group_scalars_match_torch_optim_cpu(kind):
    from endoscopy_image_classification_b200.fused_step import group_row
    torch.manual_seed(0)
    w = torch.nn.Parameter(torch.randn(257))
    b = torch.nn.Parameter(torch.randn(31))
    opt = _make_opt(kind, [{"params": [w]}, {"params": [b], "weight_decay": 0.0}], lr=3e-3)
    ema = [w.detach().clone(), b.detach().clone()]
    mine = [dict(p=t.detach().numpy().copy(), s1=np.zeros(t.numel(), np.float32), s2=np.zeros(t.numel(), np.float32),
                 e=t.detach().numpy().copy()) for t in (w, b)]
    d = 0.999
    d32, o32 = np.float32(d), np.float32(1.0 - d)
    for step in range(1, 6):
        grads = [torch.randn_like(w), torch.randn_like(b)]
        opt.param_groups[0]["lr"] = opt.param_groups[1]["lr"] = 3e-3 * (1.0 - 0.1 * step)     # a scheduler moves the lr
        for t, g in zip((w, b), grads):
            t.grad = g.clone()
        opt.step()
        for e, t in zip(ema, (w, b)):                       # code/ema.py:53-56, applied twice for an aliased storage
            for _ in range(2):
                e.mul_(d).add_((1.0 - d) * t.detach())
        for gi, (m, g) in enumerate(zip(mine, grads)):
            row = group_row(KINDS[kind], opt.param_groups[gi], step)
            m["p"], m["s1"], m["s2"], m["e"] = _kernel_math(KINDS[kind], row, m["p"], g.numpy(), m["s1"], m["s2"], m["e"], d32, o32, 2)
    for m, t, e in zip(mine, (w, b), ema):
        np.testing.assert_allclose(m["p"], t.detach().numpy(), rtol=2e-6, atol=1e-7)
        np.testing.assert_allclose(m["e"], e.numpy(), rtol=2e-6, atol=1e-7)
    st = opt.state[w]
    if kind == "sgd":
        np.testing.assert_allclose(mine[0]["s1"], st["momentum_buffer"].numpy(), rtol=2e-6, atol=1e-7)
    else:
        np.testing.assert_allclose(mine[0]["s1"], st["exp_avg"].numpy(), rtol=2e-6, atol=1e-7)
        np.testing.assert_allclose(mine[0]["s2"], st["exp_avg_sq"].numpy(), rtol=2e-6, atol=1e-9)


def test_fused_step_rejects_what_it_does_not_implement():
    from endoscopy_image_classification_b200.fused_step import FusedOptimizerEMA
    w = torch.nn.Parameter(torch.randn(4))
    with pytest.raises(NotImplementedError):
        FusedOptimizerEMA(torch.optim.RMSprop([w]))
    with pytest.raises(NotImplementedError):
        FusedOptimizerEMA(torch.optim.Adam([w], amsgrad=True))
    with pytest.raises(NotImplementedError):
        FusedOptimizerEMA(torch.optim.SGD([w], lr=0.1, momentum=0.9, dampening=0.5))
    f = FusedOptimizerEMA(torch.optim.Adam([w]))
    w.grad = torch.zeros(4)
    with pytest.raises(RuntimeError, match="CUDA"):          # no CPU path
        f.step()


class _Net(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.conv = torch.nn.Conv2d(3, 24, 3, bias=False)      # 648 weights
        self.bn = torch.nn.BatchNorm2d(24)                     # 1-D params (no decay) + float / int64 buffers
        self.fc1 = torch.nn.Linear(24, 517)                    # 12408 weights: several 4096-chunks, ragged tail
        self.fc2 = torch.nn.Linear(517, 23)
        self.frozen = torch.nn.Linear(7, 5)
        self.frozen.requires_grad_(False)

    def forward(self, x):
        h = torch.relu(self.bn(self.conv(x))).mean((2, 3))
        return self.fc2(torch.relu(self.fc1(h)))


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["sgd", "adam", "adamw"])
@pytest.mark.parametrize("use_ema", [True, False])
def test_fused_optimizer_ema_matches_torch_optim_then_ema(kind, use_ema):
    from endoscopy_image_classification_b200.ema import ModelEMA
    from endoscopy_image_classification_b200.fused_step import FusedOptimizerEMA
    from endoscopy_image_classification_b200.optimizer import set_weight_decay
    dev = torch.device("cuda:0")
    torch.manual_seed(1)
    ref = _Net().to(dev)
    new = copy.deepcopy(ref)
    opt_r = _make_opt(kind, set_weight_decay(ref), lr=2e-3)
    opt_n = _make_opt(kind, set_weight_decay(new), lr=2e-3)
    ema_r = ModelEMA(ref, 0.99, device=dev) if use_ema else None
    ema_n = ModelEMA(new, 0.99, device=dev) if use_ema else None
    fused = FusedOptimizerEMA(opt_n, ema_n, new)
    g = torch.Generator(device=dev).manual_seed(3)
    for step in range(5):
        x = torch.randn(8, 3, 9, 9, device=dev, generator=g)
        y = torch.randint(0, 23, (8,), device=dev, generator=g)
        torch.nn.functional.cross_entropy(ref(x), y).backward()
        with torch.no_grad():
            new(x)                                          # BatchNorm running statistics move in both models
        for pr, pn in zip(ref.parameters(), new.parameters()):      # identical gradients for both optimizers
            if pr.grad is not None:
                if pn.grad is None:
                    pn.grad = pr.grad.clone()
                else:
                    pn.grad.copy_(pr.grad)
        for o in (opt_r, opt_n):
            for grp in o.param_groups:
                grp["lr"] = 2e-3 * (1.0 - 0.15 * step)
        opt_r.step()
        if use_ema:
            ema_r.update(ref)
        opt_r.zero_grad(set_to_none=False)
        fused.step()
        fused.zero_grad()
        if step == 2:                                       # a gradient that moved is detected: tables are rebuilt
            opt_n.zero_grad(set_to_none=True)
    tol = dict(rtol=2e-6, atol=1e-7)
    for (n, a), (_, b) in zip(ref.state_dict().items(), new.state_dict().items()):
        torch.testing.assert_close(b, a, **tol, msg=lambda m, n=n: f"model {n}: {m}")
    if use_ema:
        for (n, a), (_, b) in zip(ema_r.ema.state_dict().items(), ema_n.ema.state_dict().items()):
            torch.testing.assert_close(b, a, **tol, msg=lambda m, n=n: f"ema {n}: {m}")
    sr, sn = opt_r.state_dict()["state"], opt_n.state_dict()["state"]
    assert sr.keys() == sn.keys()
    for k in sr:
        assert sr[k].keys() == sn[k].keys(), (k, sr[k].keys(), sn[k].keys())
        for name in sr[k]:
            torch.testing.assert_close(sn[k][name].float().cpu(), sr[k][name].float().cpu(), rtol=2e-6, atol=1e-9,
                                       msg=lambda m, k=k, name=name: f"state[{k}][{name}]: {m}")
    # the wrapped optimizer can simply take over again
    for pn in new.parameters():
        if pn.requires_grad:
            pn.grad = torch.ones_like(pn)
    opt_n.step()


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["sgd", "adam", "adamw"])
def test_captured_fused_step_replays_bit_exact(kind):
    """SURVEY 8 f1 inside a CUDA graph: ``FusedOptimizerEMA.capture()`` keeps the group scalars in device memory, so the
    replayed launch follows the LR schedule and Adam's bias corrections; 6 replays == 6 eager fused steps, bit for bit
    (weights, optimizer state, EMA, step counters)."""
    from endoscopy_image_classification_b200.ema import ModelEMA
    from endoscopy_image_classification_b200.fused_step import FusedOptimizerEMA
    from endoscopy_image_classification_b200.optimizer import set_weight_decay
    dev = torch.device("cuda:0")
    torch.manual_seed(2)
    a = _Net().to(dev)
    b = copy.deepcopy(a)
    gen = torch.Generator(device=dev).manual_seed(9)
    for pa, pb in zip(a.parameters(), b.parameters()):
        pa.grad = torch.randn(pa.shape, device=dev, generator=gen)
        pb.grad = pa.grad.clone()
    opt_a, opt_b = _make_opt(kind, set_weight_decay(a), lr=2e-3), _make_opt(kind, set_weight_decay(b), lr=2e-3)
    ema_a, ema_b = ModelEMA(a, 0.99, device=dev), ModelEMA(b, 0.99, device=dev)
    eager, graphed = FusedOptimizerEMA(opt_a, ema_a, a), FusedOptimizerEMA(opt_b, ema_b, b)
    eager.step()                                   # first step eagerly on both sides (SGD: the momentum buffer is born here)
    graphed.step()
    cap = graphed.capture()
    before = [p.detach().clone() for p in b.parameters()]
    for p, q in zip(b.parameters(), before):       # capturing records, it does not step
        assert torch.equal(p, q)
    for step in range(6):
        for o in (opt_a, opt_b):
            for grp in o.param_groups:
                grp["lr"] = 2e-3 * (1.0 - 0.1 * step)
        for pa, pb in zip(a.parameters(), b.parameters()):
            pa.grad.copy_(torch.randn(pa.shape, device=dev, generator=gen))
            pb.grad.copy_(pa.grad)
        eager.step()
        cap.replay()
    torch.cuda.synchronize()
    for (n, x), (_, y) in zip(a.state_dict().items(), b.state_dict().items()):
        assert torch.equal(x, y), n
    for (n, x), (_, y) in zip(ema_a.ema.state_dict().items(), ema_b.ema.state_dict().items()):
        assert torch.equal(x, y), n
    sa, sb = opt_a.state_dict()["state"], opt_b.state_dict()["state"]
    for k in sa:
        for name in sa[k]:
            assert torch.equal(torch.as_tensor(sa[k][name]).cpu(), torch.as_tensor(sb[k][name]).cpu()), (k, name)
    for p in b.parameters():                       # a gradient that moves invalidates the captured tables
        p.grad = p.grad.clone()
        break
    with pytest.raises(RuntimeError):
        cap.replay()


class _StubLib:
    """Records the C-ABI calls of the fused step so its host logic can run on CPU tensors (no GPU here)."""

    def __init__(self):
        self.calls = []

    def b200ssl_opt_ema_multi_tensor(self, blocks, n_blocks, groups, n_groups, decay, omd, stream):
        import ctypes
        from endoscopy_image_classification_b200.fused_step import GROUP
        raw = ctypes.string_at(groups, n_groups * GROUP.itemsize)
        self.calls.append(dict(n_blocks=n_blocks, groups=np.frombuffer(raw, dtype=GROUP).copy(), decay=decay, omd=omd))
        return 0

    def b200ssl_ema_multi_tensor(self, *a):
        self.calls.append(dict(rest=a))
        return 0


def test_fused_step_host_logic_tables_counters_and_hooks(monkeypatch):
    """Block / group tables, EMA pairing (aliased storages, entries left to the plain EMA launch), lazily written step
    counters and the rebuild on a moved gradient -- with the native launch stubbed out."""
    from endoscopy_image_classification_b200 import _native as N
    from endoscopy_image_classification_b200 import ema as ema_mod
    from endoscopy_image_classification_b200 import fused_step as F
    stub = _StubLib()
    monkeypatch.setattr(N, "lib", lambda: stub)
    monkeypatch.setattr(N, "require_cuda", lambda *a, **k: torch.device("cpu"))
    monkeypatch.setattr(N, "stream_ptr", lambda d: 0)

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.body = torch.nn.Linear(100, 90)            # 9000 weights -> 3 chunks of <= 4096
            self.bn = torch.nn.BatchNorm1d(90)
            self.alias = self.body                          # the same storages under a second name (custom_model.py:194-200)
            self.frozen = torch.nn.Linear(3, 3).requires_grad_(False)

    model = Net()
    ema = ema_mod.ModelEMA.__new__(ema_mod.ModelEMA)        # the constructor insists on CUDA; the fields are all we need
    ema.ema, ema.decay = copy.deepcopy(model), 0.99
    opt = torch.optim.AdamW([{"params": [model.body.weight]}, {"params": [model.body.bias, model.bn.weight, model.bn.bias],
                                                               "weight_decay": 0.0}], lr=1e-3, weight_decay=0.05)
    fused = F.FusedOptimizerEMA(opt, ema, model)
    for p in opt.param_groups[0]["params"] + opt.param_groups[1]["params"]:
        p.grad = torch.ones_like(p)
    fused.step()
    t = fused._tables
    tbl = np.frombuffer(t["blocks"].numpy().tobytes(), dtype=F.BLOCK)
    assert t["n_blocks"] == len(tbl) == 3 + 1 + 1 + 1
    w = model.body.weight
    assert list(tbl["count"][:3]) == [4096, 4096, 9000 - 8192] and list(tbl["group"]) == [0, 0, 0, 1, 1, 1]
    assert list(tbl["param"][:3]) == [w.data_ptr() + 4 * o for o in (0, 4096, 8192)]
    assert list(tbl["grad"][:3]) == [w.grad.data_ptr() + 4 * o for o in (0, 4096, 8192)]
    assert list(tbl["s1"][:3]) == [opt.state[w]["exp_avg"].data_ptr() + 4 * o for o in (0, 4096, 8192)]
    assert list(tbl["ema"][:3]) == [ema.ema.body.weight.data_ptr() + 4 * o for o in (0, 4096, 8192)]
    assert list(tbl["ema_repeat"]) == [2, 2, 2, 2, 1, 1]    # body.* appears twice in state_dict(), bn.* once
    # everything the optimizer does not own stays with the plain EMA launch: bn buffers + the frozen layer
    rest = t["rest"]
    assert rest.n_unique == 3 + 2 and any("rest" in c for c in stub.calls)
    g0, g1 = stub.calls[0]["groups"]
    assert g0["kind"] == g1["kind"] == F.OPT_ADAMW and g0["first_step"] == 1
    assert np.isclose(g0["decay_factor"], 1 - 1e-3 * 0.05) and g1["decay_factor"] == 1.0
    assert np.isclose(g0["step_size"], 1e-3 / (1 - 0.9)) and np.isclose(g0["bias2_sqrt"], np.sqrt(1 - 0.999))
    # step counters: written back lazily, visible through state_dict() and before a plain optimizer.step()
    fused.step()
    assert float(opt.state[w]["step"]) == 0.0 and fused._pending == 2
    assert stub.calls[-2]["groups"][0]["first_step"] == 0
    assert np.isclose(stub.calls[-2]["groups"][0]["step_size"], 1e-3 / (1 - 0.9 ** 2))
    assert all(float(s["step"]) == 2.0 for s in opt.state_dict()["state"].values()) and fused._pending == 0
    # a gradient that moved (zero_grad(set_to_none=True) + backward) rebuilds the table; in-place zeroing does not
    fused.zero_grad()
    assert fused._valid() and float(w.grad.abs().sum()) == 0.0
    w.grad = torch.ones_like(w)
    assert not fused._valid()
    fused.step()
    assert fused._tables is not t and fused._tables["base"] == [2, 2] and fused._pending == 1
    opt.step()                                              # hands back: the hook flushes first
    assert float(opt.state[w]["step"]) == 4.0
    opt.load_state_dict(opt.state_dict())
    assert fused._tables is None
