"""CPU restatements (numpy / torch fp64) of the numerics the round-2 kernels rest on -- no GPU, no library call.  Each test
states the claim a kernel makes about its arithmetic and checks it against fp64 on BASELINE-shaped data:

* K3 / K6 with fp32 storage run on bf16 operands: which split (terms per operand, cross products kept) meets the 1e-5 bar
  (csrc/bank_tc.cu NT = 2, csrc/contrast_tc.cu NT = 2);
* the contrastive forward closes its second pass in the moments of its first one (csrc/contrast_tc.cu, kZLim).
"""
import math

import numpy as np
import torch

C, D, T, TH = 23, 64, 0.2, 0.8


def bf(x):
    return x.to(torch.bfloat16).to(torch.float64)


def split(x, n):
    out, r = [], x.clone()
    for _ in range(n):
        h = bf(r)
        out.append(h)
        r = r - h
    return out


def clustered(g, n, protos, noise=0.35):
    lab = torch.randint(0, C, (n,), generator=g)
    return torch.nn.functional.normalize(protos[lab] + noise * torch.randn(n, D, generator=g, dtype=torch.float64), dim=1), lab


def test_k3_split_operands_meet_the_fp32_bar():
    """Smoothed probabilities from bf16 hi + mid operands with three cross terms per GEMM are ~1e-6 from fp64 (the reference's
    fp32 torch.mm: ~3e-7); one bf16 term (the bf16-storage kernel) is ~1e-3; dropping the split of P / Qp alone costs 5e-4."""
    g = torch.Generator().manual_seed(0)
    protos = torch.nn.functional.normalize(torch.randn(C, D, generator=g, dtype=torch.float64), dim=1)
    f, _ = clustered(g, 448, protos)
    q, ql = clustered(g, 2560, protos)
    f, q = f.float().double(), q.float().double()                    # fp32 storage
    qp = torch.softmax(3 * torch.randn(2560, C, generator=g, dtype=torch.float64) + 4 * torch.nn.functional.one_hot(ql, C), 1).float().double()
    A = torch.exp(f @ q.t() / T)
    ref = (A @ qp) / A.sum(1, keepdim=True)

    def run(nf, pairs1, npq, pairs2):
        fs, qs = split(f, nf), split(q, nf)
        S = sum(fs[a] @ qs[b].t() for a, b in pairs1)
        E = torch.exp2(S.float() * np.float32(1.4426950408889634 / T)).double()      # fp32 epilogue
        ps, qps = split(E, npq), split(qp, npq)
        num = sum(ps[a] @ qps[b] for a, b in pairs2)
        rs = sum(p.sum(1, keepdim=True) for p in ps)
        return float(((num / rs) - ref).abs().max() / ref.abs().max())

    p3 = [(1, 0), (0, 1), (0, 0)]
    assert run(2, p3, 2, p3) < 4e-6                                   # what the kernel does
    assert run(1, [(0, 0)], 1, [(0, 0)]) > 1e-4                       # bf16 storage: the 1e-2 bar, not the 1e-5 one
    assert run(2, p3, 1, [(0, 0)]) > 1e-4                             # P and Qp need their second term too


def test_k6_gradient_gemm_needs_three_terms():
    """dF0 = dZ F1 is a sum with heavy cancellation (sum_j dZ_ij = 0): two bf16 terms per operand leave > 1e-5 of the largest
    gradient, three terms with the six leading cross products are at fp32 level."""
    g = torch.Generator().manual_seed(1)
    rows = 448
    protos = torch.nn.functional.normalize(torch.randn(C, D, generator=g, dtype=torch.float64), dim=1)
    f0, lab = clustered(g, rows, protos)
    f1 = torch.nn.functional.normalize(f0 + 0.3 * torch.randn(rows, D, generator=g, dtype=torch.float64), dim=1)
    f0, f1 = f0.float().double(), f1.float().double()
    probs = torch.softmax(6.0 * torch.nn.functional.one_hot(lab, C).double() + torch.randn(rows, C, generator=g, dtype=torch.float64), 1)
    E = torch.exp(f0 @ f1.t() / T)
    P = E / E.sum(1, keepdim=True)
    Q = probs @ probs.t()
    Q.fill_diagonal_(1.0)
    Qm = torch.where(Q >= TH, Q, torch.zeros_like(Q))
    qn = Qm / Qm.sum(1, keepdim=True)
    G = -(qn / (P + 1e-7)) / rows
    dZ = (P * (G - (G * P).sum(1, keepdim=True))).float().double()
    exact = dZ @ f1
    scale = exact.abs().max()

    def err(nz, nf, terms):
        zs, fs = split(dZ, nz), split(f1, nf)
        return float((sum(zs[a] @ fs[b] for a, b in terms) - exact).abs().max() / scale)

    assert err(2, 2, [(0, 0), (0, 1), (1, 0)]) > 5e-6                 # round-2's first attempt: 1.5e-5 on the reference goldens
    assert err(3, 3, [(1, 1), (0, 2), (2, 0), (0, 1), (1, 0), (0, 0)]) < 2e-7      # what the kernel does


def test_k6_forward_closes_pass_b_in_the_moments_of_pass_a():
    """loss_i = -sum_j qn log(P + eps) and r_i = sum_j qn P / (P + eps) from sum q, sum q log2 E, sum q E^-k (k = 1..4): exact to
    the Taylor remainder z^5 / 5 with z = eps rs / E, i.e. far below either bar while max z < kZLim -- and the guard fires
    when it is not (embeddings far from the unit sphere's range), where the kernel falls back to the exact second pass."""
    g = torch.Generator().manual_seed(2)
    rows = 896
    protos = torch.nn.functional.normalize(torch.randn(C, D, generator=g, dtype=torch.float64), dim=1)
    f0, lab = clustered(g, rows, protos)
    f1 = torch.nn.functional.normalize(f0 + 0.4 * torch.randn(rows, D, generator=g, dtype=torch.float64), dim=1)
    probs = torch.softmax(6.0 * torch.nn.functional.one_hot(lab, C).double() + torch.randn(rows, C, generator=g, dtype=torch.float64), 1)

    def both(f0, f1, temp):
        y = (f0 @ f1.t()) * (1.4426950408889634 / temp)               # log2 of E
        E = torch.exp2(y)
        rs = E.sum(1)
        Q = probs @ probs.t()
        Q.fill_diagonal_(1.0)
        qm = torch.where(Q >= TH, Q, torch.zeros_like(Q))
        qs = qm.sum(1)
        P = E / rs[:, None]
        li = -(qm / qs[:, None] * torch.log(P + 1e-7)).sum(1)         # the second pass, exactly
        rr = (qm / qs[:, None] * P / (P + 1e-7)).sum(1)
        m1 = (qm * y).sum(1)
        c = 1e-7 * rs
        s = [c ** k * (qm * E ** (-k)).sum(1) for k in (1, 2, 3, 4)]
        li_c = -(math.log(2.0) * m1 - torch.log(rs) * qs + (s[0] - s[1] / 2 + s[2] / 3 - s[3] / 4)) / qs
        rr_c = (qs - s[0] + s[1] - s[2] + s[3]) / qs
        zmax = (c[:, None] / torch.where(qm > 0, E, torch.full_like(E, float("inf")))).max()
        return li, rr, li_c, rr_c, float(zmax)

    li, rr, li_c, rr_c, zmax = both(f0, f1, T)
    assert zmax < 1e-3                                                # the usual regime: z ~ 1e-5
    assert float((li_c - li).abs().max() / li.abs().max()) < 1e-9 and float((rr_c - rr).abs().max()) < 1e-9
    # a temperature of 0.02 spreads E over 43 decades: pairs of the graph with E ~ 1e-15 of their row sum -> z >> kZLim
    li, rr, li_c, rr_c, zmax = both(f0, f1, 0.02)
    assert zmax > 0.25                                                # the guard of the kernel (kZLim) fires: exact pass B
    # in between: at the guard the series is still inside the bf16 bar, at the fp32 guard inside the fp32 bar
    for zlim, bar in ((0.25, 1e-3), (0.03, 1e-7)):
        z = torch.tensor(zlim, dtype=torch.float64)
        assert abs(float(torch.log1p(z) - (z - z ** 2 / 2 + z ** 3 / 3 - z ** 4 / 4))) < bar
        assert abs(float(1 / (1 + z) - (1 - z + z ** 2 - z ** 3 + z ** 4))) < 10 * bar
