"""Multi-rank parity of the peer-memory CoMatch bank on ONE GPU: ``world`` emulated ranks of one process
(``peer.LocalArenaSet``) run the real multi-rank kernels -- K3 over all shards (row loop, one tensor map per shard), the
rank-major ring enqueue into the owning shard / every replica, the epoch-flag protocol -- rank after rank in lock step
(``comatch_head.lockstep_total_loss``), against the single-process oracle for the concatenated batch with the whole bank.

Why emulation: kernels of different ranks that wait for each other's flags must not be run as concurrent launches /
processes on one GPU (nothing guarantees co-scheduling).  In lock step every flag a kernel reads was published by an
earlier launch of the same stream, so nothing ever spins.  ``tests/test_gpu_sharded.py`` runs the same modes with real
processes when the box has >= 2 GPUs."""
import pytest
import torch

from conftest import rel_err
from oracle import ssl_oracle as O

pytestmark = pytest.mark.gpu
C, D, B, MU, THR = 23, 64, 16, 7, 0.9
N_ROWS = B + B * MU


def _inputs(seed, dtype=torch.bfloat16):
    from endoscopy_image_classification_b200.synthetic import comatch_step_inputs, rownorm
    g = torch.Generator().manual_seed(seed)
    protos = rownorm(torch.randn(C, D, generator=torch.Generator().manual_seed(7)))
    b = comatch_step_inputs(g, B, MU, D, C, protos, dtype)
    b.pop("logits_x")
    return b


def _run(world, exchange, K, steps, prefill, ptr0=0):
    from endoscopy_image_classification_b200.comatch_head import CoMatchHead, lockstep_total_loss
    from endoscopy_image_classification_b200.peer import LocalArenaSet
    dev = torch.device("cuda", 0)
    ranks = LocalArenaSet(world, dev)
    heads = [CoMatchHead(C, D, K, THR, enqueue_mode="always", device=dev, dtype=torch.bfloat16, exchange=exchange,
                         local_ranks=(ranks, r)) for r in range(world)]
    assert all(h.exchange == exchange and h._shards is not None for h in heads)
    state = O.CoMatchState.zeros(K, D, C)
    if prefill:
        g0 = torch.Generator().manual_seed(5)
        qf = torch.nn.functional.normalize(torch.randn(K, D, generator=g0), dim=1).to(torch.bfloat16)
        qp = torch.softmax(2.0 * torch.randn(K, C, generator=g0), 1).to(torch.bfloat16)
        for h in heads:
            lo, hi = (0, K) if exchange == "replicated" else (h.geom.shard_begin, h.geom.shard_begin + h.geom.shard_rows)
            h.queue_feats.copy_(qf[lo:hi])
            h.queue_probs.copy_(qp[lo:hi])
            h.queue_probs_t[:C].copy_(qp[lo:hi].t())
        state.queue_feats.copy_(qf.float())
        state.queue_probs.copy_(qp.float())
    if ptr0:
        for h in heads:
            h.queue_ptr = ptr0
        state.queue_ptr = ptr0
    hist = [[] for _ in range(world)]
    tol = 1e-2
    for step in range(steps):
        cpu = [_inputs(100 * step + r) for r in range(world)]
        batches = []
        for b in cpu:
            d = {k: v.cuda() for k, v in b.items()}
            for k in ("logits_u_s0", "feats_u_s0", "feats_u_s1"):
                d[k].requires_grad_(True)
            batches.append(d)
        outs = lockstep_total_loss(heads, batches, lambda_u=2.0, lambda_c=0.5)
        for o in outs:
            o[0].backward()
        state.queue_probs = state.queue_probs.to(torch.bfloat16).float()      # the device bank stores bf16 rows
        ref = O.comatch_head_sharded(state, hist, [{k: (v.float() if v.is_floating_point() else v) for k, v in b.items()} for b in cpu],
                                     thr=THR, num_classes=C)
        for r, h in enumerate(heads):
            assert rel_err(h.last["probs"], ref[r]["probs"]) < tol, (step, r)
            if bool((h.last["mask"].cpu() == ref[r]["mask"]).all()):
                want = float(2.0 * ref[r]["loss_u"] + 0.5 * ref[r]["loss_contrast"])
                assert abs(float(outs[r][0]) - want) < tol * abs(want), (step, r)
            assert rel_err(batches[r]["feats_u_s0"].grad.float(), 0.5 * ref[r]["grad_feats_s0"]) < tol
            assert h.queue_ptr == int(h.ptr_state[0]) == state.queue_ptr == (ptr0 + (step + 1) * world * N_ROWS) % K
    torch.cuda.synchronize()
    assert all(h.peer_timeouts() == 0 for h in heads)
    if exchange == "replicated":
        banks = [(h.queue_feats.float().cpu(), h.queue_probs.float().cpu()) for h in heads]
    else:
        banks = [(torch.cat([h.queue_feats.float().cpu() for h in heads]), torch.cat([h.queue_probs.float().cpu() for h in heads]))]
    for bank_f, bank_p in banks:
        assert torch.equal(bank_f, state.queue_feats)          # copied embedding rows: bit-exact, ring positions included
        assert float((bank_p - state.queue_probs).abs().max()) < tol
    for h in heads:
        h.close()
    ranks.close()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("exchange", ["direct", "replicated"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_emulated_ranks_ring_wrap(world, exchange):
    """The ring holds 3 steps of `world` ranks: step 4 wraps; the shards (3 blocks = 384 rows) are not a multiple of the
    128-key tile, so every shard ends in a partial TMA tile."""
    _run(world, exchange, K=3 * world * N_ROWS, steps=4, prefill=False)


@pytest.mark.timeout(600)
@pytest.mark.parametrize("world,K", [(2, 8192), (8, 65536), (4, 8 * 1000)])
def test_emulated_ranks_prefilled_bank(world, K):
    """Smoothing against a full (pre-filled) sharded bank -- incl. BASELINE cfg 4's 65536 rows over 8 ranks and shards of
    2000 rows (8-row aligned, not tile aligned).  The ring pointer starts 100 rows before a shard boundary and is not a
    multiple of 8: blocks straddle two shards and take the unaligned store path."""
    _run(world, "direct", K=K, steps=2, prefill=True, ptr0=K // world - 100)
