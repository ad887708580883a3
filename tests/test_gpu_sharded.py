"""2-GPU (NCCL) parity of the rank-sharded CoMatch bank: real kernels on every rank against the
single-process oracle for the concatenated batch with the full bank.  Skipped with
fewer than two GPUs (run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_sharded.py -m gpu`)."""
import os
import socket
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
REPO = Path(__file__).resolve().parents[1]
C, D, B, MU, THR, STEPS = 23, 64, 16, 7, 0.9, 4
K = 3 * 2 * (B + B * MU)          # 3 steps of 2 ranks: the ring wraps at step 4


def _inputs(seed, dtype=torch.float32):
    sys.path.insert(0, str(REPO))
    from endoscopy_image_classification_b200.synthetic import comatch_step_inputs, rownorm
    g = torch.Generator().manual_seed(seed)
    protos = rownorm(torch.randn(C, D, generator=torch.Generator().manual_seed(7)))
    b = comatch_step_inputs(g, B, MU, D, C, protos, dtype)
    b.pop("logits_x")
    return b


def _check_peer_primitives(rank, world, dev):
    """PeerArena launches against the NCCL collectives on the same data: many epochs (slot parity, epoch
    counters), two-part gathers, and a CUDA-graph replay of both exchanges."""
    from endoscopy_image_classification_b200.peer import PeerArena
    arena = PeerArena(dist.group.WORLD, dev, {0: 471 * 128, 1: 471 * 96, 3: 4096})
    g = torch.Generator(device=dev).manual_seed(11 + rank)
    for it in range(9):
        a = torch.randn(448 - 8 * it, 64, generator=g, device=dev).to(torch.bfloat16)
        b = torch.randn(23, 64, generator=g, device=dev).to(torch.bfloat16)
        got = arena.all_gather(0, [a, b])
        want = torch.empty(world * (a.shape[0] + 23), 64, dtype=torch.bfloat16, device=dev)
        dist.all_gather_into_tensor(want, torch.cat([a, b]))
        assert torch.equal(got, want), ("all_gather", it)
        part = torch.randn(world * (471 - it), 24, generator=g, device=dev)
        got = arena.reduce_scatter(1, part)
        want = torch.empty(471 - it, 24, device=dev)
        dist.reduce_scatter_tensor(want, part.clone())
        if world == 2:
            assert torch.equal(got, want), ("reduce_scatter", it)      # two addends: order cannot matter
        torch.testing.assert_close(got, want, rtol=1e-6, atol=1e-6)
    # graph replay: epochs live on the device
    src = torch.zeros(64, 16, device=dev)
    part = torch.zeros(world * 32, 8, device=dev)
    torch.cuda.synchronize()
    dist.barrier()
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        arena.all_gather(3, [src])
        arena.reduce_scatter(1, part)
        torch.cuda.synchronize()
        with torch.cuda.graph(graph, stream=side):
            out_g = arena.all_gather(3, [src])
            out_r = arena.reduce_scatter(1, part)
    torch.cuda.current_stream().wait_stream(side)
    for it in range(5):
        src.fill_(float(10 * it + rank))
        part.fill_(float(it + 1 + rank))
        graph.replay()
        torch.cuda.synchronize()
        want = torch.cat([torch.full((64, 16), float(10 * it + r)) for r in range(world)])
        assert torch.equal(out_g.cpu(), want), ("graph all_gather", it)
        assert torch.equal(out_r.cpu(), torch.full((32, 8), float(sum(it + 1 + r for r in range(world))))), ("graph reduce_scatter", it)
    assert arena.timeouts() == 0
    del graph
    arena.close()


def _worker(rank, world, port, out_dir, dtype_name, exchange):
    sys.path.insert(0, str(REPO))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from endoscopy_image_classification_b200.comatch_head import CoMatchHead
    dtype = getattr(torch, dtype_name)
    if exchange == "peer" and dtype_name == "float32":
        _check_peer_primitives(rank, world, torch.device("cuda", rank))
    head = CoMatchHead(C, D, K, THR, enqueue_mode="always", device=f"cuda:{rank}", dtype=dtype, process_group=dist.group.WORLD,
                       exchange=exchange)
    outs = []
    for step in range(STEPS):
        inp = {k: v.cuda() for k, v in _inputs(100 * step + rank, dtype).items()}
        for k in ("logits_u_s0", "feats_u_s0", "feats_u_s1"):
            inp[k].requires_grad_(True)
        total, lu, lc, mm = head.total_loss(**inp, lambda_u=2.0, lambda_c=0.5)
        total.backward()
        outs.append(dict(total=total.detach().cpu(), probs=head.last["probs"].cpu(), mask=head.last["mask"].cpu(),
                         g_f0=inp["feats_u_s0"].grad.float().cpu(), ptr=head.queue_ptr, dev_ptr=int(head.ptr_state[0])))
    torch.save(dict(outs=outs, qf=head.queue_feats.float().cpu(), qp=head.queue_probs.float().cpu()),
               os.path.join(out_dir, f"rank{rank}.pt"))
    assert head.peer_timeouts() == 0
    assert head.exchange == exchange and (head._arena is not None) == (exchange != "collective")
    head.close()
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(300)
@pytest.mark.parametrize("exchange", ["replicated", "direct", "peer", "collective"])
@pytest.mark.parametrize("dtype_name,tol", [("float32", 1e-5), ("bfloat16", 1e-2)])
def test_two_gpu_sharded_bank(tmp_path, dtype_name, tol, exchange):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    if exchange in ("direct", "replicated") and dtype_name != "bfloat16":
        pytest.skip("the peer-memory resident bank is the bf16 tensor-core layout")
    sys.path.insert(0, str(REPO))
    from oracle import ssl_oracle as O
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), dtype_name, exchange), nprocs=world, join=True)
    res = [torch.load(tmp_path / f"rank{r}.pt", weights_only=False) for r in range(world)]
    dtype = getattr(torch, dtype_name)
    state = O.CoMatchState.zeros(K, D, C)
    hist = [[] for _ in range(world)]
    n = B + B * MU
    for step in range(STEPS):
        inputs = [{k: (v.float() if v.is_floating_point() else v) for k, v in _inputs(100 * step + r, dtype).items()} for r in range(world)]
        if dtype != torch.float32:
            state.queue_probs = state.queue_probs.to(dtype).float()
        ref = O.comatch_head_sharded(state, hist, inputs, thr=THR, num_classes=C)
        for r in range(world):
            got = res[r]["outs"][step]
            err = float((got["probs"].double() - ref[r]["probs"].double()).abs().max() / ref[r]["probs"].abs().max())
            assert err < tol, (step, r, err)
            if bool((got["mask"] == ref[r]["mask"]).all()):
                want = float(2.0 * ref[r]["loss_u"] + 0.5 * ref[r]["loss_contrast"])
                assert abs(float(got["total"]) - want) < tol * abs(want)
            gref = 0.5 * ref[r]["grad_feats_s0"]
            assert float((got["g_f0"] - gref).abs().max() / gref.abs().max()) < tol
            assert got["ptr"] == got["dev_ptr"] == state.queue_ptr == ((step + 1) * world * n) % K
    # shards concatenate to the oracle's ring; with exchange='replicated' every rank holds the whole ring
    banks = ([(res[r]["qf"], res[r]["qp"]) for r in range(world)] if exchange == "replicated" else
             [(torch.cat([res[r]["qf"] for r in range(world)]), torch.cat([res[r]["qp"] for r in range(world)]))])
    for bank_f, bank_p in banks:
        assert torch.equal(bank_f, state.queue_feats)
        assert float((bank_p - state.queue_probs).abs().max()) < tol
