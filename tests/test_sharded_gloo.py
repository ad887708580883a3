"""World-size-2 gloo test (CPU) of the N>1 host path: the rank-sharded memory bank choreography
of CoMatchHead (all-gather of the enqueue block, local-shard smoothing for all ranks' queries,
reduce-scatter of the partial sums, rank-major sharded enqueue, pointer arithmetic).

The CUDA kernels cannot run here, so a TEST-ONLY subclass replaces each one-kernel wrapper
(`_k_*`) by the oracle's math on CPU tensors; everything between the kernels -- the code under
test -- is the product's.  The 2-rank result must equal the single-process oracle for the
concatenated batch with the full bank."""
import os
import socket
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = Path(__file__).resolve().parents[1]
C, D, B, MU, K, THR, STEPS = 7, 16, 2, 3, 48, 0.5, 5


def _inputs(seed):
    g = torch.Generator().manual_seed(seed)
    Bu = B * MU
    nf = lambda n: torch.nn.functional.normalize(torch.randn(n, D, generator=g), dim=1)
    return dict(logits_u_w=2 * torch.randn(Bu, C, generator=g), logits_u_s0=2 * torch.randn(Bu, C, generator=g), feats_u_w=nf(Bu),
                feats_u_s0=nf(Bu), feats_u_s1=nf(Bu), feats_x=nf(B), targets_x=torch.randint(0, C, (B,), generator=g))


def _make_head_class():
    import torch.nn.functional as F

    from endoscopy_image_classification_b200.bank import local_segments
    from endoscopy_image_classification_b200.comatch_head import CoMatchHead
    from oracle import ssl_oracle as O

    class OracleBackedHead(CoMatchHead):
        def _check_backend(self):      # CPU tensors are fine for this test double
            pass

        def _k_da(self, lw):
            hist = self.prob_list
            O.comatch_da(lw, hist, self.da_window)
            self.prob_list = hist
            self.prob_avg.copy_(torch.stack(hist).mean(0))

        def _k_smooth(self, queries, packed_ld=0):
            A = torch.exp(queries @ self.queue_feats.t() / self.temperature)
            if not packed_ld:
                return A.sum(1), A @ self.queue_probs
            packed = torch.zeros(queries.shape[0], packed_ld)
            packed[:, :self.num_classes] = A @ self.queue_probs
            packed[:, self.num_classes] = A.sum(1)
            return packed

        def _k_finalize(self, lw, ls0, rowsum, numer, lds=(0, 0)):
            p = torch.softmax(lw, 1) / self.prob_avg
            po = p / p.sum(1, keepdim=True)
            rows, C = lw.shape
            if rowsum is not None:
                numer, rowsum = numer[:rows, :C], rowsum[:rows, 0] if rowsum.dim() == 2 else rowsum[:rows]
            probs = self.alpha * po + (1 - self.alpha) * numer / rowsum[:, None] if rowsum is not None else po
            scores, lbs = probs.max(1)
            mask = scores.ge(self.thr).float()
            with torch.enable_grad():
                s = ls0.clone().requires_grad_(True)
                lu = O.comatch_focal_softce(s, probs, mask, self.gamma)
                lu.backward()
            return {"probs": probs, "probs_orig": po, "scores": scores, "mask": mask, "lbs": lbs, "grad_s0": s.grad,
                    "scalars": torch.tensor([float(lu), float(mask.mean()), 0.0, 0.0]), "probs_hl": None}

        def _k_enqueue(self, fw, fx, probs_orig, tx, block_offset, advance):
            rows_f = torch.cat([fw, fx])
            rows_p = torch.cat([probs_orig, F.one_hot(tx, self.num_classes).float().reshape(-1, self.num_classes)])
            ptr = int(self.ptr_state[0])
            for src, dst, ln in local_segments((ptr + block_offset) % self.queue_size, rows_f.shape[0], self.geom):
                self.queue_feats[dst:dst + ln] = rows_f[src:src + ln]
                self.queue_probs[dst:dst + ln] = rows_p[src:src + ln]
            if advance:
                self.ptr_state[0] = (ptr + advance) % self.queue_size

        def _k_contrast_fwd(self, fs0, fs1, probs, scalars, lambda_u=1.0, lambda_c=1.0, probs_hl=None):
            lc = O.comatch_contrast(fs0, fs1, probs, self.temperature, self.contrast_th)
            scalars[2] = lc
            scalars[3] = lambda_u * scalars[0] + lambda_c * lc
            return torch.zeros(3, fs0.shape[0]), lc

        def _k_contrast_bwd(self, f0, f1, probs, stats, g_c, factor=1.0, probs_hl=None, scale=None):
            with torch.enable_grad():
                a, b = f0.clone().requires_grad_(True), f1.clone().requires_grad_(True)
                O.comatch_contrast(a, b, probs, self.temperature, self.contrast_th).backward()
            if scale is not None:
                scale[0].mul_(scale[1] * scale[2])
            return a.grad * g_c * factor, b.grad * g_c * factor

        def _k_scale(self, grad, g, factor=1.0):
            return grad * g * factor

    return OracleBackedHead


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(REPO))
    sys.path.insert(0, str(REPO / "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    Head = _make_head_class()
    head = Head(C, D, K, THR, enqueue_mode="always", device="cpu", process_group=dist.group.WORLD)
    head.fuse_rows = False             # the test double replaces the three separate row kernels
    assert head.geom.shard_rows == K // world and head.queue_feats.shape == (K // world, D)
    outs = []
    for step in range(STEPS):
        inp = _inputs(100 * step + rank)
        for k in ("logits_u_s0", "feats_u_s0", "feats_u_s1"):
            inp[k].requires_grad_(True)
        total, lu, lc, mm = head.total_loss(**inp, lambda_u=2.0, lambda_c=0.5)
        total.backward()
        outs.append(dict(total=total.detach(), probs=head.last["probs"], mask=head.last["mask"],
                         g_s0=inp["logits_u_s0"].grad, g_f0=inp["feats_u_s0"].grad, g_f1=inp["feats_u_s1"].grad,
                         ptr=head.queue_ptr, dev_ptr=int(head.ptr_state[0])))
    torch.save(dict(outs=outs, qf=head.queue_feats, qp=head.queue_probs), os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(300)
def test_two_rank_sharded_bank_equals_single_process_oracle(tmp_path):
    sys.path.insert(0, str(REPO))
    from oracle import ssl_oracle as O
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(tmp_path / f"rank{r}.pt", weights_only=False) for r in range(world)]
    state = O.CoMatchState.zeros(K, D, C)
    hist = [[] for _ in range(world)]
    n = B + B * MU
    for step in range(STEPS):
        inputs = [_inputs(100 * step + r) for r in range(world)]
        ref = O.comatch_head_sharded(state, hist, inputs, thr=THR, num_classes=C)
        for r in range(world):
            got = res[r]["outs"][step]
            torch.testing.assert_close(got["probs"], ref[r]["probs"], rtol=1e-5, atol=1e-6)
            assert torch.equal(got["mask"], ref[r]["mask"])
            want = 2.0 * ref[r]["loss_u"] + 0.5 * ref[r]["loss_contrast"]
            torch.testing.assert_close(got["total"], want, rtol=1e-5, atol=1e-6)
            torch.testing.assert_close(got["g_s0"], 2.0 * ref[r]["grad_logits_s0"], rtol=1e-4, atol=1e-7)
            torch.testing.assert_close(got["g_f0"], 0.5 * ref[r]["grad_feats_s0"], rtol=1e-4, atol=1e-7)
            torch.testing.assert_close(got["g_f1"], 0.5 * ref[r]["grad_feats_s1"], rtol=1e-4, atol=1e-7)
            assert got["ptr"] == got["dev_ptr"] == state.queue_ptr == ((step + 1) * world * n) % K
    bank_f = torch.cat([res[r]["qf"] for r in range(world)])
    bank_p = torch.cat([res[r]["qp"] for r in range(world)])
    assert torch.equal(bank_f, state.queue_feats)          # rank-major rows, wrapped ring: bit-exact copies
    torch.testing.assert_close(bank_p, state.queue_probs, rtol=1e-5, atol=1e-6)
