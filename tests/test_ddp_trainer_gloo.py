"""SURVEY 8 f3 on the CPU: the data-parallel wiring of the trainer shell in a world-size-2 gloo job -- DistributedDataParallel
around the backbone (gradients averaged in backward), rank-aware samplers with ``set_epoch``, LR schedule index
(``epoch * EVAL_STEP + batch_idx``, reference fixmatch.py:124), checkpoint written by rank 0 only.  The criteria are CUDA
kernels with no CPU path, so on this box they are swapped for the oracle's restatement (test infrastructure); the same
trainer runs the real kernels under NCCL in tests/test_gpu_trainers.py::test_ddp_single_rank_nccl."""
import os
import socket
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

REPO = Path(__file__).resolve().parents[1]
C = 5


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _Views(torch.utils.data.Dataset):
    def __init__(self, n, labeled, seed):
        g = torch.Generator().manual_seed(seed)
        self.x, self.y, self.labeled = torch.randn(n, 8, generator=g), torch.randint(0, C, (n,), generator=g), labeled

    def __len__(self):
        return len(self.x)

    def __getitem__(self, i):
        if self.labeled:
            return self.x[i], self.y[i]
        return (self.x[i], self.x[i] + 0.1), self.y[i]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(REPO))
    sys.path.insert(0, str(REPO / "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from endoscopy_image_classification_b200 import fixmatch as fm
    from endoscopy_image_classification_b200 import utils
    from oracle import ssl_oracle as O
    fm.ce_loss = lambda logits, targets, class_weights=None, reduction="mean", type_loss="none": O.ce_loss(
        logits, targets, class_weights, reduction=reduction, type_loss=type_loss)
    fm.consistency_loss = lambda w, s, T=1.0, p_cutoff=0.0, device=None: O.consistency_loss(w, s, T=T, p_cutoff=p_cutoff)
    torch.manual_seed(0)                                 # identical replicas
    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.body, self.fc = nn.Linear(8, 16), nn.Linear(16, C)

        def forward(self, x):
            return self.fc(torch.relu(self.body(x)))
    net = Net()
    A = utils.AttrDict
    cfg = A(DATA=A(BATCH_SIZE=4, MU=2, TARGET_NAME="t"), MODEL=A(NUM_CLASSES=C, NAME="x"),
            TRAIN=A(EPOCHS=4, WARMUP_EPOCHS=1, DECAY_EPOCHS=1, WARMUP_LR=0.01, SCH_NAME="cosine", BASE_LR=0.1, EVAL_STEP=3,
                    USE_EMA=False, EMA_DECAY=0.99, IS_FREEZE=False, CLS_WEIGHT=False, THRES=0.3, T=1.0, LAMBDA_U=1.0,
                    FREQ_EVAL=1, SAVE_CP=os.path.join(out_dir, "ckpt")))
    tr = fm.FixMatch(net, opt_func="SGD", device="cpu")
    assert tr.rank == rank and tr.world_size == world
    lab = tr.distributed_loader(_Views(32, True, 1), batch_size=4)
    unl = tr.distributed_loader(_Views(64, False, 2), batch_size=8)
    tr.get_dataloader((lab, unl), [])
    tr.get_config(cfg)
    from torch.nn.parallel import DistributedDataParallel
    assert isinstance(tr.net, DistributedDataParallel) and tr.net.module is tr.model and tr.model is net
    seen = []
    orig = tr.lr_scheduler.step_update
    tr.lr_scheduler.step_update = lambda i: (seen.append(i), orig(i))[1]
    w0 = [p.detach().clone() for p in net.parameters()]
    meter = tr.train_one(epoch=1)
    assert seen == [3, 4, 5]                              # epoch * EVAL_STEP + batch_idx
    del tr.lr_scheduler.step_update                       # the instance attribute; state_dict() must stay picklable
    assert lab.sampler.epoch == 1 and unl.sampler.epoch == 1
    # the ranks saw different data (disjoint sampler shards) but hold identical weights after the averaged steps
    mine = torch.cat([p.detach().flatten() for p in net.parameters()])
    both = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(both, mine)
    assert torch.equal(both[0], both[1]) and not torch.equal(mine, torch.cat([p.flatten() for p in w0]))
    idx = torch.tensor(list(iter(lab.sampler)))
    other = [torch.empty_like(idx) for _ in range(world)]
    dist.all_gather(other, idx)
    assert not set(other[0].tolist()) & set(other[1].tolist())
    tr.epoch = 1
    path = tr.save_checkpoint(cfg.TRAIN.SAVE_CP)
    assert (path is not None) == (rank == 0)
    files = os.listdir(cfg.TRAIN.SAVE_CP)
    assert len(files) == 1                                # after the barrier: exactly rank 0's file
    torch.save({"loss": meter.avg}, os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_ddp_trainer_two_ranks_gloo(tmp_path):
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert all((tmp_path / f"rank{r}.pt").exists() for r in range(2))
