"""CPU tests: the C-ABI library loads and exports every symbol include/b200ssl.h
declares, rejects bad arguments before touching the GPU, and the host-side logic
(EMA block table, shard geometry, config knobs, no-CPU-fallback guarantees)."""
import ctypes as C
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from conftest import REPO

HEADER = REPO / "include" / "b200ssl.h"
PKG = REPO / "endoscopy-image-classification_b200"


@pytest.fixture(scope="module")
def native():
    lib_path = PKG / "libb200ssl.so"
    if not lib_path.exists():
        import __graft_entry__ as g
        g.build()
    from endoscopy_image_classification_b200 import _native
    _native.lib()
    return _native


def declared_symbols():
    text = HEADER.read_text()
    return sorted(set(re.findall(r"B200SSL_API[^;(]*?\b(b200ssl_\w+)\s*\(", text)))


def test_header_symbols_exported_and_bound(native):
    syms = declared_symbols()
    assert len(syms) >= 13
    lib = C.CDLL(str(PKG / "libb200ssl.so"))
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/b200ssl.h but not exported"
    assert sorted(native.SIGNATURES) == syms, "ctypes SIGNATURES out of sync with the header"
    exported = subprocess.run(["nm", "-D", "--defined-only", str(PKG / "libb200ssl.so")], capture_output=True, text=True).stdout
    extra = [l.split()[-1] for l in exported.splitlines() if " T " in l and not l.split()[-1].startswith("b200ssl_")]
    assert not extra, f"unexpected exported symbols: {extra[:5]}"


def test_no_torch_types_in_header():
    text = HEADER.read_text()
    assert "torch" not in text.lower().replace("pytorch", "") and "at::" not in text and "#include <cuda" not in text


def test_version_and_workspace(native):
    lib = native.lib()
    assert lib.b200ssl_version() == 100
    small = lib.b200ssl_workspace_bytes(448, 23, 2560)
    big = lib.b200ssl_workspace_bytes(14336, 23, 65536)
    assert 256 + 65536 < small and 256 + 65536 < big and small % 256 == 0 and big % 256 == 0


def test_argument_validation_without_gpu(native):
    """Bad arguments are rejected with negative codes before any CUDA call."""
    lib = native.lib()
    buf = (C.c_char * 4096)()
    p = C.addressof(buf)
    p256 = (p + 255) & ~255
    wsb = lib.b200ssl_workspace_bytes(16, 23, 0)
    E_NULL, E_SHAPE, E_DTYPE, E_ALIGN, E_WS, E_ARG = -1, -2, -3, -4, -5, -6
    f = lib.b200ssl_fixmatch_head_fwd_bwd
    assert f(None, p, None, p, None, 16, 23, 0, 0.95, 1.0, 1, p, None, None, p256, wsb, None) == E_NULL
    assert b"NULL" in lib.b200ssl_last_error_string()
    assert f(p, p, None, p, None, 0, 23, 0, 0.95, 1.0, 1, p, None, None, p256, wsb, None) == E_SHAPE
    assert f(p, p, None, p, None, 16, 5000, 0, 0.95, 1.0, 1, p, None, None, p256, wsb, None) == E_SHAPE
    assert f(p, p, None, p, None, 16, 23, 9, 0.95, 1.0, 1, p, None, None, p256, wsb, None) == E_DTYPE
    assert f(p, p, p, p, None, 16, 23, 0, 0.95, 1.0, 1, p, None, None, p256, wsb, None) == E_NULL   # s2 without grad_s2
    assert f(p, p, None, p, None, 16, 23, 0, 0.95, 1.0, 1, p, None, None, p256 + 4, wsb, None) == E_ALIGN
    assert f(p, p, None, p, None, 16, 23, 0, 0.95, 1.0, 1, p, None, None, p256, 64, None) == E_WS
    assert lib.b200ssl_comatch_da(p, 16, 23, 0, p, p, 100, p, None, p256, wsb, None) == E_ARG      # window > 64
    assert lib.b200ssl_bank_smooth_partial(p, p, p, None, 16, 64, 12, 23, 0, 0.2, p, p, 0, 0, None, p256, wsb, None) == E_SHAPE  # dim % 8
    assert lib.b200ssl_bank_smooth_partial(p, p, p, None, 16, 64, 64, 23, 0, 0.0, p, p, 0, 0, None, p256, wsb, None) == E_ARG   # tau
    assert lib.b200ssl_bank_enqueue(p, p, None, p, p, p, p, 4, 4, 64, 23, 0, 100, None, 0, 0, 64, 0, 64, None) == E_ARG  # ptr >= K
    assert lib.b200ssl_bank_enqueue(p, p, None, p, p, p, p, 4, 4, 64, 23, 0, 0, None, 0, 0, 64, 32, 64, None) == E_ARG   # shard outside
    assert lib.b200ssl_bank_enqueue(p, p, None, p, p, p, p, 4, 4, 64, 23, 0, 0, None, 8, 0, 64, 0, 64, None) == E_ARG    # advance w/o state
    assert lib.b200ssl_contrast_fwd(p, p, p, None, 16, 64, 500, 0, 0.2, 0.8, p, p, None, 1.0, 1.0, None, p256, wsb, None) == E_SHAPE
    assert lib.b200ssl_ema_multi_tensor(None, 4, 0, 1, 0.999, 0.001, 0, None) == E_NULL
    assert lib.b200ssl_ema_multi_tensor(p256, 4, 3, 1, 0.999, 0.001, 0, None) == E_DTYPE
    assert lib.b200ssl_ema_multi_tensor(p256, 4, 0, 1, 0.999, 0.001, 7, None) == E_ARG
    assert lib.b200ssl_scale_inplace(None, 4, 0, p, 1.0, None) == E_NULL
    # the forms of the update that run next to other kernels (ema.ModelEMA(overlap=True))
    ctas, masked = lib.b200ssl_ema_multi_tensor_ctas, lib.b200ssl_ema_multi_tensor_masked
    assert ctas(None, 4, 0, 1, 0.999, 0.001, 0, 400, None) == E_NULL
    assert ctas(p256, 4, 0, 1, 0.999, 0.001, 0, -1, None) == E_ARG                    # negative grid cap
    assert ctas(p256 + 8, 4, 0, 1, 0.999, 0.001, 0, 400, None) == E_ALIGN
    assert masked(p256, 4, 0, 1, 0.999, 0.001, 0, None, p, None) == E_NULL            # no SM mask
    assert masked(p256, 4, 0, 1, 0.999, 0.001, 0, p, None, None) == E_NULL            # no scheduler words
    assert masked(p256, 0, 0, 1, 0.999, 0.001, 0, p, p, None) == E_SHAPE
    assert masked(p256, 4, 9, 1, 0.999, 0.001, 0, p, p, None) == E_DTYPE
    assert lib.b200ssl_probe_sm_set(6, 8, None, p, None) == E_NULL
    assert lib.b200ssl_probe_sm_set(6, 3, p, p, None) == E_ARG                        # cluster not a power of two
    assert lib.b200ssl_probe_sm_set(12, 8, p, p, None) == E_ARG                       # more than half the SMs
    assert lib.b200ssl_stream_delay(-5, None) == E_ARG
    assert lib.b200ssl_stream_delay(10**9, None) == E_ARG
    assert lib.b200ssl_stream_delay(0, None) == 0                                     # nothing to launch
    # multi-rank bank / peer memory entry points
    out = C.c_void_p()
    assert lib.b200ssl_peer_alloc(16, C.byref(out), p) == E_ARG                      # smaller than the control page
    assert lib.b200ssl_peer_alloc(1 << 20, None, p) == E_NULL
    ctl = lib.b200ssl_peer_control_bytes()
    assert ctl >= 1024 and ctl % 256 == 0
    ag, rs = lib.b200ssl_peer_all_gather, lib.b200ssl_peer_reduce_scatter_f32
    assert ag(None, 64, None, 0, p, p, ctl, 256, 0, 0, 2, None) == E_NULL
    assert ag(p, 64, None, 0, p256, p, ctl, 256, 0, 0, 1, None) == E_ARG             # world < 2
    assert ag(p256, 64, None, 0, p256, p, ctl, 256, 9, 0, 2, None) == E_ARG          # exchange id
    assert ag(p256, 60, None, 0, p256, p, ctl, 256, 0, 0, 2, None) == E_ALIGN        # bytes % 16
    assert ag(p256, 512, None, 0, p256, p, ctl, 256, 0, 0, 2, None) == E_ALIGN       # slot < bytes
    assert rs(p256, p256, 0, p, ctl, 256, 1, 0, 2, None) == E_SHAPE
    assert rs(p256, p256, 16, p, 64, 256, 1, 0, 2, None) == E_ALIGN                  # region inside the control page
    sh = native.BankShards(2, 0, 64, p, p, 4096, 4096 + 8192, 4096 + 16384, 0, 0)
    assert C.sizeof(native.BankShards) == 64
    enq = lib.b200ssl_bank_enqueue_peer
    assert enq(p, p, p, p, 4, 4, 64, 23, 1, p, None, None) == E_NULL                  # no shard table
    assert enq(p, p, p, p, 4, 4, 64, 23, 0, p, C.byref(sh), None) == E_DTYPE         # fp32 bank
    assert enq(p, p, p, p, 4, 4, 64, 23, 1, None, C.byref(sh), None) == E_NULL       # no ptr_state
    assert enq(p256, p256, p, p, 40, 40, 64, 23, 1, p, C.byref(sh), None) == E_SHAPE  # world*n > bank rows
    bad = native.BankShards(1, 0, 64, p, p, 4096, 8192, 16384, 0, 0)
    sm = lib.b200ssl_bank_smooth_partial
    assert sm(p256, None, None, None, 16, 64, 64, 23, 1, 0.2, p, p, 0, 0, C.byref(bad), p256, wsb, None) == E_ARG    # one rank
    assert sm(p256, None, None, None, 16, 100, 64, 23, 1, 0.2, p, p, 0, 0, C.byref(sh), p256, wsb, None) == E_ARG    # K != world*shard
    assert sm(p256, None, None, None, 16, 128, 64, 23, 0, 0.2, p, p, 0, 0, C.byref(sh), p256, wsb, None) == E_DTYPE  # fp32


def test_product_has_no_cpu_fallback_and_no_oracle_import(native):
    from endoscopy_image_classification_b200 import comatch_head, ema, loss
    w, s = torch.randn(8, 23), torch.randn(8, 23)
    with pytest.raises(RuntimeError, match="CUDA"):
        loss.consistency_loss(w, s)
    with pytest.raises(RuntimeError, match="CUDA"):
        loss.ce_loss(w, torch.zeros(8, dtype=torch.long), reduction="mean")
    with pytest.raises(RuntimeError, match="CUDA"):
        comatch_head.CoMatchHead(23, 64, 512, 0.9, device="cpu")
    m = torch.nn.Linear(4, 4)
    with pytest.raises(RuntimeError, match="CUDA"):
        ema.ModelEMA(m, 0.9).update(m)
    # nothing in the product imports the oracle or reads the reference tree
    for f in PKG.rglob("*.py"):
        src = f.read_text()
        assert "oracle" not in src.replace("no reference oracle", "").replace("without reference oracle", ""), f
        assert "/root/reference" not in src, f


def test_missing_library_fails_loudly(tmp_path):
    code = ("import os, sys; os.environ['B200SSL_LIB']=%r; sys.path.insert(0, %r);"
            "from endoscopy_image_classification_b200 import _native as N\n"
            "try:\n    N.lib()\nexcept N.NativeLibraryError as e:\n    print('LOUD', e); sys.exit(0)\nsys.exit(1)") % (
        str(tmp_path / "nope.so"), str(REPO))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0 and "LOUD" in r.stdout, r.stderr[-500:]


def test_ema_block_table():
    from endoscopy_image_classification_b200 import _native as N
    from endoscopy_image_classification_b200.ema import BLOCK_DTYPE, build_block_table
    entries = [(0x1000, 0x9000, 10000, 4, N.F32, 2), (0x20000, 0x30000, 1, 8, N.I64, 1), (0x40000, 0x50000, 4096, 2, N.BF16, 1),
               (0x60000, 0x70000, 0, 4, N.F32, 1)]
    t = build_block_table(entries)
    assert t.dtype == BLOCK_DTYPE and t.dtype.itemsize == 32
    assert t["count"].tolist() == [4096, 4096, 1808, 1, 4096]
    assert t["ema"].tolist() == [0x1000, 0x1000 + 4096 * 4, 0x1000 + 8192 * 4, 0x20000, 0x40000]
    assert t["model"][2] == 0x9000 + 8192 * 4
    assert t["repeat"].tolist() == [2, 2, 2, 1, 1] and t["dtype"].tolist() == [0, 0, 0, 3, 1]
    assert int(t["count"].sum()) == 10000 + 1 + 4096
    assert C.sizeof(N.EmaBlock) == 32


def test_shard_geometry_and_segments():
    from endoscopy_image_classification_b200.bank import ShardGeometry, local_segments
    with pytest.raises(ValueError):
        ShardGeometry(100, 8, 0)
    K, R, n = 96, 4, 16
    covered = np.zeros(K, dtype=int)
    ptr = 80
    for rank in range(R):
        g = ShardGeometry(K, R, rank)
        assert g.shard_rows == 24 and g.shard_begin == 24 * rank
        for src, dst, ln in local_segments(ptr, R * n, g):
            rows = (ptr + src + np.arange(ln)) % K
            assert (rows == g.shard_begin + dst + np.arange(ln)).all()
            covered[rows] += 1
    want = np.zeros(K, dtype=int)
    want[(ptr + np.arange(R * n)) % K] = 1
    assert (covered == want).all()
    g = ShardGeometry(K, R, 1)
    assert g.next_ptr(80, n) == (80 + 64) % 96
    assert g.should_enqueue(n, "always") and not g.should_enqueue(n, "reference")
    assert ShardGeometry(64, 4, 0).should_enqueue(16, "reference")
    with pytest.raises(ValueError):
        g.should_enqueue(100, "always")


def test_config_roundtrip_and_reference_module_names(tmp_path):
    import endoscopy_image_classification_b200 as eic
    from endoscopy_image_classification_b200.utils import AverageMeter, get_config
    cfg = get_config(PKG / "configs" / "comatch_r50_hyperkvasir.yaml")
    assert cfg.TRAIN.THRES == 0.9 and cfg.DATA.MU == 7 and cfg.MODEL.LOW_DIM == 64 and cfg.MODEL.NUM_CLASSES == 23
    assert cfg.TRAIN.EMA_DECAY == 0.999 and cfg.TRAIN.T == 1.0 and cfg.TRAIN.USE_EMA is True
    m = AverageMeter()
    m.update(2.0, 4)
    m.update(4.0, 4)
    assert m.avg == 3.0
    eic.install_as_reference_modules(("loss", "ema", "utils"))
    import ema as ref_named_ema
    import loss as ref_named_loss
    assert ref_named_loss.consistency_loss is eic.loss.consistency_loss
    assert ref_named_ema.ModelEMA.__init__.__code__.co_varnames[:4] == ("self", "model", "decay", "device")
    for n in ("loss", "ema", "utils"):
        sys.modules.pop(n, None)


def test_trainer_shells_and_schedules_on_cpu():
    """Host logic that needs no GPU: optimizer groups, schedules, FixMatch get_config wiring."""
    import torch.nn as nn
    from endoscopy_image_classification_b200 import utils
    from endoscopy_image_classification_b200.fixmatch import FixMatch
    from endoscopy_image_classification_b200.lr_scheduler import build_scheduler
    from endoscopy_image_classification_b200.optimizer import build_optimizer
    net = nn.Sequential(nn.Linear(4, 8), nn.BatchNorm1d(8), nn.Linear(8, 3))
    opt = build_optimizer(net, "sgd", lr=0.1)
    assert len(opt.param_groups) == 2 and opt.param_groups[1]["weight_decay"] == 0.0
    assert len(opt.param_groups[0]["params"]) == 2 and len(opt.param_groups[1]["params"]) == 4   # weights | biases + BN
    assert build_optimizer(net, "nope") is None
    A = utils.AttrDict
    cfg = A(DATA=A(BATCH_SIZE=4, MU=2), MODEL=A(NUM_CLASSES=3, NAME="x"),
            TRAIN=A(EPOCHS=10, WARMUP_EPOCHS=2, DECAY_EPOCHS=3, WARMUP_LR=0.01, SCH_NAME="cosine", LR_DECAY=0.5, BASE_LR=0.1,
                    EVAL_STEP=5, USE_EMA=True, EMA_DECAY=0.99, IS_FREEZE=False, CLS_WEIGHT=False, THRES=0.95, T=1.0, LAMBDA_U=1))
    sch = build_scheduler(cfg, opt, 5)
    sch.step_update(0)
    assert abs(opt.param_groups[0]["lr"] - 0.01) < 1e-12
    sch.step_update(10)
    assert abs(opt.param_groups[0]["lr"] - (5e-6 + 0.5 * (0.1 - 5e-6) * (1 + np.cos(np.pi * 10 / 50)))) < 1e-9
    sch.step_update(50)
    assert opt.param_groups[0]["lr"] == 5e-6
    cfg.TRAIN.SCH_NAME = "step"
    s2 = build_scheduler(cfg, build_optimizer(net, "adam", lr=0.1), 5)
    s2.step_update(31)
    assert abs(s2.optimizer.param_groups[0]["lr"] - 0.1 * 0.5 ** 2) < 1e-12
    tr = FixMatch(net, opt_func="Adam", device="cpu")
    tr.get_dataloader(([], []), [])
    tr.get_config(cfg)
    assert tr.ema_model.decay == 0.99 and not tr.ema_model.ema.training and tr.class_weights is None
    assert tr.epoch_start == 1 and tr.best_valid_perf is None


def test_peer_arena_layout_with_stubbed_allocation(native, monkeypatch):
    """Region / slot / named-area arithmetic of PeerArena (pure host logic); the allocation and the process group are
    stand-ins, the kernels are not called."""
    import torch.distributed as dist
    from endoscopy_image_classification_b200 import peer

    class Lib:
        def b200ssl_peer_control_bytes(self):
            return 4096

        def b200ssl_peer_alloc(self, nbytes, out, handle):
            C.cast(out, C.POINTER(C.c_void_p))[0] = 0x7F0000000000
            self.nbytes = nbytes
            return 0

    lib = Lib()
    monkeypatch.setattr(native, "lib", lambda: lib)
    monkeypatch.setattr(torch.cuda, "device", lambda d: __import__("contextlib").nullcontext())
    monkeypatch.setattr(torch.cuda, "synchronize", lambda d=None: None)
    monkeypatch.setattr(dist, "get_rank", lambda pg=None: 1)
    monkeypatch.setattr(dist, "get_world_size", lambda pg=None: 4)
    monkeypatch.setattr(dist, "barrier", lambda group=None: None)

    def fake_gather(out, obj, group=None):
        out[:] = [obj] * len(out)
    monkeypatch.setattr(dist, "all_gather_object", fake_gather)

    class OpenLib(Lib):
        opened = 0

        def b200ssl_peer_open(self, handle, out):
            OpenLib.opened += 1
            C.cast(out, C.POINTER(C.c_void_p))[0] = 0x7E0000000000 + OpenLib.opened * (1 << 32)
            return 0
    lib = OpenLib()
    a = peer.PeerArena(None, "cpu", {0: 60288, 2: 1000}, named={"qf": 320 * 128, "qpt": 32 * 320 * 2})
    assert a.rank == 1 and a.world == 4 and OpenLib.opened == 3
    assert a.slot == {0: 60416, 2: 1024}                                   # rounded up to 256 bytes
    assert a.offset[0] == 4096 and a.offset[2] == 4096 + 2 * 4 * 60416     # [2 parities][world][slot] per exchange
    end = a.offset[2] + 2 * 4 * 1024
    assert a.named_offset == {"qf": end, "qpt": end + 320 * 128} and a.bytes == lib.nbytes == end + 320 * 128 + 32 * 320 * 2
    assert a.fits(0, 60288) and not a.fits(0, 60417) and not a.fits(1, 16)
    assert int(a.bases[1]) == 0x7F0000000000 and len(set(a.bases.tolist())) == 4      # own base at index `rank`
    with pytest.raises(ValueError):
        a.all_gather(0, [torch.zeros(3, 5)])                               # 60 bytes: not a multiple of 16
    with pytest.raises(ValueError):
        a.reduce_scatter(2, torch.zeros(8, 3, dtype=torch.float64))


# ---- tensor-core K3: launch plan and the FMA-pipe exponential (host-side restatements) ---------------------------
def test_k3_launch_plan_invariants(native):
    """csrc/bank_tc.cu smooth_tc_plan: every plan splits the key tiles over at most one CTA per SM and wave, a directly
    addressed sharded bank with <= 4 row tiles uses the row loop (each remote key tile crosses NVLink once), and the
    workspace the sizing call reports covers the plan's outer partials."""
    lib = native.lib()
    out = (C.c_int32 * 3)()
    for rows in (1, 100, 448, 512, 896, 1792, 3000, 3584, 7168, 14336):
        for K in (8, 80, 2560, 4104, 8192, 16384, 20480, 32768, 65536):
            for remote in (0, 1):
                assert lib.b200ssl_debug_smooth_plan(rows, K, remote, out) == 0
                mt, cl, no = list(out)
                row_tiles, ktiles = (rows + 127) // 128, (K + 127) // 128
                assert 1 <= mt <= 4 and cl in (1, 2, 4, 8) and no >= 1
                assert cl * no <= ktiles, (rows, K, remote, mt, cl, no)          # every CTA owns at least one key tile
                if remote and row_tiles <= 4:
                    assert mt == row_tiles
                    assert no <= lib.b200ssl_debug_max_active_clusters(cl)         # one group of row tiles: one wave of clusters
                groups = (row_tiles + mt - 1) // mt
                need = 256 + 65536 + (no * groups * mt * 128 * 24 * 4 if no > 1 else 0)
                assert lib.b200ssl_workspace_bytes(rows, 23, K) >= need, (rows, K, remote)
    # BASELINE cfg 4 on 8 ranks (448 queries per rank, 65536 rows): the whole chip shares the key tiles
    # (a cluster lives inside one GPC: B200 holds 15 clusters of 8 one-CTA-per-SM blocks at once, not 148 / 8 = 18)
    assert [lib.b200ssl_debug_max_active_clusters(c) for c in (1, 2, 4, 8)] == [148, 74, 33, 15]    # the no-device table (= what a B200 reports)
    lib.b200ssl_debug_smooth_plan(448, 65536, 1, out)
    mt, cl, no = list(out)
    assert mt == 4 and 100 <= cl * no <= 148 and no <= lib.b200ssl_debug_max_active_clusters(cl)
    # the sweep corner: no half-empty second wave (round 1 ran 224 CTAs on 148 SMs)
    lib.b200ssl_debug_smooth_plan(3584, 65536, 0, out)
    mt, cl, no = list(out)
    assert ((28 + mt - 1) // mt) * no <= lib.b200ssl_debug_max_active_clusters(cl)


def test_head_sm_budget_shapes_the_k3_plan(native):
    """b200ssl_set_head_sm_budget (set by ModelEMA(overlap=True)): the K3 planner keeps to that many CTAs while the
    constrained plan costs at most 1.5 x the free one; big problems and shards read over NVLink ignore it."""
    lib = native.lib()
    out = (C.c_int32 * 3)()

    def ctas(rows, K, remote=0):
        assert lib.b200ssl_debug_smooth_plan(rows, K, remote, out) == 0
        mt, cl, no = list(out)
        return ((rows + 127) // 128 + mt - 1) // mt * cl * no

    assert lib.b200ssl_set_head_sm_budget(0) == 0
    free_n8, free_big, free_remote = ctas(448, 20480), ctas(3584, 65536), ctas(448, 65536, 1)
    assert free_n8 > 48                                          # cfg 2's ring at N = 8: the free plan spreads over the chip
    try:
        assert lib.b200ssl_set_head_sm_budget(48) == 0            # returns the previous value
        assert ctas(448, 2560) <= 48 and ctas(448, 20480) <= 48   # the head stays inside the SMs the capped update leaves free
        assert ctas(3584, 65536) == free_big                      # 28 row tiles x 512 key tiles: a 48-CTA plan would cost > 1.5 x
        assert ctas(448, 65536, 1) == free_remote                 # shards over NVLink live off the tile streams in flight
        assert lib.b200ssl_set_head_sm_budget(1000) == 48         # out of range = no budget
        assert ctas(448, 20480) == free_n8
    finally:
        lib.b200ssl_set_head_sm_budget(0)


def test_k3_polynomial_exp2_math():
    """ex2_poly of csrc/bank_tc.cu restated in numpy fp32: round-to-nearest split through 1.5 * 2^23, degree-3 minimax
    polynomial, exponent patched in with an integer shift-add.  Relative error <= 7.6e-5 (bf16 half-ulp: 2e-3) over the
    whole logit range of unit-norm embeddings at temperature 0.2, and a finite tiny value below the clamp."""
    def fma(a, b, c):
        return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)

    def ex2_poly(s, scale, s_min):
        magic = np.float32(12582912.0)
        s = np.maximum(s, s_min)
        sc = np.full_like(s, scale)
        t = fma(s, sc, np.full_like(s, magic))
        f = fma(s, sc, -(t - magic).astype(np.float32))
        c3, c2, c1, c0 = (np.float32(float.fromhex(h)) for h in ("0x1.c3f76p-5", "0x1.f0de1ap-3", "0x1.62f31ap-1", "0x1.fff692p-1"))
        p = fma(np.full_like(s, c3), f, np.full_like(s, c2))
        p = fma(p, f, np.full_like(s, c1))
        p = fma(p, f, np.full_like(s, c0))
        bits = (p.view(np.uint32).astype(np.uint64) + ((t.view(np.uint32).astype(np.uint64) << 23) & 0xFFFFFFFF)) & 0xFFFFFFFF
        return bits.astype(np.uint32).view(np.float32)

    src = (PKG / "csrc" / "bank_tc.cu").read_text()
    for h in ("0x1.c3f76p-5f", "0x1.f0de1ap-3f", "0x1.62f31ap-1f", "0x1.fff692p-1f", "12582912.f"):
        assert h in src, f"the kernel's constant {h} changed: update this restatement"
    scale = np.float32(1.4426950408889634 / 0.2)
    s_min = np.float32(-126.0 / scale)
    s = np.linspace(-1.2, 1.2, 400001).astype(np.float32)
    got = ex2_poly(s, scale, s_min).astype(np.float64)
    ref = np.exp2(s.astype(np.float64) * float(scale))
    assert np.abs(got / ref - 1).max() < 7.6e-5
    low = ex2_poly(np.array([-30.0, -1e4], np.float32), scale, s_min)
    assert np.all(np.isfinite(low)) and np.all(low >= 0) and np.all(low < 1e-37)
