"""CPU: host-side logic, the C-ABI surface, and the multi-process (gloo, world size 2) path."""
import ctypes
import os
import random
import re
import types

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "pcrl_b200.h")).read()
    return re.findall(r"^\s*(?:int|const char\*)\s+(pcrl_\w+)\s*\(", src, flags=re.M)


def test_library_exports_every_declared_symbol():
    from pcrlv2_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)       # loads without a GPU; no compute call is made here
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/pcrl_b200.h but not exported"
    # the python binding table covers exactly the compute entry points of the header
    assert set(_lib.SIGNATURES) == set(syms) - {"pcrl_last_error", "pcrl_version"}
    lib.pcrl_version.restype = ctypes.c_int
    assert lib.pcrl_version() >= 100


def test_header_cites_reference_lines():
    src = open(os.path.join(ROOT, "include", "pcrl_b200.h")).read()
    assert src.count("pcrlv2_model_3d.py:") >= 6 and "train_3d.py:48-51" in src


def test_product_path_fails_loudly_without_gpu():
    from pcrlv2_b200.models import PCRLv23d
    from pcrlv2_b200 import _lib
    m = PCRLv23d()
    with pytest.raises(RuntimeError, match="CUDA only"):
        m(torch.zeros(2, 1, 16, 16, 16))
    with pytest.raises(_lib.PcrlError, match="CPU tensor"):
        _lib.call("pcrl_zero_pad_rows", torch.zeros(8), 1, 1, 8)
    # the head / loss functions are CUDA kernels as well: no silent CPU evaluation
    from pcrlv2_b200 import functional as Fn
    with pytest.raises(_lib.PcrlError, match="CPU tensor"):
        Fn.mse_loss(torch.zeros(4, 4), torch.zeros(4, 4))
    with pytest.raises(_lib.PcrlError, match="CPU tensor"):
        Fn.cosine_mean(torch.randn(4, 8), torch.randn(4, 8))
    with pytest.raises(_lib.PcrlError, match="CPU tensor"):
        Fn.batch_norm1d(torch.randn(4, 8), torch.nn.BatchNorm1d(8))
    # nothing in the product package imports the oracle
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pcrlv2_b200")):
        for f in files:
            if f.endswith(".py"):
                assert "oracle" not in open(os.path.join(dirpath, f)).read().replace("# oracle", ""), f


def test_model_surface_matches_reference_layout():
    from oracle import pcrlv2_oracle as orc
    from pcrlv2_b200.models import PCRLv23d
    for norm, act in (("bn", "relu"), ("in", "relu"), ("bn", "prelu"), ("bn", "elu")):
        m = PCRLv23d(norm=norm, act=act)
        sd = m.state_dict()
        spec = orc.state_spec(norm=norm, act=act)
        assert [k for k, _, _ in spec] == list(sd.keys())
        for k, shape, _ in spec:
            assert tuple(sd[k].shape) == tuple(shape), k
    # the reference cannot build norm='gn' either (GroupNorm(8, 1) raises ValueError)
    with pytest.raises((NotImplementedError, ValueError)):
        PCRLv23d(norm="gn")
    with pytest.raises(ValueError):
        PCRLv23d(norm="xx")
    with pytest.raises(ValueError):
        PCRLv23d(act="leaky")
    import inspect
    sig = inspect.signature(PCRLv23d.__init__)
    # the reference's arguments in the reference's order; `precision` is an added keyword
    assert list(sig.parameters)[1:7] == ["n_class", "act", "norm", "in_channels", "low_dim", "student"]
    # same RNG consumption as the reference construction order -> load_state_dict round trip
    m = PCRLv23d()
    m.load_state_dict(orc.init_state(3))


def test_cos_loss_and_schedule():
    from oracle import pcrlv2_oracle as orc
    from pcrlv2_b200 import train_3d as T
    from pcrlv2_b200.utils import AverageMeter, adjust_learning_rate
    torch.manual_seed(0)
    a = [[torch.randn(4, 8), torch.randn(4, 8)] for _ in range(3)]
    b = [[torch.randn(4, 8), torch.randn(4, 8)] for _ in range(3)]
    random.seed(11)
    l1, i1 = T.cos_loss(torch.nn.CosineSimilarity(), a, b)
    l2, i2 = orc.cos_loss(random.Random(11), a, b)
    assert i1 == i2 and torch.allclose(l1, l2)
    opt = types.SimpleNamespace(param_groups=[{"lr": 0.0}])
    args = types.SimpleNamespace(lr=1e-3, epochs=240)
    for e in (0, 60, 240):
        adjust_learning_rate(e, args, opt)
        assert abs(opt.param_groups[0]["lr"] - orc.lr_at(e, 1e-3, 240)) < 1e-15
    m = AverageMeter()
    m.update(1.0, 2)
    m.update(4.0, 1)
    assert m.avg == 2.0 and m.val == 4.0 and m.count == 3


def test_cli_flags_match_reference():
    from pcrlv2_b200.main import build_parser
    ns = build_parser().parse_args([])
    for k, v in dict(model="pcrlv2", phase="pretask", b=16, epochs=100, lr=1e-3, n="luna", d=3, workers=4,
                     gpus="0,1,2,3", ratio=0.8, momentum=0.9, weight_decay=1e-4, seed=42, amp=False).items():
        assert getattr(ns, k) == v, k
    ns = build_parser().parse_args("--b 32 --epochs 240 --lr 1e-3 --n luna --d 3 --gpus 0,1,2,3 --ratio 1.0 --amp".split())
    assert ns.b == 32 and ns.amp and ns.ratio == 1.0


def test_synthetic_batch_contract():
    from pcrlv2_b200.data import DataGenerator
    args = types.SimpleNamespace(data="synthetic", b=4, workers=0, seed=42, synthetic_items=8)
    loader = DataGenerator(args).pcrlv2_luna_pretask()["train"]
    x1, x2, gt, gt2, local = next(iter(loader))
    assert tuple(x1.shape) == (4, 1, 64, 64, 32) == tuple(gt.shape)
    assert len(local) == 6 and tuple(local[0].shape) == (4, 1, 16, 16, 16)
    assert 0 <= gt.min() and gt.max() < 1


def _ddp_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pcrlv2_b200.train_3d import allreduce_flat_gradients
    # every rank holds the gradient of ITS shard (mean over the shard); the exchange step must
    # yield the gradient of the global mean on every rank
    g = torch.full((1000,), float(rank + 1))
    scale = allreduce_flat_gradients(g, None)
    torch.testing.assert_close(g * scale, torch.full((1000,), 1.5))
    # the 13 scale draws of a step must agree on all ranks (same seed)
    random.seed(42)
    draws = torch.tensor([random.randint(0, 2) for _ in range(13)])
    gathered = [torch.zeros_like(draws) for _ in range(world)]
    dist.all_gather(gathered, draws)
    assert all(torch.equal(gathered[0], t) for t in gathered)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        open(out, "w").write("ok")


def test_two_rank_gradient_exchange_gloo(tmp_path):
    out = str(tmp_path / "ok.txt")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_ddp_worker, args=(2, port, out), nprocs=2, join=True)
    assert open(out).read() == "ok"


def test_bench_reference_arm_prints_one_json_line():
    """bench.py --impl reference (the CPU arm): exactly one line on stdout, valid JSON with the keys
    of the contract, whatever libraries print while it runs."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "volumes/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["metric"].startswith("LUNA 64x64x32")


def test_c_abi_argument_errors_without_gpu():
    """Error behaviour of the C ABI (include/pcrl_b200.h): bad arguments are rejected with
    PCRL_ERR_ARG (-1) and a message BEFORE any CUDA call, so this runs without a GPU."""
    from pcrlv2_b200 import _lib
    lib = _lib.lib()
    P = ctypes.c_void_p
    # NULL operand
    assert lib.pcrl_linear_fwd(None, None, None, None, 4, 8, 8, None) == -1
    assert b"NULL" in lib.pcrl_last_error()
    assert lib.pcrl_mse_fwd(None, None, None, 16, None) == -1
    fake = P(4096)      # never dereferenced: the shape checks fire first
    # weight gradient needs Cout % 64 == 0
    assert lib.pcrl_conv3d_k3_wgrad(fake, fake, fake, 1, 4, 4, 4, 64, 48, 0, None) == -1
    assert b"multiple of 64" in lib.pcrl_last_error()
    # unknown storage type
    assert lib.pcrl_conv3d_k3_fprop(fake, fake, fake, None, 0, 0, 1, 4, 4, 4, 64, 64, 7, None) == -1
    assert b"dtype" in lib.pcrl_last_error()
    # BatchNorm1d in training mode needs more than one row (torch raises the same condition)
    assert lib.pcrl_bn1d_fwd(fake, fake, fake, None, None, None, fake, fake, fake, 1, 8, 0, 1, 0.1, 1e-5, None) == -1
    assert b"more than 1 value per channel" in lib.pcrl_last_error()
    # norm/act channel count must be 8 * 2^k
    assert lib.pcrl_norm_act_fwd(fake, fake, fake, None, fake, None, None, 0, 0, 0, 1, 4, 4, 4, 24, 0, None) == -1
    assert b"8 * 2^k" in lib.pcrl_last_error()


# ------------------------------------------------------------------------------ 2-D path, host side
def test_model2d_surface_matches_reference_layout():
    from oracle import pcrlv2_oracle_2d as orc2
    from pcrlv2_b200.models import PCRLv2
    m = PCRLv2()
    sd = m.state_dict()
    spec = orc2.state_spec()
    assert [k for k, _, _ in spec] == list(sd.keys())            # = the reference PCRLv2().state_dict() (make_golden_2d.py)
    for k, shape, _ in spec:
        assert tuple(sd[k].shape) == tuple(shape), k
    m.load_state_dict(orc2.init_state(1))
    # the checkpoint of train_2d.py:99 is the ENCODER's state_dict: torchvision resnet18 minus fc
    import torchvision
    ref = torchvision.models.resnet18()
    del ref.fc
    ref.load_state_dict(m.model.encoder.state_dict(), strict=True)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(2, 3, 32, 32))
    with pytest.raises(NotImplementedError):
        PCRLv2(n_class=1)
    with pytest.raises(ValueError):
        PCRLv2(precision="fp16")


def test_reached_parameter_rule_2d_matches_the_oracles_autograd():
    """note N3 for the 2-D model: GraphedStep2d._reached(draws) (what the captured step tells SGD) against the set of
    parameters that actually receive a gradient in the oracle's autograd graph, for several draw patterns."""
    from oracle import pcrlv2_oracle_2d as orc2
    from pcrlv2_b200.models import PCRLv2
    from pcrlv2_b200.train_2d import GraphedStep2d

    class FixedDraws:
        def __init__(self, seq):
            self.seq = list(seq)

        def randint(self, a, b):
            v = self.seq.pop(0)
            assert a <= v <= b
            return v

    m = PCRLv2()
    names = [n for n, _ in m.named_parameters()]
    fake = types.SimpleNamespace(model=m, opt=types.SimpleNamespace(_ps=[p for _, p in m.named_parameters()]))
    sd0 = orc2.init_state(0)
    x1, x2, gt, lv = orc2.synthetic_batch(2, seed=1, size=(32, 32), local=(32, 32), n_local=2)
    for draws in ([0, 0, 0, 0, 0], [3, 1, 1, 4, 1], [4, 2, 0, 2, 0], [2, 2, 2, 2, 2]):
        sd = orc2.clone_state(sd0)
        keys = [k for k in sd if orc2.is_param(k)]
        for k in keys:
            sd[k].requires_grad_(True)
        loss, _, got = orc2.step_loss(sd, x1, x2, gt, lv, 0, FixedDraws(draws))
        assert got == draws
        grads = torch.autograd.grad(loss, [sd[k] for k in keys], allow_unused=True)
        reached_ref = {k: g is not None for k, g in zip(keys, grads)}
        rule = dict(zip(names, GraphedStep2d._reached(fake, draws)))
        assert rule == reached_ref, [k for k in names if rule[k] != reached_ref[k]][:8]
        fake.__dict__.pop("_names", None)


def test_synthetic_chest_batch_contract_and_cli_dispatch():
    from pcrlv2_b200.data import DataGenerator
    args = types.SimpleNamespace(data="synthetic", b=4, workers=0, seed=42, synthetic_items=8)
    loader = DataGenerator(args).pcrlv2_chest_pretask()["train"]
    x1, x2, g1, g2, lv = next(iter(loader))
    assert x1.shape == (4, 3, 224, 224) and g2.shape == x2.shape and len(lv) == 6 and lv[0].shape == (4, 3, 96, 96)
    assert float(g1.min()) >= 0.0 and float(g1.max()) < 1.0
    from pcrlv2_b200 import main as M
    assert M.train_pcrlv2.__module__ == "pcrlv2_b200.train_2d"      # main.py:47-48 dispatches --d 2 to train_2d
