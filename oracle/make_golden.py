"""Pins oracle/pcrlv2_oracle.py against the reference and writes tests/golden/*.npz.

Run in the BUILD container only (it reads /root/reference, which does not exist on the GPU
box):  python oracle/make_golden.py

What it does
  1. imports the reference model by file path and checks the oracle's state layout against
     ``PCRLv23d().state_dict()`` (keys, shapes, dtypes, order);
  2. forward parity (train mode, BN buffers included) oracle vs reference at b=2, 64x64x32 and
     at b=12, 16^3 with local=True, for norm='bn' (default) and norm='in';
  3. runs the REAL reference ``train_3d.train_pcrlv2_inner`` for two iterations on CPU
     (``.cuda()``/``torch.cuda.synchronize`` patched to no-ops, smp stubbed because
     models/__init__.py imports the 2-D model) and checks the oracle's ``train_step`` reproduces
     its parameters, momentum buffers, BN buffers and loss meters;
  4. writes digests of all of the above as golden fixtures;
  2b. the same forward + restoration-gradient check for act='prelu' and act='elu' (32x32x16);
  3b. the real reference trainer again at a batch where BatchNorm1d is well conditioned (b=16,
      32x32x16 volumes, 6 local 16^3 views, two iterations, lr 1e-2): train_2steps_b16.npz carries
      every loss term of both steps, the final parameters / momentum buffers (1024 samples per tensor)
      and the BN buffers -- the fixture tests/test_step_gpu.py asserts per-parameter updates against.
      ``python oracle/make_golden.py b16`` runs this part alone.
"""
from __future__ import annotations

import importlib.util
import os
import random
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
sys.path.insert(0, ROOT)
from oracle import pcrlv2_oracle as orc  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
N_SAMPLES = 256


def digest(t: torch.Tensor, n_samples: int = N_SAMPLES) -> np.ndarray:
    """[sum, abs-sum, sq-sum, numel, <=n_samples strided samples] in float64."""
    f = t.detach().double().flatten()
    stride = max(1, f.numel() // n_samples)
    samp = f[::stride][:n_samples]
    head = torch.tensor([f.sum(), f.abs().sum(), (f * f).sum(), float(f.numel())], dtype=torch.float64)
    return torch.cat([head, samp]).numpy()


def load_ref_model_module():
    spec = importlib.util.spec_from_file_location("ref_pcrlv2_model_3d",
                                                  os.path.join(REF, "models", "pcrlv2_model_3d.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_ref_train_module():
    """Import the reference train_3d.py unmodified.  models/__init__.py pulls in the 2-D model,
    which needs segmentation_models_pytorch (absent): stub it (never called on the 3-D path)."""
    for name in ["segmentation_models_pytorch", "segmentation_models_pytorch.base",
                 "segmentation_models_pytorch.base.modules",
                 "segmentation_models_pytorch.base.initialization"]:
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
    sys.modules["segmentation_models_pytorch.base"].modules = sys.modules[
        "segmentation_models_pytorch.base.modules"]
    sys.modules["segmentation_models_pytorch"].base = sys.modules["segmentation_models_pytorch.base"]
    init = sys.modules["segmentation_models_pytorch.base.initialization"]
    init.initialize_decoder = init.initialize_head = lambda *a, **k: None
    sys.path.insert(0, REF)
    try:
        import train_3d  # noqa
    finally:
        sys.path.remove(REF)
    return sys.modules["train_3d"]


def check_close(name, a, b, tol):
    a, b = a.detach().double(), b.detach().double()
    err = (a - b).abs().max().item() if a.numel() else 0.0
    ref = max(b.abs().max().item() if b.numel() else 0.0, 1e-30)
    status = "ok" if err <= tol * max(ref, 1.0) else "MISMATCH"
    print(f"  {name:58s} max|d|={err:.3e} (ref max {ref:.3e}) {status}")
    assert status == "ok", name
    return err


def run_reference_trainer(train_3d, sd0, batches, lr, seed):
    """The unmodified reference ``train_pcrlv2_inner`` on CPU (``.cuda()`` / ``synchronize`` no-ops).
    Returns (state_dict, momentum buffers by name, mg_avg, local_avg)."""
    model = train_3d.PCRLv23d()
    model.load_state_dict(orc.clone_state(sd0))
    args = types.SimpleNamespace(lr=lr, momentum=0.9, weight_decay=1e-4, amp=False, epochs=240)
    opt = torch.optim.SGD(model.parameters(), lr=args.lr, momentum=args.momentum,
                          weight_decay=args.weight_decay)
    orig_cuda, orig_sync = torch.Tensor.cuda, torch.cuda.synchronize
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.cuda.synchronize = lambda *a, **k: None
    random.seed(seed)
    try:
        loader = [(b[0], b[1], b[2], b[2], b[3]) for b in batches]
        mg_avg, local_avg = train_3d.train_pcrlv2_inner(args, 0, loader, model, opt,
                                                        torch.nn.MSELoss(), torch.nn.CosineSimilarity())
    finally:
        torch.Tensor.cuda, torch.cuda.synchronize = orig_cuda, orig_sync
    name_of = {id(p): n for n, p in model.named_parameters()}
    mom = {name_of[id(p)]: st["momentum_buffer"] for p, st in opt.state.items()
           if st.get("momentum_buffer") is not None}
    return model.state_dict(), mom, mg_avg, float(local_avg)


B16 = dict(bsz=16, vol=(32, 32, 16), seeds=(42, 43), lr=1e-2, rng_seed=1234, n_samples=1024)


def golden_b16(train_3d):
    """3b: reference trainer vs oracle at b=16, 32x32x16 (BatchNorm1d over 16 / 96 rows).

    At this size the reference's own fp32 arithmetic is not reproducible to 2e-5 any more: the
    BatchNorm backward cancellation amplifies fp32 rounding towards the stem, and two fp32 evaluation
    orders of the SAME math (the reference's module graph vs the oracle's functional graph) differ by
    ~2e-3 in the stem's momentum buffer.  The oracle is therefore also run in fp64 ("truth"): the
    fixture carries the truth, and per tensor the FLOOR = distance of the reference's fp32 result from
    the truth; oracle-fp32 must sit within 4 floors of the reference trainer."""
    c = B16
    sd0 = orc.init_state(0)
    batches = [orc.synthetic_batch(c["bsz"], seed=s, vol=c["vol"]) for s in c["seeds"]]
    ref_sd, ref_mom, mg_avg, local_avg = run_reference_trainer(train_3d, sd0, batches, c["lr"], c["rng_seed"])

    def run_oracle(dtype):
        sd = orc.clone_state(sd0, dtype)
        bufs, rng, scal, draws_all, g1 = {}, random.Random(c["rng_seed"]), [], [], None
        for b in batches:
            bb = [b[0].to(dtype), b[1].to(dtype), b[2].to(dtype), [v.to(dtype) for v in b[3]]]
            s_, draws, grads = orc.train_step(sd, bufs, bb[0], bb[1], bb[2], bb[3], 0, c["lr"], rng)
            g1 = grads if g1 is None else g1
            scal.append(s_)
            draws_all.append(draws)
        return sd, bufs, scal, draws_all, g1

    sd, bufs, scal, draws_all, g1_32 = run_oracle(torch.float32)
    sd64, bufs64, scal64, draws64, g1_64 = run_oracle(torch.float64)
    assert draws64 == draws_all
    print("[3b] trainer parity at b=16, 32x32x16 (reference train_pcrlv2_inner, 2 iterations) draws:", draws_all)

    def upd_err(a, b, k):      # rel-L2 of the 2-step update of parameter k
        i0 = sd0[k].double()
        da, db = a.double() - i0, b.double() - i0
        return ((da - db).norm() / db.norm().clamp_min(1e-30)).item()

    floors = {}
    for k, v in ref_sd.items():
        if orc.is_param(k) and k in ref_mom and orc.is_cancelling(k):
            floors[k] = float("inf")       # exact gradient is zero: the update is weight decay + noise
        elif orc.is_param(k) and k in ref_mom:
            floors[k] = upd_err(v, sd64[k], k)
            e = upd_err(sd[k], v, k)
            status = "ok" if e <= max(2e-5, 4 * floors[k]) else "MISMATCH"
            print(f"  update {k:52s} oracle-fp32 vs reference {e:.3e}; reference vs fp64 truth {floors[k]:.3e} {status}")
            assert status == "ok", k
        else:
            check_close(f"state {k}", sd[k].double(), v.double(), 2e-4 if not orc.is_param(k) else 0.0)
    assert set(ref_mom) == set(bufs) == set(bufs64), sorted(set(ref_mom) ^ set(bufs))
    for k, v in ref_mom.items():
        if orc.is_cancelling(k):
            continue
        fl = ((v.double() - bufs64[k]).norm() / bufs64[k].norm().clamp_min(1e-30)).item()
        e = ((bufs[k].double() - v.double()).norm() / v.double().norm().clamp_min(1e-30)).item()
        assert e <= max(2e-5, 4 * fl), (k, e, fl)
    mg = sum(s_["loss1"] for s_ in scal) / 2
    lc = sum(s_["local_loss"] for s_ in scal) / 2
    print(f"  mg_loss avg oracle {mg:.8f} ref {mg_avg:.8f}; local avg oracle {lc:.8f} ref {local_avg:.8f}")
    assert abs(mg - mg_avg) < 2e-6 and abs(lc - local_avg) < 2e-5
    tr = {"draws": np.array(draws_all), "mg_avg": np.float64(mg_avg), "local_avg": np.float64(local_avg),
          "lr": np.float64(c["lr"])}
    for i, s_ in enumerate(scal64):
        for k, v in s_.items():
            tr[f"step{i}.{k}"] = np.float64(v)                       # fp64 truth of every loss term
            tr[f"step{i}.f32.{k}"] = np.float64(scal[i][k])
    for k, v in ref_sd.items():
        tr[f"state.{k}"] = digest(v, c["n_samples"]) if v.numel() > 1 else v.numpy()
        if orc.is_param(k) and k in ref_mom:
            tr[f"truth.{k}"] = digest(sd64[k], c["n_samples"])
            tr[f"floor.{k}"] = np.float64(floors[k])
            # step-1 gradient: fp64 truth samples and the distance of the fp32 evaluation from it
            if g1_64[k] is not None:      # None: reached by no loss term of step 1 (note N3)
                tr[f"grad1.{k}"] = digest(g1_64[k], c["n_samples"])
                tr[f"floor1.{k}"] = np.float64(float("inf") if orc.is_cancelling(k) else
                                               ((g1_32[k].double() - g1_64[k]).norm() / g1_64[k].norm().clamp_min(1e-30)).item())
    for k, v in ref_mom.items():
        tr[f"mom.{k}"] = digest(v, c["n_samples"])
    np.savez_compressed(os.path.join(GOLD, "train_2steps_b16.npz"), **tr)
    print("wrote train_2steps_b16.npz")


def main():
    torch.set_num_threads(os.cpu_count())
    os.makedirs(GOLD, exist_ok=True)
    if sys.argv[1:] == ["b16"]:
        golden_b16(load_ref_train_module())
        return
    refmod = load_ref_model_module()
    report = []

    # ---- 1. state layout
    for norm in ("bn", "in"):
        ref = refmod.PCRLv23d(norm=norm)
        rsd = ref.state_dict()
        spec = orc.state_spec(norm=norm)
        assert [k for k, _, _ in spec] == list(rsd.keys()), f"key order mismatch ({norm})"
        for k, shape, _ in spec:
            assert tuple(rsd[k].shape) == tuple(shape), (k, rsd[k].shape, shape)
        print(f"[1] state layout norm={norm}: {len(spec)} entries match the reference")
    n_par = sum(v.numel() for k, v in refmod.PCRLv23d().state_dict().items() if orc.is_param(k))
    assert n_par == 17111434, n_par

    # ---- 2. forward parity + fixtures
    out = {}
    for norm in ("bn", "in"):
        sd0 = orc.init_state(0, norm=norm)
        x1, x2, gt, lv = orc.synthetic_batch(2, seed=42)
        ref = refmod.PCRLv23d(norm=norm)
        ref.load_state_dict(orc.clone_state(sd0))
        ref.train()
        sd = orc.clone_state(sd0)
        with torch.no_grad():
            r_out, r_feats, r_masks = ref(x1)
            o_out, o_feats, o_masks = orc.forward(sd, x1, False, True, "relu", norm)
            r_lout, r_lfeats, r_lmasks = ref(torch.cat(lv, 0), local=True)
            o_lout, o_lfeats, o_lmasks = orc.forward(sd, torch.cat(lv, 0), True, True, "relu", norm)
        print(f"[2] forward parity norm={norm}")
        check_close("out", o_out, r_out, 1e-6)
        for s in range(3):
            check_close(f"pro[{s}]", o_feats[s][0], r_feats[s][0], 1e-5)
            check_close(f"pre[{s}]", o_feats[s][1], r_feats[s][1], 1e-5)
            check_close(f"mask[{s}]", o_masks[s], r_masks[s], 1e-6)
            check_close(f"local pro[{s}]", o_lfeats[s][0], r_lfeats[s][0], 1e-5)
            check_close(f"local pre[{s}]", o_lfeats[s][1], r_lfeats[s][1], 1e-5)
        assert r_lmasks == [] and o_lmasks == []
        for k, v in ref.state_dict().items():
            if not orc.is_param(k):
                check_close(f"buffer {k}", sd[k].double(), v.double(), 1e-6)
        out[f"{norm}.out"] = digest(r_out)
        out[f"{norm}.local_out"] = digest(r_lout)
        for s in range(3):
            out[f"{norm}.pro{s}"] = r_feats[s][0].numpy()
            out[f"{norm}.pre{s}"] = r_feats[s][1].numpy()
            out[f"{norm}.mask{s}"] = digest(r_masks[s])
            out[f"{norm}.local_pro{s}"] = r_lfeats[s][0].numpy()
            out[f"{norm}.local_pre{s}"] = r_lfeats[s][1].numpy()
        for k, v in ref.state_dict().items():
            if k.endswith("running_mean") or k.endswith("running_var"):
                out[f"{norm}.buf.{k}"] = digest(v)
    np.savez_compressed(os.path.join(GOLD, "forward_b2.npz"), **out)
    print("wrote forward_b2.npz")

    # ---- 2b. activation variants (reference models/pcrlv2_model_3d.py:20-27): forward + gradients of
    # the restoration terms at b=2, 32x32x16
    out = {}
    for act in ("prelu", "elu"):
        sd0 = orc.init_state(0, act=act)
        x1, _, gt, _ = orc.synthetic_batch(2, seed=7, vol=(32, 32, 16))
        ref = refmod.PCRLv23d(act=act)
        ref.load_state_dict(orc.clone_state(sd0))
        ref.train()
        r_out, _, r_masks = ref(x1)
        r_loss = torch.nn.functional.mse_loss(r_out, gt) + torch.nn.functional.mse_loss(r_masks[1], gt)
        r_loss.backward()
        r_grads = {n: p.grad for n, p in ref.named_parameters()}
        sd = orc.clone_state(sd0)
        keys = [k for k in sd if orc.is_param(k)]
        for k in keys:
            sd[k].requires_grad_(True)
        o_out, _, o_masks = orc.forward(sd, x1, False, True, act, "bn")
        o_loss = torch.nn.functional.mse_loss(o_out, gt) + torch.nn.functional.mse_loss(o_masks[1], gt)
        o_grads = dict(zip(keys, torch.autograd.grad(o_loss, [sd[k] for k in keys], allow_unused=True)))
        print(f"[2b] act={act}: loss oracle {o_loss.item():.8f} ref {r_loss.item():.8f}")
        check_close("out", o_out, r_out, 1e-6)
        for s_ in range(3):
            check_close(f"mask[{s_}]", o_masks[s_], r_masks[s_], 1e-6)
        for k in keys:
            if r_grads[k] is None:
                assert o_grads[k] is None, k
            else:
                check_close(f"grad {k}", o_grads[k], r_grads[k], 2e-5)
        out[f"{act}.loss"] = np.float64(r_loss.item())
        out[f"{act}.out"] = digest(r_out)
        for s_ in range(3):
            out[f"{act}.mask{s_}"] = digest(r_masks[s_])
        for k in keys:
            if r_grads[k] is not None and (k.endswith("activation.weight") or k.endswith("conv1.weight")):
                out[f"{act}.grad.{k}"] = digest(r_grads[k])
    np.savez_compressed(os.path.join(GOLD, "acts_b2.npz"), **out)
    print("wrote acts_b2.npz")

    # ---- 3. the real reference trainer, two iterations on CPU
    train_3d = load_ref_train_module()
    sd0 = orc.init_state(0)
    batches = [orc.synthetic_batch(2, seed=42), orc.synthetic_batch(2, seed=43)]
    model = train_3d.PCRLv23d()
    model.load_state_dict(orc.clone_state(sd0))
    args = types.SimpleNamespace(lr=1e-3, momentum=0.9, weight_decay=1e-4, amp=False, epochs=240)
    opt = torch.optim.SGD(model.parameters(), lr=args.lr, momentum=args.momentum,
                          weight_decay=args.weight_decay)
    orig_cuda, orig_sync = torch.Tensor.cuda, torch.cuda.synchronize
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.cuda.synchronize = lambda *a, **k: None
    random.seed(1234)
    try:
        loader = [(b[0], b[1], b[2], b[2], b[3]) for b in batches]
        mg_avg, local_avg = train_3d.train_pcrlv2_inner(args, 0, loader, model, opt,
                                                        torch.nn.MSELoss(), torch.nn.CosineSimilarity())
    finally:
        torch.Tensor.cuda, torch.cuda.synchronize = orig_cuda, orig_sync
    ref_sd = model.state_dict()
    name_of = {id(p): n for n, p in model.named_parameters()}
    ref_mom = {name_of[id(p)]: st["momentum_buffer"] for p, st in opt.state.items()
               if st.get("momentum_buffer") is not None}

    sd = orc.clone_state(sd0)
    bufs = {}
    rng = random.Random(1234)
    scal, draws_all, l1s, locs = [], [], [], []
    for b in batches:
        s, draws, _ = orc.train_step(sd, bufs, b[0], b[1], b[2], b[3], 0, 1e-3, rng)
        scal.append(s)
        draws_all.append(draws)
    print("[3] trainer parity (reference train_pcrlv2_inner, 2 iterations) draws:", draws_all)
    worst = 0.0
    for k, v in ref_sd.items():
        worst = max(worst, check_close(f"state {k}", sd[k].double(), v.double(), 2e-5))
    assert set(ref_mom) == set(bufs), (sorted(set(ref_mom) ^ set(bufs)))
    for k, v in ref_mom.items():
        check_close(f"momentum {k}", bufs[k], v, 2e-5)
    mg = sum(s["loss1"] for s in scal) / 2
    lc = sum(s["local_loss"] for s in scal) / 2
    print(f"  mg_loss avg oracle {mg:.8f} ref {mg_avg:.8f}; local avg oracle {lc:.8f} ref {float(local_avg):.8f}")
    assert abs(mg - mg_avg) < 1e-6 and abs(lc - float(local_avg)) < 1e-6

    tr = {"draws": np.array(draws_all), "mg_avg": np.float64(mg_avg),
          "local_avg": np.float64(float(local_avg))}
    for i, s in enumerate(scal):
        for k, v in s.items():
            tr[f"step{i}.{k}"] = np.float64(v)
    for k, v in ref_sd.items():
        tr[f"state.{k}"] = digest(v) if v.numel() > 1 else v.numpy()
    for k, v in ref_mom.items():
        tr[f"mom.{k}"] = digest(v)
    np.savez_compressed(os.path.join(GOLD, "train_2steps_b2.npz"), **tr)
    print("wrote train_2steps_b2.npz")

    # ---- 3b. the same at a well-conditioned batch
    golden_b16(train_3d)


if __name__ == "__main__":
    main()
