"""Pins oracle/pcrlv2_oracle_2d.py against the reference's 2-D path and writes tests/golden/train2d_*.npz.

Run in the BUILD container only (reads /root/reference):  python oracle/make_golden_2d.py

segmentation_models_pytorch is absent, so the reference's models/pcrlv2_model.py is imported over
oracle/smp_stub (restatement of the handful of smp classes it touches; the encoder there is
torchvision's own ResNet class).  Everything else is the reference's unmodified code:
  1. state layout: oracle state_spec == PCRLv2().state_dict() (keys, order, shapes);
  2. forward parity (train mode, BN buffers) oracle vs reference model at b=4, 64x64 global and
     24 x 32x32 local views;
  3. the REAL ``train_2d.train_pcrlv2_inner`` for two iterations on CPU (b=8, 64x64 + 6 x 32x32, lr 1e-2)
     vs oracle.train_step: parameters, momentum buffers, BN buffers, loss meters -- with the fp64
     evaluation of the oracle as truth and the reference's own distance from it as per-tensor floor
     (same scheme as train_2steps_b16.npz of the 3-D path);
  4. writes tests/golden/train2d_2steps_b8.npz (digests: 512 samples per tensor) and forward2d_b4.npz.
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
from oracle import pcrlv2_oracle_2d as orc  # noqa: E402
from oracle.make_golden import digest, check_close  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
CFG = dict(bsz=8, size=(64, 64), local=(32, 32), seeds=(42, 43), lr=1e-2, rng_seed=1234, n_samples=512)


def load_ref():
    """(reference models/pcrlv2_model.py module, reference train_2d module), imported unmodified."""
    sys.path.insert(0, os.path.join(HERE, "smp_stub"))
    sys.path.insert(0, REF)
    try:
        import train_2d  # noqa  (pulls models/__init__ -> pcrlv2_model.py -> the stubbed smp)
    finally:
        sys.path.remove(REF)
    return sys.modules["models.pcrlv2_model"], sys.modules["train_2d"]


def run_reference_trainer(train_2d, sd0, batches, lr, seed):
    model = train_2d.PCRLv2()
    model.load_state_dict(orc.clone_state(sd0))
    args = types.SimpleNamespace(lr=lr, momentum=0.9, weight_decay=1e-4, amp=False, epochs=240)
    opt = torch.optim.SGD(model.parameters(), lr=args.lr, momentum=args.momentum, weight_decay=args.weight_decay)
    orig_cuda, orig_sync = torch.Tensor.cuda, torch.cuda.synchronize
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.cuda.synchronize = lambda *a, **k: None
    random.seed(seed)
    try:
        loader = [(b[0], b[1], b[2], b[2], b[3]) for b in batches]
        cos_avg, mg_avg, local_avg = train_2d.train_pcrlv2_inner(args, 0, loader, model, opt, torch.nn.MSELoss(),
                                                                 torch.nn.CosineSimilarity())
    finally:
        torch.Tensor.cuda, torch.cuda.synchronize = orig_cuda, orig_sync
    name_of = {id(p): n for n, p in model.named_parameters()}
    mom = {name_of[id(p)]: st["momentum_buffer"] for p, st in opt.state.items()
           if st.get("momentum_buffer") is not None}
    return model.state_dict(), mom, float(cos_avg), float(mg_avg), float(local_avg)


def main():
    torch.set_num_threads(os.cpu_count())
    os.makedirs(GOLD, exist_ok=True)
    refmod, train_2d = load_ref()

    # ---- 1. state layout
    rsd = refmod.PCRLv2().state_dict()
    spec = orc.state_spec()
    assert [k for k, _, _ in spec] == list(rsd.keys()), "key order mismatch"
    for k, shape, _ in spec:
        assert tuple(rsd[k].shape) == tuple(shape), (k, rsd[k].shape, shape)
    n_par = sum(v.numel() for k, v in rsd.items() if orc.is_param(k))
    print(f"[1] state layout: {len(spec)} entries match the reference, {n_par} parameters")
    assert n_par == 14678226, n_par

    # ---- 2. forward parity
    sd0 = orc.init_state(0)
    x1, x2, gt, lv = orc.synthetic_batch(4, seed=42, size=CFG["size"], local=CFG["local"])
    ref = refmod.PCRLv2()
    ref.load_state_dict(orc.clone_state(sd0))
    ref.train()
    sd = orc.clone_state(sd0)
    with torch.no_grad():
        r_dec, r_mask, r_mm = ref(x1)
        o_dec, o_mask, o_mm = orc.forward(sd, x1)
        r_ldec, r_lmask, r_lmm = ref(torch.cat(lv, 0), local=True)
        o_ldec, o_lmask, o_lmm = orc.forward(sd, torch.cat(lv, 0), local=True)
    print("[2] forward parity")
    out = {}
    check_close("mask", o_mask, r_mask, 1e-5)
    assert r_lmask is None and o_lmask is None
    for s in range(5):
        check_close(f"pro[{s}]", o_dec[s][0], r_dec[s][0], 1e-5)
        check_close(f"pre[{s}]", o_dec[s][1], r_dec[s][1], 1e-5)
        check_close(f"middle mask[{s}]", o_mm[s], r_mm[s], 1e-5)
        check_close(f"local pro[{s}]", o_ldec[s][0], r_ldec[s][0], 1e-5)
        check_close(f"local pre[{s}]", o_ldec[s][1], r_ldec[s][1], 1e-5)
        check_close(f"local middle mask[{s}]", o_lmm[s], r_lmm[s], 1e-5)
        out[f"pro{s}"] = r_dec[s][0].numpy()
        out[f"pre{s}"] = r_dec[s][1].numpy()
        out[f"mm{s}"] = digest(r_mm[s])
        out[f"local_pro{s}"] = r_ldec[s][0].numpy()
        out[f"local_pre{s}"] = r_ldec[s][1].numpy()
    out["mask"] = digest(r_mask)
    for k, v in ref.state_dict().items():
        if not orc.is_param(k):
            check_close(f"buffer {k}", sd[k].double(), v.double(), 1e-5)
            if v.numel() > 1:
                out[f"buf.{k}"] = digest(v)
    np.savez_compressed(os.path.join(GOLD, "forward2d_b4.npz"), **out)
    print("wrote forward2d_b4.npz")

    # ---- 3. the reference trainer, two iterations
    c = CFG
    batches = [orc.synthetic_batch(c["bsz"], seed=s, size=c["size"], local=c["local"]) for s in c["seeds"]]
    ref_sd, ref_mom, cos_avg, mg_avg, local_avg = run_reference_trainer(train_2d, sd0, batches, c["lr"], c["rng_seed"])

    def run_oracle(dtype):
        sdx = orc.clone_state(sd0, dtype)
        bufs, rng, scal, draws_all, g1 = {}, random.Random(c["rng_seed"]), [], [], None
        for b in batches:
            bb = [b[0].to(dtype), b[1].to(dtype), b[2].to(dtype), [v.to(dtype) for v in b[3]]]
            s_, draws, grads = orc.train_step(sdx, bufs, bb[0], bb[1], bb[2], bb[3], 0, c["lr"], rng)
            g1 = grads if g1 is None else g1
            scal.append(s_)
            draws_all.append(draws)
        return sdx, bufs, scal, draws_all, g1

    sd32, bufs32, scal, draws_all, g1_32 = run_oracle(torch.float32)
    sd64, bufs64, scal64, draws64, g1_64 = run_oracle(torch.float64)
    assert draws64 == draws_all
    print("[3] trainer parity at b=8, 64x64 (reference train_2d.train_pcrlv2_inner, 2 iterations) draws:", draws_all)

    def upd_err(a, b, k):
        i0 = sd0[k].double()
        da, db = a.double() - i0, b.double() - i0
        return ((da - db).norm() / db.norm().clamp_min(1e-30)).item()

    floors, worst = {}, 0.0
    for k, v in ref_sd.items():
        if orc.is_param(k) and k in ref_mom and orc.is_cancelling(k):
            floors[k] = float("inf")
        elif orc.is_param(k) and k in ref_mom:
            floors[k] = upd_err(v, sd64[k], k)
            e = upd_err(sd32[k], v, k)
            worst = max(worst, e)
            status = "ok" if e <= max(2e-5, 4 * floors[k]) else "MISMATCH"
            if status != "ok" or e > 1e-3:
                print(f"  update {k:60s} oracle-fp32 vs reference {e:.3e}; reference vs fp64 truth {floors[k]:.3e} {status}")
            assert status == "ok", k
        elif v.numel() > 1:
            check_close(f"state {k}", sd32[k].double(), v.double(), 5e-4 if not orc.is_param(k) else 0.0)
    print(f"  {len(floors)} parameter updates checked, worst oracle-fp32 vs reference rel-L2 {worst:.3e}")
    assert set(ref_mom) == set(bufs32) == set(bufs64), sorted(set(ref_mom) ^ set(bufs32))
    mg = sum(s_["loss1"] for s_ in scal) / 2
    lc = sum(s_["local_loss"] for s_ in scal) / 2
    cs = sum(s_["loss2"] for s_ in scal) / 2
    print(f"  meters: mg {mg:.8f}/{mg_avg:.8f} local {lc:.8f}/{local_avg:.8f} cos {cs:.8f}/{cos_avg:.8f} (oracle/reference)")
    assert abs(mg - mg_avg) < 5e-6 and abs(lc - local_avg) < 5e-5 and abs(cs - cos_avg) < 5e-5
    tr = {"draws": np.array(draws_all), "mg_avg": np.float64(mg_avg), "local_avg": np.float64(local_avg),
          "cos_avg": np.float64(cos_avg), "lr": np.float64(c["lr"])}
    for i, s_ in enumerate(scal64):
        for k, v in s_.items():
            tr[f"step{i}.{k}"] = np.float64(v)
            tr[f"step{i}.f32.{k}"] = np.float64(scal[i][k])
    for k, v in ref_sd.items():
        tr[f"state.{k}"] = digest(v, c["n_samples"]) if v.numel() > 1 else v.numpy()
        if orc.is_param(k) and k in ref_mom:
            tr[f"truth.{k}"] = digest(sd64[k], c["n_samples"])
            tr[f"floor.{k}"] = np.float64(floors[k])
            if g1_64[k] is not None:
                tr[f"grad1.{k}"] = digest(g1_64[k], c["n_samples"])
                tr[f"floor1.{k}"] = np.float64(float("inf") if orc.is_cancelling(k) else
                                               ((g1_32[k].double() - g1_64[k]).norm() / g1_64[k].norm().clamp_min(1e-30)).item())
    for k, v in ref_mom.items():
        tr[f"mom.{k}"] = digest(v, c["n_samples"])
    np.savez_compressed(os.path.join(GOLD, "train2d_2steps_b8.npz"), **tr)
    print("wrote train2d_2steps_b8.npz", os.path.getsize(os.path.join(GOLD, "train2d_2steps_b8.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
