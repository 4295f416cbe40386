"""The captured-graph training step (pcrlv2_b200/train_3d.py:GraphedStep) against the eager step it
replaces, the analytic "reached parameters" rule against autograd's own record, and the data-driven
controls of the graph (draws, beta, learning rate, skip guard)."""
import os
import random
import types

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LOG = os.path.join(ROOT, "gpurun_out", "graph_parity.txt")

from oracle import pcrlv2_oracle as orc  # noqa: E402

if torch.cuda.is_available():
    from pcrlv2_b200.models import PCRLv23d
    from pcrlv2_b200 import train_3d as T


def log(msg):
    os.makedirs(os.path.dirname(LOG), exist_ok=True)
    with open(LOG, "a") as f:
        f.write(msg + "\n")
    print(msg)


def rl2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def build(precision="fp32", lr=1e-2):
    sd = orc.init_state(0)
    m = PCRLv23d(precision=precision)
    m.load_state_dict(orc.clone_state(sd))
    m = m.cuda().train()
    opt = T.FlatSGD(m.parameters(), lr=lr, momentum=0.9, weight_decay=1e-4)
    return m, opt, sd


def loader(n, bsz=4, vol=(32, 32, 16), scale_gt=1.0):
    out = []
    for i in range(n):
        b = orc.synthetic_batch(bsz, seed=100 + i, vol=vol)
        out.append((b[0], b[1], b[2] * scale_gt, b[2], b[3]))
    return out


ARGS = types.SimpleNamespace(lr=1e-2, momentum=0.9, weight_decay=1e-4, amp=False, epochs=240)


def run_trainer(precision, graph, epoch=0, n=3, seed=77, scale_gt=1.0, lr=1e-2):
    os.environ["PCRL_GRAPH"] = "1" if graph else "0"
    try:
        m, opt, sd0 = build(precision, lr)
        random.seed(seed)
        mg, local = T.train_pcrlv2_inner(ARGS, epoch, loader(n, scale_gt=scale_gt), m, opt,
                                         torch.nn.MSELoss(), torch.nn.CosineSimilarity())
        torch.cuda.synchronize()
        used_graph = bool(opt.__dict__.get("_graphed"))
        assert used_graph == graph
        return m, opt, sd0, mg, local
    finally:
        os.environ.pop("PCRL_GRAPH", None)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_graphed_step_matches_eager_step(precision):
    """train_pcrlv2_inner eager vs captured graph.  ONE iteration from identical state: same loss terms,
    same reached-parameter set, same update up to the run-to-run noise of floating-point atomics
    (calibrated by a second eager run).  THREE iterations: same structure of the state (momentum-buffer
    ownership, BatchNorm counters); values are compared loosely -- at batch 4 the contrastive terms are
    chaotic from the second iteration on (BatchNorm1d over 4 rows), in the eager path as well."""
    me, oe, sd0, mg_e, lo_e = run_trainer(precision, False, n=1)
    me2, oe2, _, mg_e2, lo_e2 = run_trainer(precision, False, n=1)
    mg_, og, _, mg_g, lo_g = run_trainer(precision, True, n=1)
    log(f"[graph {precision}] 1 step: mg eager {mg_e:.7f} / {mg_e2:.7f} graph {mg_g:.7f}; "
        f"local eager {lo_e:.7f} / {lo_e2:.7f} graph {lo_g:.7f}")
    tol = 2e-6 if precision == "fp32" else 2e-5
    assert abs(mg_g - mg_e) < max(tol, 10 * abs(mg_e2 - mg_e))
    assert abs(lo_g - lo_e) < max(tol, 10 * abs(lo_e2 - lo_e))
    assert oe._has_buf == og._has_buf
    i0 = torch.cat([sd0[n].flatten() for n, _ in me.named_parameters()]).double()

    def upd(opt):
        return torch.cat([p.detach().flatten().double().cpu() for p in opt._ps]) - i0
    noise = ((upd(oe2) - upd(oe)).norm() / upd(oe).norm()).item()
    diff = ((upd(og) - upd(oe)).norm() / upd(oe).norm()).item()
    log(f"[graph {precision}] 1-step update rel-L2 graph vs eager {diff:.3e}; eager run-to-run {noise:.3e}")
    assert diff < max(5 * noise, 1e-3)
    for (k, a), (_, b) in zip(me.state_dict().items(), mg_.state_dict().items()):
        if k.endswith("num_batches_tracked"):
            assert int(a) == int(b) == 3, k
        elif k.endswith("running_mean") or k.endswith("running_var"):
            assert rl2(b, a) < 1e-2, k      # BatchNorm1d statistics over 4 rows: atomics-order noise (2e-3 seen)
    me3, oe3, _, mg_e3, _ = run_trainer(precision, False, n=3)
    mg3, og3, _, mg_g3, _ = run_trainer(precision, True, n=3)
    log(f"[graph {precision}] 3 steps: mg eager {mg_e3:.7f} graph {mg_g3:.7f}")
    assert abs(mg_g3 - mg_e3) < 1e-2 * abs(mg_e3)
    assert oe3._has_buf == og3._has_buf
    for (k, a), (_, b) in zip(me3.state_dict().items(), mg3.state_dict().items()):
        if k.endswith("num_batches_tracked"):
            assert int(a) == int(b) == 9, k
    gs = next(iter(og3._graphed.values()))
    assert gs.launches > 100 and 1 <= len(gs.graphs) <= 3


def test_reached_parameters_rule_matches_autograd():
    """FlatSGD.reached_from_draws (used by the captured step, where autograd hooks do not run) against
    the flags autograd's own hooks set in the eager step, for ordinary draws and for draws that miss
    a scale entirely (SURVEY note N3)."""
    m, opt, _ = build("fp32")
    b = orc.synthetic_batch(2, seed=5, vol=(32, 32, 16))
    crit, cos = torch.nn.MSELoss(), torch.nn.CosineSimilarity()
    cases = [None, None, [1] * 13, [2] + [0] * 12, [0, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2]]
    for i, forced in enumerate(cases):
        random.seed(10 + i)
        state = random.getstate()
        draws = T.draw_scales(6) if forced is None else forced
        random.setstate(state)
        orig = T.draw_scales
        if forced is not None:
            T.draw_scales = lambda n, k=3, f=forced: list(f)
        try:
            loss, *_ = T.pcrlv2_step_loss(m, b[0].cuda(), b[1].cuda(), b[2].cuda(), [v.cuda() for v in b[3]], 0, crit, cos)
        finally:
            T.draw_scales = orig
        opt.zero_grad()
        loss.backward()
        want = opt.reached_from_draws(m, draws)
        assert list(opt._touched) == want, (draws, [n for n, a, b_ in zip(opt._names, opt._touched, want) if a != b_])
        n_un = want.count(False)
        log(f"[reached] draws {draws}: {n_un} parameters unreached")
        assert n_un in (8, 16, 24)      # 2 ds heads (8) [+ one or two scales' projection/prediction heads (8 each)]


def test_unreached_heads_stay_bit_identical_in_graph_mode():
    orig = T.draw_scales
    T.draw_scales = lambda n, k=3: [1] * (1 + 2 * n)
    try:
        m, opt, sd0, _, _ = run_trainer("fp32", True, n=2)
    finally:
        T.draw_scales = orig
    moved = 0
    for n, p in m.named_parameters():
        head = (("deep_supervision_head" in n or ".bn." in n or "predictor_head" in n)
                and (n.startswith("up_tr256") or n.startswith("up_tr64")))
        if head:
            assert torch.equal(p.detach().cpu(), sd0[n]), f"{n} must not move (no weight decay, no momentum)"
        else:
            moved += int(not torch.equal(p.detach().cpu(), sd0[n]))
    assert moved > 60


def test_graph_reads_learning_rate_beta_and_skip_guard_from_data():
    # lr: same batch, same draws, two learning rates -> the update scales with lr (first step: buf = g + wd p)
    ups = []
    for lr in (1e-2, 1e-3):
        orig = T.draw_scales
        T.draw_scales = lambda n, k=3: [2] * (1 + 2 * n)
        try:
            m, opt, sd0, _, _ = run_trainer("fp32", True, n=1, lr=lr)
        finally:
            T.draw_scales = orig
        ups.append(m.out_tr.final_conv.weight.detach().cpu().double() - sd0["out_tr.final_conv.weight"].double())
    ratio = (ups[0].norm() / ups[1].norm()).item()
    log(f"[graph ctl] update ratio for lr 1e-2 / 1e-3: {ratio:.4f}")
    assert abs(ratio - 10.0) < 0.05
    # skip guard (reference train_3d.py:140): epoch > 10 and loss > 1000 -> no update, meters untouched
    for graph in (False, True):
        m, opt, sd0, mg, local = run_trainer("fp32", graph, epoch=11, n=1, scale_gt=100.0)
        assert mg == 0 and local == 0
        for n, p in m.named_parameters():
            assert torch.equal(p.detach().cpu(), sd0[n]), (graph, n)
        assert not any(opt._has_buf)
        # the same batch at epoch 5 is NOT skipped
        m, opt, sd0, mg, local = run_trainer("fp32", graph, epoch=5, n=1, scale_gt=100.0)
        assert mg > 1000 and any(opt._has_buf)
