"""Two-GPU data-parallel equivalence (one process per GPU over NCCL; skipped on a single-GPU box).

The reference shards a batch over GPUs with nn.DataParallel (train_3d.py:54): every replica
normalises with the BatchNorm statistics of ITS shard and the replicas' gradients are summed into one
parameter set.  Here every rank runs the step on its shard and the flat gradient buffer is
all-reduced.  Checked on rank 0 against a single process that evaluates the two shards one after the
other (each with its own BatchNorm statistics) and averages the two gradients:
  * the all-reduced, 1/world-scaled gradient equals that average (up to atomics noise);
  * after the update every rank holds bit-identical parameters, in the eager step and in the
    captured-graph step; both steps agree with each other.
Evidence of a run on 2 x B200: profiles/r02_ddp2_pytest.log."""
import os
import random
import types

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import pcrlv2_oracle as orc  # noqa: E402


def _worker(rank, world, port, out_path):
    import faulthandler
    import sys
    import torch.distributed as dist
    faulthandler.dump_traceback_later(150, exit=True, file=sys.stderr)      # a hang prints where, then exits
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(world), RANK=str(rank),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from pcrlv2_b200.models import PCRLv23d
    from pcrlv2_b200 import train_3d as T
    dev = torch.device("cuda", rank)
    sd0 = orc.init_state(0)
    PB = 8      # per-rank batch: BatchNorm1d over 8 / 48 rows (over 2 rows the contrastive gradient is chaotic)
    full = orc.synthetic_batch(PB * world, seed=9, vol=(32, 32, 16))

    def shard(r):
        sl = slice(PB * r, PB * r + PB)
        return full[0][sl], full[1][sl], full[2][sl], [v[sl] for v in full[3]]

    def fresh():
        m = PCRLv23d(precision="fp32")
        m.load_state_dict(orc.clone_state(sd0))
        return m.to(dev).train()

    crit, cos = torch.nn.MSELoss(), torch.nn.CosineSimilarity()
    log = []

    def rl2(a, b):
        return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()

    # ---- eager data-parallel step
    m = fresh()
    opt = T.FlatSGD(m.parameters(), lr=1e-2, momentum=0.9, weight_decay=1e-4)
    assert opt._distributed
    x1, x2, gt, lv = [t.to(dev) if torch.is_tensor(t) else [v.to(dev) for v in t] for t in shard(rank)]
    random.seed(3)
    loss, *_ = T.pcrlv2_step_loss(m, x1, x2, gt, lv, 0, crit, cos)
    opt.zero_grad()
    loss.backward()
    opt.step()
    del loss
    g_dp = opt._flat_g[:opt._total].clone() / world
    p_dp = opt._flat_p.clone()
    gathered = [torch.zeros_like(p_dp) for _ in range(world)]
    dist.all_gather(gathered, p_dp)
    assert all(torch.equal(gathered[0], t) for t in gathered), "replicas diverged after the eager step"

    # ---- single-process restatement on rank 0: shard by shard, own BatchNorm statistics each
    if rank == 0:
        def single_process_average():
            gs_sum = None
            for r in range(world):
                ms = fresh()
                os_ = T.FlatSGD(ms.parameters(), lr=1e-2, momentum=0.9, weight_decay=1e-4, distributed=False)
                a1, a2, agt, alv = [t.to(dev) if torch.is_tensor(t) else [v.to(dev) for v in t] for t in shard(r)]
                random.seed(3)
                l_, *_ = T.pcrlv2_step_loss(ms, a1, a2, agt, alv, 0, crit, cos)
                os_.zero_grad()
                l_.backward()
                T.join_side_streams()
                torch.cuda.synchronize()
                g = os_._flat_g[:os_._total].clone()
                gs_sum = g if gs_sum is None else gs_sum + g
                del l_
            return gs_sum / world
        ref_a, ref_b = single_process_average(), single_process_average()
        noise = rl2(ref_b, ref_a)         # run-to-run noise of the SAME computation (floating-point atomics)
        e = rl2(g_dp, ref_a)
        log.append(f"all-reduced gradient vs average of per-shard gradients: rel-L2 {e:.3e} "
                   f"(run-to-run noise of the single-process evaluation {noise:.3e})")
        assert e < max(5 * noise, 1e-3), (e, noise)

    # ---- captured-graph data-parallel steps: replicas stay in sync, result agrees with the eager step
    m2 = fresh()
    opt2 = T.FlatSGD(m2.parameters(), lr=1e-2, momentum=0.9, weight_decay=1e-4)
    s = shard(rank)
    args = types.SimpleNamespace(lr=1e-2, momentum=0.9, weight_decay=1e-4, amp=False, epochs=240)
    random.seed(3)
    T.train_pcrlv2_inner(args, 0, [(s[0], s[1], s[2], s[2], s[3])], m2, opt2, crit, cos)
    assert opt2.__dict__.get("_graphed"), "the trainer did not take the captured-graph step"
    p_g = opt2._flat_p.clone()
    dist.all_gather(gathered, p_g)
    assert all(torch.equal(gathered[0], t) for t in gathered), "replicas diverged after the graph step"
    i0 = torch.cat([sd0[n].flatten() for n, _ in m.named_parameters()]).to(dev)
    # flat buffers pad segments to 16 bytes: compare parameter by parameter
    du_e = torch.cat([p.detach().flatten() for p in opt._ps]) - i0
    du_g = torch.cat([p.detach().flatten() for p in opt2._ps]) - i0
    e = rl2(du_g, du_e)
    log.append(f"rank {rank}: update of the graph step vs the eager step rel-L2 {e:.3e}")
    assert e < 0.2, e          # same noise source; the tight comparison is the one above
    # a second replay with another batch keeps them in sync as well
    s2 = orc.synthetic_batch(PB * world, seed=10, vol=(32, 32, 16))
    sl = slice(PB * rank, PB * rank + PB)
    T.train_pcrlv2_inner(args, 0, [(s2[0][sl], s2[1][sl], s2[2][sl], s2[2][sl], [v[sl] for v in s2[3]])], m2, opt2, crit, cos)
    dist.all_gather(gathered, opt2._flat_p.clone())
    assert all(torch.equal(gathered[0], t) for t in gathered)
    # ---- the 2-D path: two captured data-parallel steps, replicas bit-identical, gradient all-reduce active
    from oracle import pcrlv2_oracle_2d as orc2
    from pcrlv2_b200.models import PCRLv2
    from pcrlv2_b200 import train_2d as T2
    m3 = PCRLv2(precision="fp32")
    m3.load_state_dict(orc2.clone_state(orc2.init_state(0)))
    m3 = m3.to(dev).train()
    opt3 = T.FlatSGD(m3.parameters(), lr=1e-2, momentum=0.9, weight_decay=1e-4)
    assert opt3._distributed
    p0 = opt3._flat_p.clone()
    random.seed(3)
    for seed in (21, 22):
        f2 = orc2.synthetic_batch(4 * world, seed=seed, size=(64, 64), local=(32, 32))
        sl = slice(4 * rank, 4 * rank + 4)
        T2.train_pcrlv2_inner(args, 0, [(f2[0][sl], f2[1][sl], f2[2][sl], f2[2][sl], [v[sl] for v in f2[3]])],
                              m3, opt3, crit, cos)
    assert opt3.__dict__.get("_graphed"), "the 2-D trainer did not take the captured-graph step"
    g3 = [torch.zeros_like(opt3._flat_p) for _ in range(world)]
    dist.all_gather(g3, opt3._flat_p.clone())
    assert all(torch.equal(g3[0], t) for t in g3), "2-D replicas diverged"
    assert not torch.equal(opt3._flat_p, p0)
    log.append(f"rank {rank}: 2-D path, two captured data-parallel steps: replicas bit-identical")
    dist.barrier()
    if rank == 0:
        open(out_path, "w").write("ok\n" + "\n".join(log))
    # captured graphs hold NCCL operations: release them before the communicator is torn down
    # (destroy_process_group otherwise waits forever)
    import gc
    opt2._graphed.clear()
    opt3._graphed.clear()
    del m2, opt2, m3, opt3
    gc.collect()
    torch.cuda.synchronize()
    dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_data_parallel_equivalence(tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "ddp.txt")
    port = 29600 + (os.getpid() % 1000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    text = open(out).read()
    print(text)
    assert text.startswith("ok")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.makedirs(os.path.join(root, "gpurun_out"), exist_ok=True)
    open(os.path.join(root, "gpurun_out", "ddp2_equivalence.txt"), "w").write(text)
