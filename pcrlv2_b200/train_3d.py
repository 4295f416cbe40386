"""B200-native mirror of the reference ``train_3d.py`` (3-D pre-training loop).

Same public functions and semantics as the reference:
  train_pcrlv2_3d(args, data_loader, out_channel=3)                       reference :42-83
  train_pcrlv2_inner(args, epoch, train_loader, model, optimizer, criterion, cosine)   :95-173
  cos_loss(cosine, output1, output2)                                      :86-92
Differences in mechanism, not in results:
  * the model is pcrlv2_b200.models.PCRLv23d (sm_100a kernels);
  * nn.DataParallel (single process, :54) is replaced by one process per GPU: every rank runs the
    step on its shard of the batch with per-rank BatchNorm statistics (what DataParallel replicas
    do) and the flat gradient buffer is all-reduced (sum, then divided by the world size) with
    NCCL before the fused SGD kernel.  The 13 scale draws of a step use Python's ``random`` and
    must be seeded identically on every rank;
  * torch.optim.SGD is replaced by FlatSGD: one fused kernel over a flat parameter buffer that
    skips parameters which received no gradient this step, exactly like torch.optim.SGD skips
    ``grad is None`` parameters (SURVEY note N3).
"""
from __future__ import print_function

import math
import os
import random
import sys
import time

import torch
import torch.distributed as dist
import torch.nn as nn

from . import functional as Fn
from . import kernels as K
from .models import PCRLv23d
from .models.pcrlv2_model_3d import bump_param_epoch, join_side_streams
from .utils import AverageMeter, adjust_learning_rate


def allreduce_flat_gradients(flat_g, group=None):
    """The single exchange step of the data-parallel path: sum the flat gradient
    buffer over the ranks; returns the factor (1/world) the fused SGD kernel applies on read."""
    dist.all_reduce(flat_g, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / dist.get_world_size(group)


class FlatSGD(torch.optim.Optimizer):
    """torch.optim.SGD(lr, momentum, weight_decay) semantics (reference train_3d.py:48-51) on flat
    fp32 buffers: parameters, gradients and momentum each live in ONE contiguous allocation, the
    parameters' ``.data`` / ``.grad`` are views into them, and ``step()`` is a single kernel.

    ``zero_grad()`` zero-fills the flat gradient buffer and marks every parameter "untouched"; a
    gradient hook marks the parameters autograd actually reached.  ``step()`` updates only
    those (no weight decay / momentum decay for the others), optionally after an NCCL all-reduce
    of the flat gradient buffer across ``process_group``.
    """

    def __init__(self, params, lr=1e-3, momentum=0.0, weight_decay=0.0, process_group=None,
                 distributed=None):
        params = list(params)
        # the remaining keys are torch.optim.SGD's own defaults: ``state_dict()`` then has exactly the
        # layout the reference's checkpoints carry (train_3d.py:74-76) and loads into torch.optim.SGD,
        # and a reference checkpoint loads into FlatSGD (load_state_dict below)
        defaults = dict(lr=float(lr), momentum=float(momentum), dampening=0, weight_decay=float(weight_decay),
                        nesterov=False, maximize=False, foreach=None, differentiable=False, fused=None)
        super().__init__(params, defaults)
        ps = [p for g in self.param_groups for p in g["params"]]
        if len(self.param_groups) != 1:
            raise ValueError("FlatSGD supports a single parameter group")
        dev = ps[0].device
        if dev.type != "cuda":
            raise RuntimeError("FlatSGD needs CUDA parameters (there is no CPU path)")
        offs = [0]
        for p in ps:
            offs.append(offs[-1] + (p.numel() + 3) // 4 * 4)   # 16-byte aligned segments
        total = offs[-1]
        self._total = total
        self._flat_p = torch.zeros(total, dtype=torch.float32, device=dev)
        # 4 tail floats: slot [total] carries the step's loss through the SAME all-reduce as the
        # gradients, so that every rank takes the reference's "loss > 1000" skip decision
        # (train_3d.py:140) on the global-batch loss without a second collective (captured-graph step)
        self._flat_g = torch.zeros(total + 4, dtype=torch.float32, device=dev)
        self._flat_m = torch.zeros(total, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, o in zip(ps, offs):
                view = self._flat_p[o:o + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
                p.grad = self._flat_g[o:o + p.numel()].view_as(p)
        self._ps = ps
        self._offs = offs
        self._seg_off = torch.tensor(offs, dtype=torch.long, device=dev)
        self._touched = [False] * len(ps)
        self._has_buf = [False] * len(ps)
        # segment flags of the fused kernel: persistent device tensors fed from one pinned host block
        # (row 0 = active, row 1 = first) with a single asynchronous copy per step
        self._flags_host = torch.zeros((2, len(ps)), dtype=torch.int32).pin_memory()
        self._flags_dev = torch.zeros((2, len(ps)), dtype=torch.int32, device=dev)
        for i, p in enumerate(ps):
            p.register_hook(self._make_hook(i))
            # lets the convolution backward accumulate its weight gradient straight into the flat
            # buffer from a side stream (models/pcrlv2_model_3d.py:_wgrad_overlapped) and mark the
            # parameter as reached itself
            p._pcrl_flat = (self, i)
        self._distributed = (dist.is_available() and dist.is_initialized()
                             and dist.get_world_size(process_group) > 1) if distributed is None else distributed
        self._pg = process_group
        bump_param_epoch()

    def _make_hook(self, i):
        def hook(grad):
            # autograd also calls tensor hooks with None when a custom Function returned no
            # gradient for this parameter: only a defined gradient counts
            if grad is not None:
                self._touched[i] = True
            return None
        return hook

    def zero_grad(self, set_to_none=True):
        # the gradient views must stay attached to the flat buffer: "none" is represented by the
        # untouched flag, not by dropping the tensor
        self._flat_g.zero_()
        if torch.cuda.is_current_stream_capturing():
            return      # the views were attached before the capture; host flags come from the draws
        for i, p in enumerate(self._ps):
            self._touched[i] = False
            o = self._offs[i]
            if p.grad is None or p.grad.data_ptr() != self._flat_g.data_ptr() + 4 * o:
                p.grad = self._flat_g[o:o + p.numel()].view_as(p)

    @torch.no_grad()
    def step(self, closure=None):
        group = self.param_groups[0]
        scale = 1.0
        join_side_streams()   # weight gradients still in flight on the side stream (no-op after backward())
        if self._distributed:
            scale = allreduce_flat_gradients(self._flat_g[:self._total], self._pg)
        self.upload_flags(self._touched)
        K.sgd_flat(self._flat_p, self._flat_g, self._flat_m, self._seg_off, self._flags_dev[0], self._flags_dev[1],
                   group["lr"], group["momentum"], group["weight_decay"], scale)
        self.mark_stepped(self._touched)

    def upload_flags(self, touched):
        """(active, first) flags of every segment -> device (pinned staging, one async copy)."""
        # the previous copy out of the pinned block must have been consumed before it is rewritten
        ev = getattr(self, "_flags_event", None)
        if ev is not None:
            ev.synchronize()
        fh = self._flags_host
        fh[0] = torch.tensor([1 if t else 0 for t in touched], dtype=torch.int32)
        fh[1] = torch.tensor([0 if b else 1 for b in self._has_buf], dtype=torch.int32)
        self._flags_dev.copy_(fh, non_blocking=True)
        self._flags_event = torch.cuda.Event()
        self._flags_event.record()

    def mark_stepped(self, touched):
        """Host bookkeeping after an update: reached parameters now own a momentum buffer."""
        for i, p in enumerate(self._ps):
            if touched[i] and not self._has_buf[i]:
                self._has_buf[i] = True
                o = self._offs[i]
                self.state[p]["momentum_buffer"] = self._flat_m[o:o + p.numel()].view_as(p)
        bump_param_epoch()

    @torch.no_grad()
    def step_static(self, hyper, loss):
        """The update as it is captured into a CUDA graph: hyper-parameters ([lr, momentum, weight_decay,
        grad_scale, skip_threshold]) and segment flags are read from device memory at replay time, the
        loss rides in the tail of the flat gradient buffer through the all-reduce and gates the update."""
        join_side_streams()
        self._flat_g[self._total:self._total + 1].copy_(loss.detach().reshape(1))
        if self._distributed:
            dist.all_reduce(self._flat_g, op=dist.ReduceOp.SUM, group=self._pg)
        K.sgd_flat_dev(self._flat_p, self._flat_g, self._flat_m, self._seg_off, self._flags_dev[0],
                       self._flags_dev[1], hyper, self._flat_g[self._total:])

    def reached_from_draws(self, model, draws):
        """Which parameters the reference's autograd graph reaches for a given list of scale draws
        (SURVEY note N3), without running autograd: everything except (a) the deep-supervision heads
        of the scales != draws[0] (only middle_masks1[index2] enters the loss, train_3d.py:137) and (b)
        the projection BatchNorm1d + prediction head of a scale that no cos_loss term drew."""
        names = getattr(self, "_names", None)
        if names is None:
            by_id = {id(p): n for n, p in model.named_parameters()}
            names = self._names = [by_id[id(p)] for p in self._ps]
        stage = {"up_tr256": 0, "up_tr128": 1, "up_tr64": 2}
        drawn = set(int(d) for d in draws)
        out = []
        for n in names:
            top, _, rest = n.partition(".")
            ok = True
            if top in stage:
                if rest.startswith("deep_supervision_head."):
                    ok = stage[top] == int(draws[0])
                elif rest.startswith("bn.") or rest.startswith("predictor_head."):
                    ok = stage[top] in drawn
            out.append(ok)
        return out

    def load_state_dict(self, state_dict):
        """Accepts the ``optimizer`` entry of a reference checkpoint (torch.optim.SGD.state_dict(),
        train_3d.py:76) or FlatSGD's own: momentum buffers are copied into the flat buffer."""
        for g in state_dict["param_groups"]:
            if g.get("nesterov") or g.get("dampening", 0) != 0 or g.get("maximize"):
                raise ValueError("FlatSGD implements SGD with momentum and weight decay only "
                                 "(nesterov / dampening / maximize are not used by the reference)")
        super().load_state_dict(state_dict)
        with torch.no_grad():
            for i, p in enumerate(self._ps):
                o = self._offs[i]
                view = self._flat_m[o:o + p.numel()].view_as(p)
                buf = self.state.get(p, {}).get("momentum_buffer")
                if buf is not None:
                    view.copy_(buf)
                    self.state[p]["momentum_buffer"] = view
                    self._has_buf[i] = True
                else:
                    view.zero_()
                    self._has_buf[i] = False

    def touched_names(self, model):
        names = {id(p): n for n, p in model.named_parameters()}
        return [names[id(p)] for i, p in enumerate(self._ps) if self._touched[i]]


def cos_loss(cosine, output1, output2):
    """reference train_3d.py:86-92 (one scale drawn with random.randint, symmetric negative
    cosine similarity against the detached projection)."""
    index = random.randint(0, len(output1) - 1)
    sample1 = output1[index]
    sample2 = output2[index]
    if _is_plain_cosine(cosine) and sample1[1].is_cuda and sample1[1].dim() == 2:
        # nn.CosineSimilarity(dim=1): fused forward + gradient kernel (csrc/losses.cu), -1/2 folded in
        loss = (Fn.cosine_mean(sample1[1], sample2[0], cosine.eps, -0.5) +
                Fn.cosine_mean(sample2[1], sample1[0], cosine.eps, -0.5))
    else:   # a caller-supplied similarity: evaluate it as the reference does
        loss = -(cosine(sample1[1], sample2[0].detach()).mean() + cosine(sample2[1],
                                                                         sample1[0].detach()).mean()) * 0.5
    return loss, index


def _is_plain_cosine(cosine):
    return type(cosine) is nn.CosineSimilarity and cosine.dim == 1 and cosine.eps == 1e-8


def _is_plain_mse(criterion):
    return type(criterion) is nn.MSELoss and criterion.reduction == "mean"


def _mse(criterion, pred, target):
    """criterion(pred, target); nn.MSELoss() (the reference's criterion, train_3d.py:56) runs on the
    fused squared-error kernels."""
    if _is_plain_mse(criterion) and pred.is_cuda and pred.dtype == torch.float32 and pred.shape == target.shape:
        return Fn.mse_loss(pred, target)
    return criterion(pred, target)


def draw_scales(n_views, n_scales=3):
    """The 1 + 2*n_views ``random.randint`` draws of one iteration in the reference's order
    (train_3d.py:87 called from :119 and, per local view, :129-130): index2, then for every view the
    draw of (decoder 1, view) and of (decoder 2, view).  Consumes Python's global RNG exactly like
    the reference's 13 cos_loss calls."""
    return [random.randint(0, n_scales - 1) for _ in range(1 + 2 * n_views)]


def pcrlv2_step_loss(model, x1, x2, gt, local_views, epoch, criterion, cosine, static=None):
    """Forward part of one iteration, reference train_3d.py:116-138.
    Returns (loss, loss1, loss2, local_loss).

    With the reference's own criterion / similarity (nn.MSELoss, nn.CosineSimilarity) the 13
    cos_loss terms are ONE kernel (csrc/losses.cu:contrastive_kernel) that reads the drawn scales
    from device memory; any other callable is evaluated term by term as the reference does.
    ``static`` (a _StaticCtl, captured-graph mode): the draws and beta come from device buffers that
    the host refreshes before every replay, and the deep-supervision term is evaluated for all
    three scales with the drawn one selected by a device-side weight."""
    bsz = x1.size(0)
    fused = (_is_plain_cosine(cosine) and x1.is_cuda) or static is not None
    mask1, decoder_outputs1, middle_masks1 = model(x1)
    mask2, decoder_outputs2, _ = model(x2)
    if not fused:
        loss2, index2 = cos_loss(cosine, decoder_outputs1, decoder_outputs2)
    local_input = torch.cat(local_views, dim=0)
    _, local_views_outputs, _ = model(local_input, local=True)
    if fused:
        if static is None:
            draws = draw_scales(len(local_views))
            index2 = draws[0]
            draws_dev = torch.tensor(draws, dtype=torch.int32, device=x1.device)
            closs, parts = Fn.contrastive_losses(decoder_outputs1, decoder_outputs2, local_views_outputs,
                                                 draws_dev, tuple(sorted(set(draws))))
        else:
            closs, parts = Fn.contrastive_losses(decoder_outputs1, decoder_outputs2, local_views_outputs,
                                                 static.draws, None)
        loss2, local_loss = parts[0], parts[1]
    else:
        local_loss = 0.0
        local_views_outputs = [torch.stack(t) for t in local_views_outputs]
        for i in range(len(local_views)):
            local_views_outputs_tmp = [t[:, bsz * i: bsz * (i + 1)] for t in local_views_outputs]
            loss_local_1, _ = cos_loss(cosine, decoder_outputs1, local_views_outputs_tmp)
            loss_local_2, _ = cos_loss(cosine, decoder_outputs2, local_views_outputs_tmp)
            local_loss += loss_local_1
            local_loss += loss_local_2
        local_loss = local_loss / (2 * len(local_views))
        closs = loss2 + local_loss
    loss1 = _mse(criterion, mask1, gt)
    if static is not None:
        # beta * MSE(middle_masks1[index2], gt): beta is data (static.w4[index2]); index2 itself selects
        # one of three captured graphs (GraphedStep), so that the backward of the two deep-supervision
        # heads the draw did NOT select is not run at all -- as in the reference (SURVEY note N2/N3)
        k = static.index2
        loss4 = Fn.mse_loss(middle_masks1[k], gt, static.w4[k:k + 1])
    else:
        beta = 0.5 * (1. + math.cos(math.pi * epoch / 240))
        loss4 = beta * _mse(criterion, middle_masks1[index2], gt)
    loss = loss1 + closs + loss4
    return loss, loss1, loss2, local_loss


class _StaticCtl:
    """Device-resident control block of a captured step: draws int32[16], w4 float32[8] (beta * one-hot
    of the drawn deep-supervision scale; 3 scales in the 3-D model, 5 in the 2-D one), hyper float32[8]
    ([lr, momentum, weight_decay, grad_scale, skip_threshold]); one pinned host mirror, one asynchronous
    copy per step."""

    def __init__(self, dev):
        self.host = torch.zeros(32, dtype=torch.int32).pin_memory()
        self.dev = torch.zeros(32, dtype=torch.int32, device=dev)
        self.draws = self.dev[0:16]
        self.w4 = self.dev[16:24].view(torch.float32)
        self.hyper = self.dev[24:32].view(torch.float32)
        self._h_draws = self.host[0:16]
        self._h_w4 = self.host[16:24].view(torch.float32)
        self._h_hyper = self.host[24:32].view(torch.float32)
        self._event = None
        self.index2 = 0          # host constant of the graph being captured (see GraphedStep)

    def upload(self, draws, beta, lr, momentum, weight_decay, grad_scale, skip_threshold):
        if self._event is not None:
            self._event.synchronize()          # the previous copy has left the pinned block
        self._h_draws.zero_()
        self._h_draws[:len(draws)] = torch.tensor(draws, dtype=torch.int32)
        self._h_w4.zero_()
        self._h_w4[draws[0]] = beta
        self._h_hyper[:5] = torch.tensor([lr, momentum, weight_decay, grad_scale, skip_threshold], dtype=torch.float32)
        self.dev.copy_(self.host, non_blocking=True)
        self._event = torch.cuda.Event()
        self._event.record()


class GraphedStep:
    """One training iteration (reference train_3d.py:109-151: three forwards, the four loss terms,
    backward, gradient all-reduce, SGD) captured ONCE into a CUDA graph and replayed per step.

    Why: the eager step issues ~600 kernel launches through Python/ctypes per ~60-100 ms of GPU work;
    with eight ranks on one host that Python time, not the GPU, set the bf16 8-GPU step time (round-1
    SCALE: 0.81 efficiency).  A replay costs the host one cudaGraphLaunch plus two small pinned copies.

    Everything that varies from step to step is DATA of the graph, refreshed by the host before the
    replay: the input batch (static device buffers), the 13 scale draws and beta (``_StaticCtl``), the
    learning rate / skip threshold (device hyper-parameters of ``sgd_flat_dev``) and the per-parameter
    "reached by autograd" flags of SGD (``FlatSGD.reached_from_draws``: unreached parameters are
    skipped entirely, SURVEY note N3).  A graph always evaluates the contrastive terms of all three
    scales (one small kernel; terms the draws did not select contribute exact zeros).  The one draw
    that changes the amount of work -- index2, which deep-supervision mask enters the loss -- selects
    one of THREE graphs, captured lazily on first use into one shared memory pool (they never run
    concurrently and hand nothing to each other but the static output buffer).

    State semantics are the eager step's: parameters, momentum, BatchNorm running statistics and
    ``num_batches_tracked`` are updated in place by the captured kernels.  Results are bit-compatible
    with the eager path up to the order of floating-point atomics (tests/test_graph_gpu.py)."""

    n_scales = 3          # scales a cos_loss draw chooses from (= graphs keyed by index2)
    in_channels = 1

    def __init__(self, model, optimizer, bsz, vol, local, n_local, warmup=1):
        if not isinstance(optimizer, FlatSGD):
            raise TypeError("GraphedStep needs a FlatSGD optimizer")
        dev = optimizer._flat_p.device
        self.model, self.opt, self.key = model, optimizer, (bsz, tuple(vol), tuple(local), n_local)
        self.x1 = torch.zeros((bsz, self.in_channels) + tuple(vol), device=dev)
        self.x2 = torch.zeros_like(self.x1)
        self.gt = torch.zeros_like(self.x1)
        self.local = torch.zeros((n_local * bsz, self.in_channels) + tuple(local), device=dev)
        self.ctl = _StaticCtl(dev)
        self.out = torch.zeros(4, device=dev)          # loss, loss1, loss2, local_loss of the last replay
        self.n_local, self.bsz = n_local, bsz
        self.graphs = {}                 # index2 -> CUDAGraph
        self.pool = None
        self.launches = 0
        self._warmup = warmup
        self._capture(0)

    # -- model-specific pieces (overridden by the 2-D path, train_2d.GraphedStep2d)
    def _step_loss(self, views, crit, cos):
        return pcrlv2_step_loss(self.model, self.x1, self.x2, self.gt, views, 0, crit, cos, static=self.ctl)

    def _reached(self, draws):
        return self.opt.reached_from_draws(self.model, draws)

    # -- the step as it is captured
    def _static_step(self):
        bump_param_epoch()     # the tensor-core operand copies of the weights are re-packed INSIDE the graph
        crit, cos = nn.MSELoss(), nn.CosineSimilarity()
        views = [self.local[i * self.bsz:(i + 1) * self.bsz] for i in range(self.n_local)]
        loss, loss1, loss2, local_loss = self._step_loss(views, crit, cos)
        self.opt.zero_grad()
        loss.backward()
        self.opt.step_static(self.ctl.hyper, loss)
        self.out.copy_(torch.stack([loss.detach(), loss1.detach(), loss2.detach(), local_loss.detach()]))

    def _capture(self, index2):
        from . import _lib
        warmup = self._warmup if not self.graphs else 1
        self.ctl.index2 = index2
        from .models import pcrlv2_model_3d as M
        model, opt = self.model, self.opt
        # the warm-up iterations run for real (they initialise lazily created CUDA state: kernel
        # attributes, the TMA driver entry point, NCCL communicators, the allocator's pools); the
        # training state they touch is put back afterwards
        saved = {"p": opt._flat_p.clone(), "m": opt._flat_m.clone(),
                 "buf": [b.detach().clone() for b in model.buffers()], "rng": random.getstate()}
        group = opt.param_groups[0]
        self.ctl.upload([index2] * (1 + 2 * self.n_local), 1.0, 0.0, group["momentum"], 0.0, 1.0, float("inf"))
        opt.upload_flags([True] * len(opt._ps))
        M._EXPLICIT_JOIN[0] = True       # the side stream is joined by step_static, not by an engine callback
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(max(1, warmup)):
                    self._static_step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            n0 = _lib.launch_count[0]
            try:
                with torch.cuda.graph(graph, pool=self.pool, capture_error_mode="thread_local"):
                    self._static_step()
            except RuntimeError as e:
                raise RuntimeError(
                    "capturing the training step into a CUDA graph failed (%s).  The usual cause: an autograd "
                    "graph of an EARLIER eager iteration is still alive (a loss tensor kept in a variable), "
                    "which pins the parameters' AccumulateGrad nodes to the default stream; drop those "
                    "references before the first captured step, or set PCRL_GRAPH=0 to keep the eager step."
                    % str(e).splitlines()[0]) from e
            self.launches = _lib.launch_count[0] - n0
            self.graphs[index2] = graph
            if self.pool is None:
                self.pool = graph.pool()
        finally:
            M._EXPLICIT_JOIN[0] = False
        with torch.no_grad():
            opt._flat_p.copy_(saved["p"])
            opt._flat_m.copy_(saved["m"])
            for b, v in zip(model.buffers(), saved["buf"]):
                b.copy_(v)
        random.setstate(saved["rng"])
        bump_param_epoch()
        torch.cuda.synchronize()

    def capture_all(self):
        """Capture the graphs of all index2 values now (otherwise: lazily on first use)."""
        for k in range(self.n_scales):
            if k not in self.graphs:
                self._capture(k)
        return self

    def load(self, x1, x2, gt, local_views):
        """Stage one batch into the static input buffers (host or device tensors, asynchronous)."""
        self.x1.copy_(x1, non_blocking=True)
        self.x2.copy_(x2, non_blocking=True)
        self.gt.copy_(gt, non_blocking=True)
        b = self.bsz
        for i, v in enumerate(local_views):
            self.local[i * b:(i + 1) * b].copy_(v, non_blocking=True)

    def run(self, epoch, skip_guard=True):
        """Draw the scales (Python's ``random``, the reference's order), refresh the control data and
        replay.  Returns the device tensor [loss, loss1, loss2, local_loss] (no host sync here)."""
        opt = self.opt
        group = opt.param_groups[0]
        draws = draw_scales(self.n_local, self.n_scales)
        if draws[0] not in self.graphs:
            self._capture(draws[0])
        beta = 0.5 * (1. + math.cos(math.pi * epoch / 240))
        world = dist.get_world_size(opt._pg) if opt._distributed else 1
        thr = 1000.0 if (skip_guard and epoch > 10) else float("inf")
        self.ctl.upload(draws, beta, group["lr"], group["momentum"], group["weight_decay"], 1.0 / world, thr)
        reached = self._reached(draws)
        opt.upload_flags(reached)
        self.graphs[draws[0]].replay()
        self.last_reached, self.last_draws = reached, draws
        self.pending = thr != float("inf")     # the skip guard may fire: the caller reports the outcome
        if not self.pending:
            self.finish(False)
        return self.out

    def finish(self, skipped):
        """Host bookkeeping of the replayed update (momentum-buffer ownership, packed-weight cache)."""
        self.pending = False
        if not skipped:
            self.opt._touched = list(self.last_reached)
            self.opt.mark_stepped(self.last_reached)


def graphed_step_for(model, optimizer, criterion, cosine, x1, local_views):
    """The cached GraphedStep of (model, optimizer) for this batch geometry, or None when the step
    cannot be captured (foreign optimizer / criterion / similarity, PCRL_GRAPH=0)."""
    if os.environ.get("PCRL_GRAPH", "1") == "0":
        return None
    if not (isinstance(model, PCRLv23d) and isinstance(optimizer, FlatSGD) and _is_plain_mse(criterion)
            and _is_plain_cosine(cosine) and model.training):
        return None
    key = (x1.shape[0], tuple(x1.shape[2:]), tuple(local_views[0].shape[2:]), len(local_views))
    cache = optimizer.__dict__.setdefault("_graphed", {})
    gs = cache.get(key)
    if gs is None:
        cache.clear()          # one geometry at a time: a captured graph pins its activation memory
        gs = cache[key] = GraphedStep(model, optimizer, key[0], key[1], key[2], key[3])
    return gs


def train_pcrlv2_inner(args, epoch, train_loader, model, optimizer, criterion, cosine):
    """one epoch of pre-training, reference train_3d.py:95-173"""
    model.train()
    batch_time = AverageMeter()
    data_time = AverageMeter()
    loss_meter = AverageMeter()
    mg_loss_meter = AverageMeter()
    prob_meter = AverageMeter()
    dev = next(model.parameters()).device
    end = time.time()
    for idx, (input1, input2, gt, gt2, local_views) in enumerate(train_loader):
        data_time.update(time.time() - end)
        bsz = input1.size(0)
        gs = graphed_step_for(model, optimizer, criterion, cosine, input1, local_views)
        if gs is not None:
            # captured step: inputs go straight into the graph's static buffers, one replay, one
            # read-back of the four loss scalars (the reference reads three .item()s per iteration)
            gs.load(input1.float(), input2.float(), gt.float(), [v.float() for v in local_views])
            vals = gs.run(epoch).tolist()
            world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
            skipped = epoch > 10 and (optimizer._flat_g[optimizer._total].item() / world if world > 1 else vals[0]) > 1000
            if gs.pending:
                gs.finish(skipped)
            if skipped:
                print('skip the step')
                continue
            loss1_v, loss2_v, local_v = vals[1], vals[2], vals[3]
        else:
            x1 = input1.float().to(dev, non_blocking=True)
            x2 = input2.float().to(dev, non_blocking=True)
            gt = gt.float().to(dev, non_blocking=True)
            local_views = [v.float().to(dev, non_blocking=True) for v in local_views]
            loss, loss1, loss2, local_loss = pcrlv2_step_loss(model, x1, x2, gt, local_views, epoch,
                                                              criterion, cosine)
            # ===================backward=====================
            if epoch > 10 and _skip_step(loss):   # reference :140; ordered to avoid a host sync early on
                print('skip the step')
                continue
            optimizer.zero_grad()
            loss.backward()
            optimizer.step()
            loss1_v, loss2_v = loss1.item(), loss2.item()
            local_v = local_loss.item() if torch.is_tensor(local_loss) else float(local_loss)
        # ===================meters=====================
        mg_loss_meter.update(loss1_v, bsz)
        loss_meter.update(loss2_v, bsz)
        prob_meter.update(local_v, bsz)
        torch.cuda.synchronize()
        batch_time.update(time.time() - end)
        end = time.time()
        if (idx + 1) % 10 == 0 and _rank() == 0:
            print('Train: [{0}][{1}/{2}]\t'
                  'BT {batch_time.val:.3f} ({batch_time.avg:.3f})\t'
                  'DT {data_time.val:.3f} ({data_time.avg:.3f})\t'
                  'cos_loss {c2l_loss.val:.3f} ({c2l_loss.avg:.3f})\t'
                  'mg loss {mg_loss.val:.3f} ({mg_loss.avg:.3f})\t'
                  'local loss {prob.val:.3f} ({prob.avg:.3f})'.format(
                epoch, idx + 1, len(train_loader), batch_time=batch_time,
                data_time=data_time, c2l_loss=loss_meter, mg_loss=mg_loss_meter, prob=prob_meter))
            sys.stdout.flush()
    return mg_loss_meter.avg, prob_meter.avg


def _skip_step(loss):
    """``loss > 1000`` of reference train_3d.py:140, decided on the GLOBAL-batch loss: under
    nn.DataParallel the reference has one loss over the whole batch; with one process per GPU every
    rank must take the same branch (a rank that skipped would leave the others' gradient all-reduce
    without a partner), so the shard losses are averaged over the ranks first -- equal shards, every
    term a batch mean, hence exactly the DataParallel loss."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        g = loss.detach().clone()
        dist.all_reduce(g, op=dist.ReduceOp.SUM)
        return bool(g.item() / dist.get_world_size() > 1000)
    return bool(loss > 1000)


def _rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def init_distributed():
    """One process per GPU (torchrun): NCCL over NVLink.  Returns (rank, world, device)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return (dist.get_rank() if world > 1 else 0), world, torch.device("cuda", local)


def train_pcrlv2_3d(args, data_loader, out_channel=3):
    """reference train_3d.py:42-83: model + SGD set-up, cosine LR per epoch, epochs+1 epochs,
    checkpoint at epoch % 100 == 0 or epoch == 240 with the reference's dict / file name."""
    train_loader = data_loader['train']
    rank, world, dev = init_distributed()
    # precision follows the reference's switch (train_3d.py:52-53, main.py:39): fp32 storage / TF32
    # tensor-core operands by default (what the reference's cuDNN convolutions compute under torch's
    # default allow_tf32), bf16 storage + operands under --amp (apex O1 runs the convolutions in half
    # precision; bf16 needs no loss scaling)
    precision = "bf16" if getattr(args, "amp", False) else "fp32"
    model = PCRLv23d(precision=precision).to(dev)
    if rank == 0:
        print("precision: %s (%s)" % (precision, "--amp" if precision == "bf16" else "default; --amp selects bf16"))
    if world > 1:
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, src=0)
    optimizer = FlatSGD(model.parameters(), lr=args.lr, momentum=float(args.momentum),
                        weight_decay=float(args.weight_decay))
    criterion = nn.MSELoss().to(dev)
    cosine = nn.CosineSimilarity().to(dev)
    for epoch in range(0, args.epochs + 1):
        adjust_learning_rate(epoch, args, optimizer)
        if rank == 0:
            print("==> training...")
        time1 = time.time()
        loss, prob = train_pcrlv2_inner(args, epoch, train_loader, model, optimizer, criterion, cosine)
        time2 = time.time()
        if rank == 0:
            print('epoch {}, total time {:.2f}'.format(epoch, time2 - time1))
        if (epoch % 100 == 0 or epoch == 240) and rank == 0:
            print('==> Saving...')
            state = {'opt': args, 'state_dict': {k: v.detach().clone() for k, v in model.state_dict().items()},
                     'optimizer': optimizer.state_dict(), 'epoch': epoch}
            save_file = os.path.join(args.output,
                                     args.model + "_" + args.n + '_' + args.phase + '_' + str(
                                         args.ratio) + '_' + str(epoch) + '.pt')
            torch.save(state, save_file)
            del state
        torch.cuda.empty_cache()
    return model
