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
        self._flat_p = torch.zeros(total, dtype=torch.float32, device=dev)
        self._flat_g = torch.zeros(total, dtype=torch.float32, device=dev)
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
            scale = allreduce_flat_gradients(self._flat_g, self._pg)
        dev = self._flat_p.device
        active = torch.tensor([1 if t else 0 for t in self._touched], dtype=torch.int32).to(dev, non_blocking=True)
        first = torch.tensor([0 if b else 1 for b in self._has_buf], dtype=torch.int32).to(dev, non_blocking=True)
        K.sgd_flat(self._flat_p, self._flat_g, self._flat_m, self._seg_off, active, first,
                   group["lr"], group["momentum"], group["weight_decay"], scale)
        for i, p in enumerate(self._ps):
            if self._touched[i] and not self._has_buf[i]:
                self._has_buf[i] = True
                o = self._offs[i]
                self.state[p]["momentum_buffer"] = self._flat_m[o:o + p.numel()].view_as(p)
        bump_param_epoch()

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
    return type(cosine) is nn.CosineSimilarity and cosine.dim == 1


def _mse(criterion, pred, target):
    """criterion(pred, target); nn.MSELoss() (the reference's criterion, train_3d.py:56) runs on the
    fused squared-error kernels."""
    if type(criterion) is nn.MSELoss and criterion.reduction == "mean" and pred.is_cuda \
            and pred.dtype == torch.float32 and pred.shape == target.shape:
        return Fn.mse_loss(pred, target)
    return criterion(pred, target)


def pcrlv2_step_loss(model, x1, x2, gt, local_views, epoch, criterion, cosine):
    """Forward part of one iteration, reference train_3d.py:116-138.
    Returns (loss, loss1, loss2, local_loss)."""
    bsz = x1.size(0)
    mask1, decoder_outputs1, middle_masks1 = model(x1)
    mask2, decoder_outputs2, _ = model(x2)
    loss2, index2 = cos_loss(cosine, decoder_outputs1, decoder_outputs2)
    local_loss = 0.0
    local_input = torch.cat(local_views, dim=0)
    _, local_views_outputs, _ = model(local_input, local=True)
    local_views_outputs = [torch.stack(t) for t in local_views_outputs]
    for i in range(len(local_views)):
        local_views_outputs_tmp = [t[:, bsz * i: bsz * (i + 1)] for t in local_views_outputs]
        loss_local_1, _ = cos_loss(cosine, decoder_outputs1, local_views_outputs_tmp)
        loss_local_2, _ = cos_loss(cosine, decoder_outputs2, local_views_outputs_tmp)
        local_loss += loss_local_1
        local_loss += loss_local_2
    local_loss = local_loss / (2 * len(local_views))
    loss1 = _mse(criterion, mask1, gt)
    beta = 0.5 * (1. + math.cos(math.pi * epoch / 240))
    loss4 = beta * _mse(criterion, middle_masks1[index2], gt)
    loss = loss1 + loss2 + loss4 + local_loss
    return loss, loss1, loss2, local_loss


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
        # ===================meters=====================
        mg_loss_meter.update(loss1.item(), bsz)
        loss_meter.update(loss2.item(), bsz)
        prob_meter.update(local_loss.item() if torch.is_tensor(local_loss) else float(local_loss), bsz)
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
