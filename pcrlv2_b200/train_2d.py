"""Host side of the 2-D pre-training path: mirror of the reference ``train_2d.py`` (same function names,
argument meaning and printed lines): ``cos_loss`` :111-117, ``train_pcrlv2_inner`` :120-195, ``train_pcrlv2``
:62-108, on ``pcrlv2_b200.models.PCRLv2`` and the flat fused SGD of the 3-D path.

Differences from the reference, all outside the arithmetic:
  * one process per GPU (torchrun / NCCL gradient all-reduce) instead of nn.DataParallel;
  * ``--amp`` selects bf16 storage / operands (no apex, no loss scaling: bf16 has fp32's exponent range);
  * the forwards whose masks the loss never reads (x2 and the local views, :142,147) skip the mask heads'
    output convolutions and upsampling (``need_masks=False``); the BatchNorm buffers of those heads are still
    updated, so ``state_dict()`` stays on the reference's trajectory.
"""
from __future__ import annotations

import math
import os
import random
import sys
import time

import torch
import torch.distributed as dist
import torch.nn as nn

from . import functional as Fn
from .models.pcrlv2_model import PCRLv2
from .train_3d import (FlatSGD, GraphedStep, _mse, _is_plain_cosine, _is_plain_mse, _rank, init_distributed)
from .utils import adjust_learning_rate, AverageMeter


def cos_loss(cosine, output1, output2):
    """reference train_2d.py:111-117."""
    index = random.randint(0, len(output1) - 1)
    sample1 = output1[index]
    sample2 = output2[index]
    if _is_plain_cosine(cosine) and sample1[1].is_cuda and sample1[1].dim() == 2:
        loss = (Fn.cosine_mean(sample1[1], sample2[0], cosine.eps, -0.5) +
                Fn.cosine_mean(sample2[1], sample1[0], cosine.eps, -0.5))
    else:
        loss = -(cosine(sample1[1], sample2[0].detach()).mean() + cosine(sample2[1],
                                                                         sample1[0].detach()).mean()) * 0.5
    return loss, index


def pcrlv2_step_loss(model, x1, x2, gt, local_views, epoch, criterion, cosine, static=None):
    """Forward part of one iteration, reference train_2d.py:141-163.  Returns (loss, loss1, loss2, local_loss).
    ``static`` (a train_3d._StaticCtl, captured-graph mode): the 13 cos_loss terms are ONE kernel that reads the
    drawn scales from device memory, beta comes from the device, and the host constant ``static.index2`` selects
    which deep-supervision mask enters the loss (one captured graph per value)."""
    bsz = x1.size(0)
    decoder_outputs1, mask1, middle_masks1 = model(x1)
    decoder_outputs2, _mask2, _ = model(x2, need_masks=False)
    if static is not None:
        local_views_outputs, _, _ = model(torch.cat(local_views, dim=0), local=True, need_masks=False)
        closs, parts = Fn.contrastive_losses(decoder_outputs1, decoder_outputs2, local_views_outputs, static.draws, None)
        loss1 = Fn.mse_loss(mask1, gt)
        k = static.index2
        loss4 = Fn.mse_loss(middle_masks1[k], gt, static.w4[k:k + 1])
        return loss1 + closs + loss4, loss1, parts[0], parts[1]
    loss2, index2 = cos_loss(cosine, decoder_outputs1, decoder_outputs2)
    local_loss = 0.0
    local_input = torch.cat(local_views, dim=0)
    local_views_outputs, _, _ = model(local_input, local=True, need_masks=False)
    # the reference stacks (pro, pre) of every scale and slices view i out of the batch dimension
    # (:148-153): the same rows, without the stack copy
    for i in range(len(local_views)):
        local_views_outputs_tmp = [(t[0][bsz * i: bsz * (i + 1)], t[1][bsz * i: bsz * (i + 1)])
                                   for t in local_views_outputs]
        loss_local_1, _ = cos_loss(cosine, decoder_outputs1, local_views_outputs_tmp)
        loss_local_2, _ = cos_loss(cosine, decoder_outputs2, local_views_outputs_tmp)
        local_loss += loss_local_1
        local_loss += loss_local_2
    local_loss = local_loss / (2 * len(local_views))
    loss1 = _mse(criterion, mask1, gt)
    beta = 0.5 * (1. + math.cos(math.pi * epoch / 240))
    loss4 = beta * _mse(criterion, middle_masks1[index2], gt)
    loss = loss1 + loss2 + local_loss + loss4
    return loss, loss1, loss2, local_loss


class GraphedStep2d(GraphedStep):
    """The 2-D iteration captured into CUDA graphs (see train_3d.GraphedStep): five scales, hence five graphs keyed by
    index2, captured lazily into one memory pool; 3-channel static input buffers.  Without it the 2-D step is
    launch-bound: ~1240 kernel launches from Python per 47 ms step at b=8 (profiles/r02s_bench_2d_eager.jsonl)."""
    n_scales = 5
    in_channels = 3

    def _step_loss(self, views, crit, cos):
        return pcrlv2_step_loss(self.model, self.x1, self.x2, self.gt, views, 0, crit, cos, static=self.ctl)

    def _reached(self, draws):
        """Parameters the reference's autograd graph reaches for these draws (note N3): everything except the
        deep-supervision heads of the blocks != draws[0] (only middle_masks1[index2] enters the loss,
        train_2d.py:158) and the BatchNorm1d + prediction head of a block no cos_loss term drew."""
        names = getattr(self, "_names", None)
        if names is None:
            by_id = {id(p): n for n, p in self.model.named_parameters()}
            names = self._names = [by_id[id(p)] for p in self.opt._ps]
        drawn = set(int(d) for d in draws)
        out = []
        for n in names:
            ok = True
            if n.startswith("model.decoder.blocks."):
                i, _, rest = n[len("model.decoder.blocks."):].partition(".")
                if rest.startswith("deep_supervision_head."):
                    ok = int(i) == int(draws[0])
                elif rest.startswith("bn.") or rest.startswith("predictor_head."):
                    ok = int(i) in drawn
            out.append(ok)
        return out


def graphed_step_for(model, optimizer, criterion, cosine, x1, local_views):
    """The cached GraphedStep2d of (model, optimizer) for this batch geometry, or None when the step cannot be
    captured (foreign optimizer / criterion / similarity, PCRL_GRAPH=0)."""
    if os.environ.get("PCRL_GRAPH", "1") == "0":
        return None
    if not (isinstance(model, PCRLv2) and isinstance(optimizer, FlatSGD) and _is_plain_mse(criterion)
            and _is_plain_cosine(cosine) and model.training):
        return None
    key = (x1.shape[0], tuple(x1.shape[2:]), tuple(local_views[0].shape[2:]), len(local_views))
    cache = optimizer.__dict__.setdefault("_graphed", {})
    gs = cache.get(key)
    if gs is None:
        cache.clear()
        gs = cache[key] = GraphedStep2d(model, optimizer, key[0], key[1], key[2], key[3])
    return gs


def train_pcrlv2_inner(args, epoch, train_loader, model, optimizer, criterion, cosine):
    """one epoch training for instance discrimination -- reference train_2d.py:120-195"""
    model.train()
    batch_time = AverageMeter()
    data_time = AverageMeter()
    loss_meter = AverageMeter()
    mg_loss_meter = AverageMeter()
    prob_meter = AverageMeter()
    all_loss_meter = AverageMeter()
    dev = next(model.parameters()).device
    end = time.time()
    for idx, (input1, input2, gt, gt2, local_views) in enumerate(train_loader):
        data_time.update(time.time() - end)
        bsz = input1.size(0)
        gs = graphed_step_for(model, optimizer, criterion, cosine, input1, local_views)
        if gs is not None:
            # captured step: inputs go into the graph's static buffers, one replay, one read-back of the scalars
            gs.load(input1.float(), input2.float(), gt.float(), [v.float() for v in local_views])
            loss_v, loss1_v, loss2_v, local_v = gs.run(epoch, skip_guard=False).tolist()
        else:
            x1 = input1.float().to(dev, non_blocking=True)
            x2 = input2.float().to(dev, non_blocking=True)
            gt = gt.float().to(dev, non_blocking=True)
            local_views = [v.float().to(dev, non_blocking=True) for v in local_views]
            loss, loss1, loss2, local_loss = pcrlv2_step_loss(model, x1, x2, gt, local_views, epoch, criterion, cosine)
            # ===================backward=====================
            optimizer.zero_grad()
            loss.backward()
            optimizer.step()
            loss_v, loss1_v, loss2_v = loss.item(), loss1.item(), loss2.item()
            local_v = local_loss.item() if torch.is_tensor(local_loss) else float(local_loss)
        # ===================meters=====================
        mg_loss_meter.update(loss1_v, bsz)
        loss_meter.update(loss2_v, bsz)
        prob_meter.update(local_v, bsz)
        all_loss_meter.update(loss_v, bsz)
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
    return loss_meter.avg, mg_loss_meter.avg, prob_meter.avg


def train_pcrlv2(args, data_loader, out_channel=3):
    """reference train_2d.py:62-108: model + SGD, cosine LR per epoch, epochs+1 epochs; the checkpoint holds
    the ENCODER's state_dict only (:99), under the reference's file name."""
    train_loader = data_loader['train']
    rank, world, dev = init_distributed()
    precision = "bf16" if getattr(args, "amp", False) else "fp32"
    model = PCRLv2(precision=precision).to(dev)
    if rank == 0:
        print("precision: %s (%s)" % (precision, "--amp" if precision == "bf16" else "default; --amp selects bf16"))
    if world > 1:
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, src=0)
    optimizer = FlatSGD(model.parameters(), lr=args.lr, momentum=float(args.momentum),
                        weight_decay=float(args.weight_decay))
    criterion = nn.MSELoss().to(dev)
    cosine = nn.CosineSimilarity().to(dev)
    loss_list, mg_loss_list = [], []
    for epoch in range(0, args.epochs + 1):
        adjust_learning_rate(epoch, args, optimizer)
        if rank == 0:
            print("==> training...")
        time1 = time.time()
        loss, mg_loss, prob = train_pcrlv2_inner(args, epoch, train_loader, model, optimizer, criterion, cosine)
        loss_list.append(loss)
        mg_loss_list.append(mg_loss)
        time2 = time.time()
        if rank == 0:
            print('epoch {}, total time {:.2f}'.format(epoch, time2 - time1))
        if (epoch % 100 == 0 or epoch == 240) and rank == 0:
            print('==> Saving...')
            state = {'opt': args,
                     'state_dict': {k: v.detach().clone() for k, v in model.model.encoder.state_dict().items()},
                     'optimizer': optimizer.state_dict(), 'epoch': epoch}
            save_file = os.path.join(args.output,
                                     args.model + "_" + args.n + '_' + args.phase + '_' + str(
                                         args.ratio) + '_' + str(epoch) + '.pt')
            torch.save(state, save_file)
            del state
        torch.cuda.empty_cache()
    return model
