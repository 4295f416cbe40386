"""Training-loop helpers on the hot path (reference utils.py:101-137)."""
import math


def adjust_learning_rate(epoch, args, optimizer):
    """Per-epoch cosine schedule, reference utils.py:111-114."""
    lr = args.lr * 0.5 * (1.0 + math.cos(math.pi * epoch / args.epochs))
    for param_group in optimizer.param_groups:
        param_group["lr"] = lr


class AverageMeter(object):
    """Running mean, reference utils.py:117-137."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val = 0
        self.avg = 0
        self.sum = 0
        self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count
