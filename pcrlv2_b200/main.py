"""CLI mirror of the reference main.py:21-50: the 3-D (--n luna --d 3) and 2-D (--n chest --d 2) pre-training paths.

Flags are the reference's (main.py:23-39).  ``--gpus`` keeps its meaning as the visible device
list, but multi-GPU runs are launched one process per GPU:
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 -m pcrlv2_b200.main --n luna --d 3 --b 256 ...
``--data synthetic`` selects the synthetic LUNA-shaped loader (pcrlv2_b200/data.py).
"""
import argparse
import os
import random
import warnings

import torch

from .data import DataGenerator
from .train_2d import train_pcrlv2
from .train_3d import train_pcrlv2_3d

warnings.filterwarnings('ignore')


def get_dataloader(args):
    generator = DataGenerator(args)
    loader_name = args.model + '_' + args.n + '_' + args.phase
    print(loader_name)
    dataloader = getattr(generator, loader_name)()
    return dataloader


# (flag, default, type, help) -- the reference's command line (main.py:23-39), flag for flag, plus --synthetic_items
_FLAGS = (
    ("data", "synthetic", str, "path to dataset ('synthetic' = generated batches of the reference's shapes)"),
    ("model", "pcrlv2", str, "choose the model"),
    ("phase", "pretask", str, "pretask or finetune or train from scratch"),
    ("b", 16, int, "batch size"),
    ("epochs", 100, int, "epochs to train"),
    ("lr", 1e-3, float, "learning rate"),
    ("output", "./model_genesis_pretrain", str, "output path"),
    ("n", "luna", str, "dataset to use"),
    ("d", 3, int, "3d or 2d to run"),
    ("workers", 4, int, "num of workers"),
    ("gpus", "0,1,2,3", str, "gpu indexs"),
    ("ratio", 0.8, float, "ratio of data used for pretraining"),
    ("momentum", 0.9, None, None),
    ("weight_decay", 1e-4, None, None),
    ("seed", 42, int, None),
    ("synthetic_items", 64, int, "items per epoch of the synthetic loader"),
)


def build_parser():
    parser = argparse.ArgumentParser(description="Self Training benchmark")
    for name, default, typ, text in _FLAGS:
        kw = {"default": default}
        if typ is not None:
            kw["type"] = typ
        if text is not None:
            kw["help"] = text
        parser.add_argument("--" + name, **kw)
    parser.add_argument("--amp", action="store_true", default=False)
    return parser


def main(argv=None):
    args = build_parser().parse_args(argv)
    if not os.path.exists(args.output):
        os.makedirs(args.output, exist_ok=True)
    if int(os.environ.get("RANK", "0")) == 0:
        print(args)
    if "LOCAL_RANK" not in os.environ:
        os.environ.setdefault("CUDA_VISIBLE_DEVICES", args.gpus)
    # the reference never uses --seed; the scale draws must agree on all ranks, so it is used here
    random.seed(args.seed)
    torch.manual_seed(args.seed)
    data_loader = get_dataloader(args)
    if args.model == 'pcrlv2' and args.phase == 'pretask' and args.d == 2:
        train_pcrlv2(args, data_loader)
    elif args.model == 'pcrlv2' and args.phase == 'pretask' and args.d == 3:
        train_pcrlv2_3d(args, data_loader)
    else:
        raise NotImplementedError("only --model pcrlv2 --phase pretask with --d 2 or --d 3 is part of this build")


if __name__ == '__main__':
    main()
