"""Input side of the 3-D pre-training path.

The reference's LUNA pipeline (data.py:63-99, datasets/lunaDataset.py) depends on torchio and on
pre-processed ``.npy`` crops; it is CPU-side and outside the hot path ("next").
What the hot path needs from it is the *batch contract*
    (input1, input2, gt1, gt2, [6 local views])
with input*/gt* of shape (B,1,64,64,32) and local views (B,1,16,16,16)
(datasets/lunaDataset.py:79-81).  ``SyntheticLunaPretask`` produces batches of exactly that
contract: inputs ~ N(0,1) (the real pipeline ends in ZNormalization, data.py:87), gt ~ U[0,1)
(HU normalised to [0,1], luna_preprocess.py:135-137).
"""
import torch


class SyntheticLunaPretask(torch.utils.data.Dataset):
    def __init__(self, length=64, vol=(64, 64, 32), local=(16, 16, 16), n_local=6, seed=42):
        self.length, self.vol, self.local, self.n_local, self.seed = length, vol, local, n_local, seed

    def __len__(self):
        return self.length

    def __getitem__(self, index):
        g = torch.Generator().manual_seed(self.seed * 1000003 + index)
        x1 = torch.randn((1,) + tuple(self.vol), generator=g)
        x2 = torch.randn((1,) + tuple(self.vol), generator=g)
        gt1 = torch.rand((1,) + tuple(self.vol), generator=g)
        gt2 = torch.rand((1,) + tuple(self.vol), generator=g)
        local = [torch.randn((1,) + tuple(self.local), generator=g) for _ in range(self.n_local)]
        return x1, x2, gt1, gt2, local


class SyntheticChestPretask(torch.utils.data.Dataset):
    """Batch contract of datasets/chestDataset.py:31-48: two global 3x224x224 crops, their targets, six local
    3x96x96 crops (inputs normalised ~ N(0,1), targets in [0,1))."""

    def __init__(self, length=64, size=(224, 224), local=(96, 96), n_local=6, seed=42):
        self.length, self.size, self.local, self.n_local, self.seed = length, size, local, n_local, seed

    def __len__(self):
        return self.length

    def __getitem__(self, index):
        g = torch.Generator().manual_seed(self.seed * 1000003 + index)
        x1 = torch.randn((3,) + tuple(self.size), generator=g)
        x2 = torch.randn((3,) + tuple(self.size), generator=g)
        gt1 = torch.rand((3,) + tuple(self.size), generator=g)
        gt2 = torch.rand((3,) + tuple(self.size), generator=g)
        local = [torch.randn((3,) + tuple(self.local), generator=g) for _ in range(self.n_local)]
        return x1, x2, gt1, gt2, local


class DataGenerator:
    """Mirror of the reference DataGenerator (data.py:9-12): the LUNA (3-D) and chest (2-D) pretask loaders."""

    def pcrlv2_chest_pretask(self):
        """reference data.py:14-61 (PNG decoding + torchvision transforms on the CPU): synthetic batches of the
        same contract here."""
        args = self.config
        if str(getattr(args, "data", "synthetic")) != "synthetic":
            raise NotImplementedError("only --data synthetic is built for the chest pretask")
        world = int(__import__("os").environ.get("WORLD_SIZE", "1"))
        rank = int(__import__("os").environ.get("RANK", "0"))
        ds = SyntheticChestPretask(length=getattr(args, "synthetic_items", 64 * max(1, args.b)),
                                   seed=getattr(args, "seed", 42) + rank)
        loader = torch.utils.data.DataLoader(ds, batch_size=max(2, args.b // world), shuffle=False,
                                             num_workers=getattr(args, "workers", 0), pin_memory=True,
                                             drop_last=True)
        return {"train": loader, "eval": loader}

    def __init__(self, config):
        self.config = config

    def pcrlv2_luna_pretask(self):
        args = self.config
        if str(getattr(args, "data", "synthetic")) != "synthetic":
            raise NotImplementedError(
                "only --data synthetic is built: the torchio LUNA pipeline of the reference "
                "(data.py:63-99) is CPU-side input staging outside the B200 hot path")
        world = int(__import__("os").environ.get("WORLD_SIZE", "1"))
        rank = int(__import__("os").environ.get("RANK", "0"))
        ds = SyntheticLunaPretask(length=getattr(args, "synthetic_items", 64 * max(1, args.b)),
                                  seed=getattr(args, "seed", 42) + rank)
        per_rank = max(2, args.b // world)
        loader = torch.utils.data.DataLoader(ds, batch_size=per_rank, shuffle=False,
                                             num_workers=getattr(args, "workers", 0),
                                             pin_memory=True, drop_last=True)
        return {"train": loader, "eval": loader}
