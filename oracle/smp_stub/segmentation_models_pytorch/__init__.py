"""Restatement of the parts of ``segmentation_models_pytorch`` the reference's 2-D model touches.
TEST INFRASTRUCTURE ONLY (lives under oracle/): it lets oracle/make_golden_2d.py import the
UNMODIFIED /root/reference/models/pcrlv2_model.py in the build container, where the real package
is absent (not installed, not vendored, version unpinned: /root/reference/README.md:7).

What the reference uses (models/pcrlv2_model.py:1,4,6,51,58,78,85-86,93,200-201,204,208):
  smp.Unet('resnet18', in_channels=3, classes=n)  -> .encoder (.out_channels, forward -> 6 features),
                                                     .decoder (replaced by the reference's own),
                                                     .segmentation_head (Conv2d(16, n, 3, padding=1))
  base.modules.Conv2dReLU, base.modules.Attention, base.initialization.initialize_decoder / _head

The published structure restated here (smp 0.3.x, encoders/resnet.py, decoders/unet/model.py,
base/heads.py): the ResNet encoder IS torchvision's ``ResNet(BasicBlock, [2, 2, 2, 2])`` without
``fc`` / ``avgpool``, returning [x, relu(bn1(conv1 x)), layer1(maxpool .), layer2, layer3, layer4];
the segmentation head is Sequential(Conv2d(k=3, padding=1), Identity, Identity).  Weights are NOT
downloaded (the real default ``encoder_weights='imagenet'`` would): random initialisation.
"""
import torch.nn as nn
from torchvision.models.resnet import ResNet, BasicBlock

from .base import modules, initialization  # noqa: F401


class ResNetEncoder(ResNet):
    def __init__(self, out_channels, depth=5, **kwargs):
        super().__init__(**kwargs)
        self._depth = depth
        self._out_channels = out_channels
        self._in_channels = 3
        del self.fc
        del self.avgpool

    @property
    def out_channels(self):
        return self._out_channels[: self._depth + 1]

    def get_stages(self):
        return [nn.Identity(), nn.Sequential(self.conv1, self.bn1, self.relu),
                nn.Sequential(self.maxpool, self.layer1), self.layer2, self.layer3, self.layer4]

    def forward(self, x):
        features = []
        for stage in self.get_stages()[: self._depth + 1]:
            x = stage(x)
            features.append(x)
        return features


class SegmentationHead(nn.Sequential):
    def __init__(self, in_channels, out_channels, kernel_size=3):
        super().__init__(nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, padding=kernel_size // 2),
                         nn.Identity(), nn.Identity())


class Unet(nn.Module):
    def __init__(self, encoder_name="resnet18", in_channels=3, classes=1, decoder_channels=(256, 128, 64, 32, 16)):
        super().__init__()
        if encoder_name != "resnet18" or in_channels != 3:
            raise NotImplementedError("stub: only Unet('resnet18', in_channels=3)")
        self.encoder = ResNetEncoder(out_channels=(3, 64, 64, 128, 256, 512), block=BasicBlock, layers=[2, 2, 2, 2])
        self.decoder = nn.Identity()     # the reference replaces it (models/pcrlv2_model.py:201)
        self.segmentation_head = SegmentationHead(decoder_channels[-1], classes, kernel_size=3)
        initialization.initialize_head(self.segmentation_head)

    def forward(self, x):
        raise NotImplementedError("stub: the reference never calls Unet.forward")
