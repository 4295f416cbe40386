"""smp.base.modules restated (see ../__init__.py): Conv2dReLU = Conv2d(bias = not use_batchnorm) ->
BatchNorm2d -> ReLU as an nn.Sequential (state_dict keys '0.weight', '1.weight', ...);
Attention(None) = Identity."""
import torch.nn as nn


class Conv2dReLU(nn.Sequential):
    def __init__(self, in_channels, out_channels, kernel_size, padding=0, stride=1, use_batchnorm=True):
        conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                         bias=not use_batchnorm)
        relu = nn.ReLU(inplace=True)
        bn = nn.BatchNorm2d(out_channels) if use_batchnorm else nn.Identity()
        super().__init__(conv, bn, relu)


class Attention(nn.Module):
    def __init__(self, name, **params):
        super().__init__()
        if name is not None:
            raise NotImplementedError("stub: attention_type=None only")
        self.attention = nn.Identity(**params)

    def forward(self, x):
        return self.attention(x)
