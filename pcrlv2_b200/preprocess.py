"""The deterministic pieces of the reference's offline crop generator (luna_preprocess.py:132-275) on the GPU
(SURVEY 8f row 4): the HU window, the box IoU test of a crop pair and the per-voxel depth scan whose 4-deep
Python loop (:217-236, 64*64*32*3 iterations per crop) dominates that script.  Not built: reading the LUNA
.mhd volumes (SimpleITK) and the skimage `resize` of crops whose drawn size is not 64x64x32.
"""
import torch

from . import _lib

HU_MIN, HU_MAX = -1000.0, 1000.0                     # setup_config, luna_preprocess.py:63-66
HU_THRED = (-150.0 - HU_MIN) / (HU_MAX - HU_MIN)
LEN_DEPTH, LUNG_MAX = 3, 0.15                        # config instance, :118-121


def hu_window(vol):
    """vol: fp32 CUDA tensor of HU values -> [0, 1] (luna_preprocess.py:133-135)."""
    assert vol.is_cuda and vol.dtype == torch.float32 and vol.is_contiguous()
    out = torch.empty_like(vol)
    _lib.call("pcrl_hu_window", vol, out, vol.numel(), HU_MIN, HU_MAX)
    return out


def cal_iou(box1, box2):
    """luna_preprocess.py:295-320; boxes are (xmin, xmax, ymin, ymax, zmin, zmax)."""
    xmin1, xmax1, ymin1, ymax1, zmin1, zmax1 = box1
    xmin2, xmax2, ymin2, ymax2, zmin2, zmax2 = box2
    s1 = (xmax1 - xmin1) * (ymax1 - ymin1) * (zmax1 - zmin1)
    s2 = (xmax2 - xmin2) * (ymax2 - ymin2) * (zmax2 - zmin2)
    w = max(0, min(xmax1, xmax2) - max(xmin1, xmin2))
    h = max(0, min(ymax1, ymax2) - max(ymin1, ymin2))
    d = max(0, min(zmax1, zmax2) - max(zmin1, zmin2))
    area = w * h * d
    return area / (s1 + s2 - area)


def depth_scan(crop, depth, len_depth=LEN_DEPTH, threshold=HU_THRED):
    """crop [X, Y, depth + len_depth (or more)] fp32 CUDA -> (t_img, d_img [X, Y, depth], sum(d_img) as a 0-dim
    fp64 tensor): luna_preprocess.py:213-241."""
    assert crop.is_cuda and crop.dtype == torch.float32 and crop.is_contiguous() and crop.dim() == 3
    x, y, zp = crop.shape
    t_img = torch.empty((x, y, depth), dtype=torch.float32, device=crop.device)
    d_img = torch.empty_like(t_img)
    total = torch.zeros((), dtype=torch.float64, device=crop.device)
    _lib.call("pcrl_depth_scan", crop, t_img, d_img, total, x, y, depth, zp, len_depth, float(threshold))
    return t_img, d_img, total


def accept_crop(crop, crop_rows, crop_cols, crop_deps, depth=32):
    """The lung-fraction test of :243-247: reject when sum(d_img) > lung_max * rows * cols * deps."""
    _, _, total = depth_scan(crop, depth)
    return bool(total.item() <= LUNG_MAX * crop_rows * crop_cols * crop_deps)
