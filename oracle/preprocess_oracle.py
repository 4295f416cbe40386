"""CPU oracle of pcrlv2_b200/preprocess.py.  TEST INFRASTRUCTURE ONLY.
Restates /root/reference/luna_preprocess.py: the HU window (:132-135), the depth-scan loops (:213-241, kept as
the reference's own nested Python loops -- small cases only -- next to a vectorised numpy form checked against
them) and cal_iou (:295-320).  The reference file itself cannot be imported here (SimpleITK, skimage absent)."""
import numpy as np

HU_MIN, HU_MAX = -1000.0, 1000.0
HU_THRED = (-150.0 - HU_MIN) / (HU_MAX - HU_MIN)


def hu_window(img):
    img = img.astype(np.float64).copy()
    img[img < HU_MIN] = HU_MIN
    img[img > HU_MAX] = HU_MAX
    return 1.0 * (img - HU_MIN) / (HU_MAX - HU_MIN)


def depth_scan_loops(crop, depth, len_depth=3, thr=HU_THRED):
    """luna_preprocess.py:213-241, loop for loop."""
    rows, cols = crop.shape[:2]
    t_img = np.zeros((rows, cols, depth), dtype=float)
    d_img = np.zeros((rows, cols, depth), dtype=float)
    for d in range(depth):
        for i in range(rows):
            for j in range(cols):
                for k in range(len_depth):
                    if crop[i, j, d + k] >= thr:
                        t_img[i, j, d] = crop[i, j, d + k]
                        d_img[i, j, d] = k
                        break
                    if k == len_depth - 1:
                        d_img[i, j, d] = k
    d_img = d_img.astype("float32")
    d_img /= (len_depth - 1)
    d_img = 1.0 - d_img
    return t_img, d_img


def depth_scan(crop, depth, len_depth=3, thr=HU_THRED):
    """Vectorised form of the loops above."""
    win = np.stack([crop[:, :, k:k + depth] for k in range(len_depth)], axis=-1)     # [X, Y, depth, len_depth]
    hit = win >= thr
    first = np.where(hit.any(-1), hit.argmax(-1), len_depth - 1)
    t_img = np.where(hit.any(-1), np.take_along_axis(win, first[..., None], -1)[..., 0], 0.0)
    d_img = 1.0 - first.astype("float32") / (len_depth - 1)
    return t_img, d_img


def cal_iou(box1, box2):
    xmin1, xmax1, ymin1, ymax1, zmin1, zmax1 = box1
    xmin2, xmax2, ymin2, ymax2, zmin2, zmax2 = box2
    s1 = (xmax1 - xmin1) * (ymax1 - ymin1) * (zmax1 - zmin1)
    s2 = (xmax2 - xmin2) * (ymax2 - ymin2) * (zmax2 - zmin2)
    xmin, ymin, zmin = max(xmin1, xmin2), max(ymin1, ymin2), max(zmin1, zmin2)
    xmax, ymax, zmax = min(xmax1, xmax2), min(ymax1, ymax2), min(zmax1, zmax2)
    w, h, d = max(0, xmax - xmin), max(0, ymax - ymin), max(0, zmax - zmin)
    area = w * h * d
    return area / (s1 + s2 - area)
