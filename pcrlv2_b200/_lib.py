"""ctypes binding of libpcrl_b200.so (include/pcrl_b200.h).

The library is the product: there is NO fallback.  If the shared object is missing or a call
fails, an exception is raised; nothing in this package routes around the CUDA path.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# PCRL_B200_LIB selects another build of the same library (e.g. the stall-accounting build of
# tools/stall_report.py); it is still this CUDA library, never a fallback.
LIB_PATH = os.environ.get("PCRL_B200_LIB") or os.path.join(_HERE, "libpcrl_b200.so")

_P = ctypes.c_void_p
_I = ctypes.c_int
_L = ctypes.c_longlong
_F = ctypes.c_float
_D = ctypes.c_double

# name -> argtypes, mirroring include/pcrl_b200.h (the header is the source of truth; the
# "not gpu" test-suite checks that every symbol declared there is exported and listed here).
SIGNATURES = {
    "pcrl_pack_conv3_weights": [_P, _P, _P, _I, _I, _I, _P],
    "pcrl_unpack_conv3_wgrad": [_P, _P, _I, _I, _P],
    "pcrl_pack_convT_weights": [_P, _P, _P, _I, _I, _I, _P],
    "pcrl_unpack_convT_wgrad": [_P, _P, _I, _I, _P],
    "pcrl_conv3d_k3_fprop": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "pcrl_conv3d_k3_dgrad": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "pcrl_conv3d_k3_dgrad_unshuffled": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "pcrl_conv3d_k3_wgrad": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "pcrl_stem_conv_fprop": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "pcrl_stem_conv_wgrad": [_P, _P, _P, _I, _I, _I, _I, _I, _P],
    "pcrl_convT3d_k2s2_fprop": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "pcrl_convT3d_k2s2_bwd": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "pcrl_norm_finalize": [_P, _D, _P, _P, _P, _P, _P, _P, _F, _F, _P, _P, _P, _P, _I, _I, _P],
    "pcrl_norm_act_fwd": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "pcrl_norm_act_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _D, _I, _I, _I, _I,
                          _I, _I, _I, _I, _I, _I, _P],
    "pcrl_zero_pad_rows": [_P, _L, _I, _L, _P],
    "pcrl_head_pack_weights": [_P, _P, _P, _P, _I, _I, _P],
    "pcrl_head_gather": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "pcrl_head_scatter": [_P, _P, _P, _I, _I, _I, _I, _I, _P],
    "pcrl_chan1_sigmoid_fwd": [_P, _P, _P, _P, _I, _I, _L, _P],
    "pcrl_chan1_sigmoid_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _D, _I, _I, _I, _L, _P],
    "pcrl_im2col27": [_P, _P, _I, _I, _I, _I, _I, _P],
    "pcrl_gemm_nt": [_P, _P, _P, _P, _L, _I, _I, _I, _I, _I, _P],
    "pcrl_gemm_nt_stats": [_P, _P, _P, _P, _L, _I, _I, _I, _P],
    "pcrl_gemm_tn": [_P, _P, _P, _L, _I, _I, _I, _P],
    "pcrl_sgd_flat": [_P, _P, _P, _P, _P, _P, _I, _F, _F, _F, _F, _P],
    "pcrl_split3_tf32": [_P, _P, _L, _I, _I, _I, _P],
    "pcrl_mse_scaled_fwd": [_P, _P, _P, _P, _L, _P],
    "pcrl_mse_scaled_bwd": [_P, _P, _P, _P, _P, _L, _P],
    "pcrl_contrastive_fwd_bwd": [_P, _P, _I, _I, _P, _P, _F, _P],
    "pcrl_contrastive_fwd_bwd_s": [_P, _P, _I, _I, _I, _P, _P, _F, _P],
    "pcrl_sgd_flat_dev": [_P, _P, _P, _P, _P, _P, _I, _P, _P, _P],
    "pcrl_aug_flip": [_P, _P, _P, _I, _I, _I, _I, _P],
    "pcrl_aug_blur_axis": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "pcrl_aug_noise_gamma": [_P, _P, _P, _P, _P, ctypes.c_ulonglong, _I, _I, _P],
    "pcrl_aug_swap": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "pcrl_aug_znorm": [_P, _P, _I, _I, _P],
    "pcrl_hu_window": [_P, _P, _L, _D, _D, _P],
    "pcrl_depth_scan": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _P],
    "pcrl_bn1d_fwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _F, _P],
    "pcrl_bn1d_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "pcrl_linear_fwd": [_P, _P, _P, _P, _I, _I, _I, _P],
    "pcrl_linear_bwd": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _P],
    "pcrl_cosine_mean_fwd_bwd": [_P, _P, _P, _P, _I, _I, _F, _F, _P],
    "pcrl_mse_fwd": [_P, _P, _P, _L, _P],
    "pcrl_mse_bwd": [_P, _P, _P, _P, _L, _P],
    "pcrl_sigmoid_fwd": [_P, _P, _L, _P],
    "pcrl_sigmoid_bwd": [_P, _P, _P, _L, _P],
    "pcrl_upsample_trilinear_fwd": [_P, _P, _I, _I, _I, _I, _I, _P],
    "pcrl_upsample_trilinear_bwd": [_P, _P, _I, _I, _I, _I, _I, _P],
    # 2-D path (planar.cu)
    "pcrl_im2col2d": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "pcrl_col2im2d": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "pcrl_maxpool2d_3x3s2_fwd": [_P, _P, _I, _I, _I, _I, _I, _P],
    "pcrl_maxpool2d_3x3s2_bwd": [_P, _P, _P, _I, _I, _I, _I, _I, _P],
    "pcrl_add_relu": [_P, _P, _P, _L, _I, _I, _P],
    "pcrl_upsample_nearest2x_fwd": [_P, _P, _I, _I, _I, _I, _I, _P],
    "pcrl_upsample_nearest2x_bwd": [_P, _P, _I, _I, _I, _I, _I, _P],
    "pcrl_bilinear2d_fwd": [_P, _P, _I, _I, _I, _I, _P],
    "pcrl_bilinear2d_bwd": [_P, _P, _I, _I, _I, _I, _P],
    "pcrl_conv2d_c3_fwd": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "pcrl_conv2d_c3_bwd": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "pcrl_pack_conv2d_weights": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "pcrl_unpack_conv2d_wgrad": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
}

_lib = None


class PcrlError(RuntimeError):
    pass


def build(force: bool = False) -> str:
    """Compile the CUDA sources in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    script = os.path.join(_HERE, "csrc", "build.sh")
    if force:
        for f in os.listdir(os.path.join(_HERE, "csrc", "build")) if os.path.isdir(
                os.path.join(_HERE, "csrc", "build")) else []:
            if f.endswith(".o"):
                os.remove(os.path.join(_HERE, "csrc", "build", f))
    subprocess.check_call(["bash", script])
    return LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PcrlError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
                f"g.build()'` (or pcrlv2_b200/csrc/build.sh).  There is no fallback path.")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.pcrl_last_error.restype = ctypes.c_char_p
        _lib.pcrl_version.restype = _I
        for name, argtypes in SIGNATURES.items():
            fn = getattr(_lib, name)
            fn.argtypes = argtypes
            fn.restype = _I
    return _lib


def _conv(a):
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        if not a.is_cuda:
            raise PcrlError("libpcrl_b200 was handed a CPU tensor; it only runs on the GPU")
        return a.data_ptr()
    return a


# kernels launched by one call of each entry point (for bench.py's gpu_launches count)
LAUNCHES = {"pcrl_convT3d_k2s2_fprop": 2, "pcrl_convT3d_k2s2_bwd": 3, "pcrl_conv3d_k3_dgrad_unshuffled": 2,
            "pcrl_linear_bwd": 3}
launch_count = [0]
# when set to a list, every call is bracketed by CUDA events on the launching stream and
# (name, int-args, start, end) is appended -- bench.py uses this for the per-kernel roofline
profile = [None]


def call(name: str, *args) -> None:
    """Invoke an entry point on torch's current CUDA stream; the stream argument is appended."""
    fn = getattr(lib(), name)
    cargs = [_conv(a) for a in args]          # raises on CPU tensors before touching CUDA
    stream = torch.cuda.current_stream().cuda_stream
    prof = profile[0]
    if prof is not None:
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = fn(*cargs, stream)
    if rc != 0:
        raise PcrlError(f"{name} failed ({rc}): {lib().pcrl_last_error().decode()}")
    launch_count[0] += LAUNCHES.get(name, 1)
    if prof is not None:
        e1.record()
        prof.append((name, tuple(a for a in args if isinstance(a, int)), e0, e1))
