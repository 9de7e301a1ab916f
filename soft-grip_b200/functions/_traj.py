"""Shared argument handling for the trajectory kernels (sg_traj_* in include/softgrip.h)."""
import ctypes as C

from .._lib import SoftGripError


def traj_args(data):
    """data: CUDA tensor (..., C), float32/float64, contiguous -> (torch, nrows, nchan, precision, device index, stream)."""
    import torch
    if not isinstance(data, torch.Tensor) or not data.is_cuda:
        raise SoftGripError("trajectory kernels need a CUDA tensor (libsoftgrip has no CPU path)")
    if data.dtype not in (torch.float32, torch.float64):
        raise TypeError("trajectory must be float32 or float64")
    if data.dim() < 2 or not data.is_contiguous():
        raise ValueError("trajectory must be a contiguous (..., C) tensor")
    nchan = int(data.shape[-1])
    nrows = int(data.numel() // max(1, nchan))
    dev = data.device.index if data.device.index is not None else torch.cuda.current_device()
    stream = C.c_void_p(torch.cuda.current_stream(data.device).cuda_stream)
    return torch, nrows, nchan, (32 if data.dtype == torch.float32 else 64), int(dev), stream


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None
