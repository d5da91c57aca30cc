"""torch is plumbing here: device memory, streams, torch.distributed.  No compute."""
import ctypes as C

import torch


def require_cuda(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("abstract_gym_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("abstract_gym_b200 only runs on CUDA devices, got %s" % device)
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return device


def ptr(t):
    """raw device/host pointer of a tensor (None -> NULL)"""
    if t is None:
        return None
    if not t.is_contiguous():
        raise ValueError("tensor must be contiguous")
    return C.c_void_p(t.data_ptr())


def stream_ptr(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def as_f64(x, device, n=None):
    t = torch.as_tensor(x, dtype=torch.float64, device=device).contiguous()
    if n is not None and t.numel() != n:
        raise ValueError("expected %d elements, got %d" % (n, t.numel()))
    return t
