"""Pre-launch tensor validation, same contract and error behaviour as
curobo/_src/curobolib/cuda_ops/tensor_checks.py:20-83 (raise ValueError; never call .contiguous())."""
from __future__ import annotations

import torch


def require_cuda(device: torch.device, message: str) -> None:
    """The product has no CPU path: every operator and every launcher refuses a non-CUDA device through this one function.
    (tests/test_simt_emulation_cpu.py replaces it -- together with the native library -- to run the host layer against the
    kernels compiled for the host SIMT emulation; nothing in the product does.)"""
    if device.type != "cuda":
        raise ValueError(message)


def check_tensors(device: torch.device, dtype: torch.dtype, **tensors: torch.Tensor) -> None:
    for name, t in tensors.items():
        if t is None:
            raise ValueError(f"{name}: expected a tensor, got None")
        if t.device != device:
            raise ValueError(f"{name}: expected device {device}, got {t.device}")
        require_cuda(device, f"{name}: curobo_b200 kernels are CUDA-only (sm_100a); got device {t.device}")
        if not t.is_contiguous():
            raise ValueError(f"{name}: expected contiguous tensor, got strides={t.stride()} for shape={tuple(t.shape)}")
        if t.dtype != dtype:
            raise ValueError(f"{name}: expected dtype {dtype}, got {t.dtype}")


def stream_ptr(device: torch.device) -> int:
    """The CURRENT torch stream of `device` (never the default stream; graph-capture safe)."""
    return _stream_of(device)


def _stream_of(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream
