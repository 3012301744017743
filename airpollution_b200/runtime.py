"""Device runtime glue: one library context per CUDA device, bound to torch's
current stream, plus the small helpers that hand torch tensors to the C ABI as
raw device pointers.  torch is used for device memory, streams and (multi-GPU)
process groups only; no torch operator runs on the solve path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


class Runtime:
    _instances: dict = {}

    def __init__(self, device: torch.device):
        if not torch.cuda.is_available():
            raise RuntimeError("airpollution_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = device
        self.lib = _lib.load()
        torch.cuda.init()
        with torch.cuda.device(device):
            torch.cuda.current_stream()  # make sure the primary context exists
            h = C.c_void_p()
            _lib.call("crbe_ctx_create", device.index or 0, C.byref(h))
        self.ctx = h
        self._bound = None

    @classmethod
    def get(cls, device=None) -> "Runtime":
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)
        device = torch.device(device)
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        key = device.index
        if key not in cls._instances:
            cls._instances[key] = Runtime(device)
        rt = cls._instances[key]
        rt.bind_stream()
        return rt

    def bind_stream(self):
        """Enqueue library work on torch's current stream of this device."""
        s = torch.cuda.current_stream(self.device).cuda_stream
        if s != self._bound:
            _lib.call("crbe_ctx_set_stream", self.ctx, C.c_void_p(s))
            self._bound = s

    def call(self, name, *args):
        _lib.call(name, *args)

    def synchronize(self):
        _lib.call("crbe_ctx_synchronize", self.ctx)

    # ---- buffers -------------------------------------------------------
    def empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.device)

    def zeros(self, shape, dtype):
        return torch.zeros(shape, dtype=dtype, device=self.device)

    def upload(self, array, dtype=None):
        a = np.ascontiguousarray(array if dtype is None else np.asarray(array).astype(dtype, copy=False))
        return torch.from_numpy(a).to(self.device)


def ptr(t):
    """Raw device address of a (contiguous) tensor, or NULL for None."""
    if t is None:
        return C.c_void_p(0)
    if not t.is_contiguous():
        raise ValueError("device buffer handed to libcrbe_b200 must be contiguous")
    return C.c_void_p(t.data_ptr())


def to_numpy(t):
    return t.detach().cpu().numpy()
