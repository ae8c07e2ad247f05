"""
Host/device buffer plumbing for the ctypes layer.

* numpy arrays (and anything ``numpy.asarray`` accepts) are HOST buffers: the library stages them through the GPU
  itself and the result comes back as a numpy array;
* ``torch`` CUDA tensors, objects exposing ``__cuda_array_interface__`` (cupy, numba, ...) and DLPack capsules /
  ``__dlpack__`` objects are DEVICE buffers: they are used in place (zero copy) and the result is a ``torch`` tensor on
  the same device, computed asynchronously on torch's current stream.

PyTorch is used for device memory and streams only (allocation of outputs, current-stream lookup).
"""

import numpy as np


def _torch():
    import torch
    return torch


def is_device_array(x):
    mod = type(x).__module__
    if mod.startswith('torch'):
        return bool(getattr(x, 'is_cuda', False))
    return hasattr(x, '__cuda_array_interface__') or (hasattr(x, '__dlpack__') and not isinstance(x, np.ndarray)
                                                      and _dlpack_on_cuda(x))


def _dlpack_on_cuda(x):
    try:
        dev_type, _ = x.__dlpack_device__()
        return int(dev_type) == 2   # kDLCUDA
    except Exception:
        return False


class Buffer(object):
    """A contiguous fp64 (or complex128) buffer, on host or device."""

    __slots__ = ('obj', 'ptr', 'shape', 'on_device', 'device')

    def __init__(self, obj, ptr, shape, on_device, device):
        self.obj, self.ptr, self.shape, self.on_device, self.device = obj, ptr, tuple(shape), on_device, device


def as_input(x, dtype='f8'):
    """Wrap ``x`` (never copies a contiguous fp64 array)."""
    if is_device_array(x):
        torch = _torch()
        if not isinstance(x, torch.Tensor):
            x = torch.as_tensor(x, device='cuda') if hasattr(x, '__cuda_array_interface__') else torch.from_dlpack(x)
        tdtype = {'f8': torch.float64, 'c16': torch.complex128}[dtype]
        if x.is_complex() and dtype == 'f8':
            raise ValueError('complex input is not supported, pass real and imaginary parts separately')
        x = x.to(tdtype).contiguous()
        return Buffer(x, x.data_ptr(), x.shape, True, x.device.index)
    if type(x).__module__.startswith('torch'):
        x = x.detach().numpy()
    x = np.asarray(x)
    if np.iscomplexobj(x) and dtype == 'f8':
        raise ValueError('complex input is not supported, pass real and imaginary parts separately')
    x = np.ascontiguousarray(x, dtype=dtype)
    return Buffer(x, x.ctypes.data, x.shape, False, None)


def empty_like_kind(ref, shape, dtype='f8'):
    """Allocate an output buffer of the same kind (host numpy / device torch) as ``ref``."""
    if ref.on_device:
        torch = _torch()
        tdtype = {'f8': torch.float64, 'c16': torch.complex128}[dtype]
        out = torch.empty(tuple(shape), dtype=tdtype, device=torch.device('cuda', ref.device))
        return Buffer(out, out.data_ptr(), shape, True, ref.device)
    out = _host_empty(tuple(shape), dtype)
    return Buffer(out, out.ctypes.data, shape, False, None)


_PINNED_MIN_BYTES = 4 << 20


def _host_empty(shape, dtype):
    """Host result buffer.  Large results are page-locked (through torch's caching host allocator) so that the
    device-to-host copy runs at full PCIe rate and overlaps with compute; small ones are plain numpy arrays."""
    nbytes = int(np.prod(shape, dtype='i8')) * np.dtype(dtype).itemsize
    if nbytes >= _PINNED_MIN_BYTES:
        try:
            torch = _torch()
            tdtype = {'f8': torch.float64, 'c16': torch.complex128}[dtype]
            return torch.empty(shape, dtype=tdtype, pin_memory=True).numpy()
        except Exception:
            pass
    return np.empty(shape, dtype=dtype)


def current_stream(device):
    """cudaStream_t (as int) device work should be ordered on: torch's current stream for device buffers."""
    torch = _torch()
    return int(torch.cuda.current_stream(device).cuda_stream)


def default_device():
    """Device index for host-buffer calls: ``$CPF_DEVICE`` if set, else torch's current device if torch has CUDA
    initialised in this process, else 0."""
    import os
    import sys
    if os.environ.get('CPF_DEVICE'):
        return int(os.environ['CPF_DEVICE'])
    torch = sys.modules.get('torch', None)
    try:
        if torch is not None and torch.cuda.is_initialized():
            return int(torch.cuda.current_device())
    except Exception:
        pass
    return 0
