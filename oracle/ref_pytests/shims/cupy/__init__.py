"""TEST INFRASTRUCTURE (oracle/ref_pytests): the two things the reference's pylibcugraph tests need from cupy — `asarray`
for building device arrays and array objects that answer `.dtype`, `.tolist()`, `len()` — served by torch tensors (CUDA
when a device is present, host memory when the library under test is the CPU emulation build).  cupy is not installed in
this image; the product never imports this module."""
import numpy as _np
import torch as _torch


def _device():
    return "cuda" if _torch.cuda.is_available() else "cpu"


def asarray(obj, dtype=None):
    a = _np.asarray(obj if not isinstance(obj, range) else list(obj), dtype=dtype)
    return _torch.as_tensor(_np.array(a, copy=True)).to(_device())


array = asarray
float32, float64, int32, int64 = _np.float32, _np.float64, _np.int32, _np.int64


# cupy arrays answer .get() with a numpy copy; results of the mirror are torch tensors (test infrastructure only)
if not hasattr(_torch.Tensor, "get"):
    _torch.Tensor.get = lambda self: self.detach().cpu().numpy()


def zeros(shape, dtype=float):
    return _torch.as_tensor(_np.zeros(shape, dtype=dtype)).to(_device())


ndarray = _torch.Tensor


def union1d(a, b):
    return _torch.unique(_torch.cat([a.reshape(-1), b.reshape(-1)]))


def _astype(self, dtype=None, **kw):
    return self.to(getattr(_torch, _np.dtype(dtype).name))


if not hasattr(_torch.Tensor, "astype"):
    _torch.Tensor.astype = _astype
