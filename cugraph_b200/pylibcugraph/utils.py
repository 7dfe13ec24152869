"""Array plumbing (reference python/pylibcugraph/pylibcugraph/utils.pyx: assert_CAI_type :85-95,
view creation :235-246, copy_to_cupy_array :162-196 — here the copy lands in a torch tensor)."""
import ctypes as C

import numpy as np

from cugraph_b200 import _capi

_TYPESTR_TO_ID = {"<i4": _capi.INT32, "<i8": _capi.INT64, "<f4": _capi.FLOAT32, "<f8": _capi.FLOAT64,
                  "<u4": _capi.UINT32, "<u8": _capi.UINT64, "|i1": _capi.INT8, "|u1": _capi.UINT8,
                  "<i2": _capi.INT16, "<u2": _capi.UINT16, "|b1": _capi.BOOL}


def assert_CAI_type(obj, var_name, allow_none=False):
    if allow_none and obj is None:
        return
    if not hasattr(obj, "__cuda_array_interface__"):
        msg = f"{var_name} must be a device array (__cuda_array_interface__)"
        if allow_none:
            msg += " or None"
        raise TypeError(msg)


class View:
    """RAII wrapper over cugraph_type_erased_device_array_view_t for a python device array."""

    def __init__(self, obj):
        self.ptr = None
        self._keep = obj
        if obj is None:
            return
        cai = obj.__cuda_array_interface__
        if cai.get("strides") is not None:
            # only contiguous 1-D arrays are accepted, like the reference
            itemsize = np.dtype(cai["typestr"]).itemsize
            if tuple(cai["strides"]) != (itemsize,) and int(np.prod(cai["shape"])) > 1:
                raise ValueError("device array must be contiguous")
        n = int(np.prod(cai["shape"])) if len(cai["shape"]) else 1
        tid = _TYPESTR_TO_ID.get(cai["typestr"])
        if tid is None:
            raise TypeError(f"unsupported dtype {cai['typestr']}")
        self.ptr = _capi.lib().cugraph_type_erased_device_array_view_create(C.c_void_p(cai["data"][0] or 0), n, tid)

    def free(self):
        if self.ptr:
            _capi.lib().cugraph_type_erased_device_array_view_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


_ID_TO_TORCH = None


def _torch_dtype(tid):
    global _ID_TO_TORCH
    import torch
    if _ID_TO_TORCH is None:
        _ID_TO_TORCH = {_capi.INT32: torch.int32, _capi.INT64: torch.int64, _capi.FLOAT32: torch.float32,
                        _capi.FLOAT64: torch.float64, _capi.UINT8: torch.uint8, _capi.INT8: torch.int8,
                        _capi.BOOL: torch.bool}
    return _ID_TO_TORCH[tid]


def copy_to_torch(handle, view_ptr):
    """Copy a result view into a freshly allocated torch CUDA tensor and free the view."""
    import torch
    L = _capi.lib()
    n = L.cugraph_type_erased_device_array_view_size(view_ptr)
    tid = L.cugraph_type_erased_device_array_view_type(view_ptr)
    out = torch.empty(n, dtype=_torch_dtype(tid), device="cuda")
    if n:
        dst = L.cugraph_type_erased_device_array_view_create(C.c_void_p(out.data_ptr()), n, tid)
        err = C.c_void_p()
        code = L.cugraph_type_erased_device_array_view_copy(handle.ptr, dst, view_ptr, C.byref(err))
        L.cugraph_type_erased_device_array_view_free(dst)
        _capi.check(code, err, "cugraph_type_erased_device_array_view_copy()")
    L.cugraph_type_erased_device_array_view_free(view_ptr)
    return out
