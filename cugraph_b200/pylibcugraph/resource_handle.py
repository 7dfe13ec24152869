"""ResourceHandle (reference python/pylibcugraph/pylibcugraph/resource_handle.pyx)."""
from cugraph_b200 import _capi


class ResourceHandle:
    """Owns a cugraph_resource_handle_t.  `handle_ptr` must be None (single-GPU handle on the current device): the
    reference passes a raft handle with NCCL comms here, which this build has no counterpart for (multi-GPU: cugraph_b200.mg)."""

    def __init__(self, handle_ptr=None, stream=None):
        self._lib = _capi.lib()
        self._stream = stream
        if stream is not None:
            # bind to the caller's CUDA stream (e.g. torch.cuda.current_stream().cuda_stream): library
            # kernels and torch.distributed collectives are then ordered without host synchronisation
            self._ptr = self._lib.cugraph_b200_create_resource_handle_on_stream(stream)
        else:
            self._ptr = self._lib.cugraph_create_resource_handle(handle_ptr)
        if not self._ptr:
            raise RuntimeError("cugraph_create_resource_handle failed (is a CUDA device available?)")

    @property
    def ptr(self):
        return self._ptr

    def get_rank(self):
        return self._lib.cugraph_resource_handle_get_rank(self._ptr)

    def get_comm_size(self):
        return self._lib.cugraph_resource_handle_get_comm_size(self._ptr)

    def order_after_caller(self):
        """Called by every wrapper before it hands tensors to the library.  A default handle owns its own
        non-blocking CUDA stream, so work the caller queued on torch's current stream (asynchronous H2D
        copies from pinned memory, generator kernels) is NOT ordered before the library's kernels: wait for
        it.  A handle bound to the caller's stream needs nothing.  (C callers: same contract as the
        reference — inputs must be complete, or ordered with respect to the handle's stream.)"""
        import torch
        if not torch.cuda.is_available():
            return
        cur = torch.cuda.current_stream()
        if self._stream is not None and int(self._stream) == int(cur.cuda_stream):
            return
        cur.synchronize()

    def launch_count(self):
        return int(self._lib.cugraph_b200_handle_launch_count(self._ptr))

    def __del__(self):
        try:
            if getattr(self, "_ptr", None):
                self._lib.cugraph_free_resource_handle(self._ptr)
                self._ptr = None
        except Exception:
            pass
