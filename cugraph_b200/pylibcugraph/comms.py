"""Multi-GPU bootstrap, one process per GPU (the role of pylibcugraph.comms / raft-dask:
python/pylibcugraph/pylibcugraph/comms/comms_wrapper.pyx:10-32).  The NCCL unique id is created on
rank 0 and distributed with torch.distributed; the library then builds its own communicators."""
import ctypes as C

from cugraph_b200 import _capi

_NCCL_ID_BYTES = 128


class Comm:
    def __init__(self, ptr):
        self.ptr = ptr

    def __del__(self):
        try:
            if self.ptr:
                _capi.lib().cugraph_b200_comm_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


def init_from_torch_distributed():
    """Build a communicator spanning torch.distributed's WORLD (must be initialised, cuda device set).
    Returns a Comm whose `.ptr` is passed to ResourceHandle(handle_ptr=...)."""
    import torch
    import torch.distributed as dist
    L = _capi.lib()
    rank, size = dist.get_rank(), dist.get_world_size()
    buf = (C.c_byte * _NCCL_ID_BYTES)()
    if rank == 0:
        err = C.c_void_p()
        code = L.cugraph_b200_get_nccl_unique_id(C.cast(buf, C.c_void_p), C.byref(err))
        _capi.check(code, err, "cugraph_b200_get_nccl_unique_id")
    t = torch.tensor(list(bytes(buf)), dtype=torch.uint8)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.broadcast(t, src=0)
    raw = bytes(t.cpu().tolist())
    buf2 = (C.c_byte * _NCCL_ID_BYTES).from_buffer_copy(raw)
    comm = C.c_void_p()
    err = C.c_void_p()
    code = L.cugraph_b200_comm_create(C.cast(buf2, C.c_void_p), rank, size, C.byref(comm), C.byref(err))
    _capi.check(code, err, "cugraph_b200_comm_create")
    return Comm(comm.value)
