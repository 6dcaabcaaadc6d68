"""Raw NCCL communicator for the opt-in global-lighting reduction
(smvsb_fit_lighting's nccl_comm argument). torch.distributed does not expose
its ncclComm_t, so a second communicator is created on the NCCL library that
ships with torch: the unique id is made on rank 0 and broadcast through the
already-initialised torch.distributed group. The library is loaded
RTLD_GLOBAL so that libsmvs_b200.so finds ncclAllReduce with dlsym."""
from __future__ import annotations

import ctypes as C
import glob
import os

NCCL_UNIQUE_ID_BYTES = 128
_lib = None


def lib():
    global _lib
    if _lib is None:
        import torch
        cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia",
                                       "nccl", "lib", "libnccl.so*"))
        cands += ["libnccl.so.2"]
        last = None
        for c in cands:
            try:
                _lib = C.CDLL(c, mode=C.RTLD_GLOBAL)
                break
            except OSError as exc:      # noqa: PERF203
                last = exc
        if _lib is None:
            raise RuntimeError(f"libnccl not found: {last}")
        _lib.ncclGetErrorString.restype = C.c_char_p
    return _lib


def _check(rc):
    if rc != 0:
        raise RuntimeError("NCCL: " + lib().ncclGetErrorString(rc).decode())


def create_comm():
    """Collective over the default torch.distributed group; returns the
    ncclComm_t as an int (pass to Context.fit_lighting(nccl_comm=...))."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    uid = (C.c_char * NCCL_UNIQUE_ID_BYTES)()
    if rank == 0:
        _check(lib().ncclGetUniqueId(C.byref(uid)))
    t = torch.frombuffer(bytearray(uid.raw), dtype=torch.uint8).clone()
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.broadcast(t, src=0)
    raw = bytes(t.cpu().numpy().tobytes())
    uid2 = (C.c_char * NCCL_UNIQUE_ID_BYTES).from_buffer_copy(raw)
    comm = C.c_void_p()
    # ncclCommInitRank(ncclComm_t*, int nranks, ncclUniqueId commId (by value), int rank)
    class _Uid(C.Structure):
        _fields_ = [("internal", C.c_char * NCCL_UNIQUE_ID_BYTES)]
    u = _Uid()
    C.memmove(C.byref(u), uid2, NCCL_UNIQUE_ID_BYTES)
    lib().ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, _Uid, C.c_int]
    _check(lib().ncclCommInitRank(C.byref(comm), world, u, rank))
    return comm.value


def destroy_comm(comm):
    lib().ncclCommDestroy.argtypes = [C.c_void_p]
    _check(lib().ncclCommDestroy(C.c_void_p(comm)))
