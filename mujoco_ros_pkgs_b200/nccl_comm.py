"""Raw NCCL communicator for b2mj_allgather_publish*, created from the libnccl that torch ships.

The C-ABI takes an `ncclComm_t` as void* (include/b2mj.h): the host application owns the communicator.  torch's
ProcessGroupNCCL does not expose its own, so the harness (bench.py, tests, tools) bootstraps one here: rank 0 makes a
unique id, torch.distributed broadcasts its 128 bytes, every rank calls ncclCommInitRank.  Plumbing only."""
from __future__ import annotations

import ctypes as C
import glob
import os


class _UniqueId(C.Structure):
    _fields_ = [("internal", C.c_byte * 128)]


def _load_nccl():
    try:
        import nvidia.nccl  # torch's bundled wheel

        cands = sorted(glob.glob(os.path.join(list(nvidia.nccl.__path__)[0], "lib", "libnccl.so*")))
        if cands:
            return C.CDLL(cands[0], mode=C.RTLD_GLOBAL)
    except Exception:
        pass
    return C.CDLL("libnccl.so.2", mode=C.RTLD_GLOBAL)


class NcclComm:
    def __init__(self, rank: int, world: int):
        import torch
        import torch.distributed as dist

        self.lib = _load_nccl()
        uid = _UniqueId()
        if rank == 0:
            rc = self.lib.ncclGetUniqueId(C.byref(uid))
            assert rc == 0, f"ncclGetUniqueId failed: {rc}"
        t = torch.frombuffer(bytearray(bytes(uid.internal)), dtype=torch.uint8).cuda()
        dist.broadcast(t, 0)
        C.memmove(C.byref(uid), bytes(t.cpu().numpy().tobytes()), 128)
        self.comm = C.c_void_p()
        self.lib.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, _UniqueId, C.c_int]
        rc = self.lib.ncclCommInitRank(C.byref(self.comm), world, uid, rank)
        assert rc == 0, f"ncclCommInitRank failed: {rc}"
        self.rank, self.world = rank, world

    @property
    def ptr(self):
        return self.comm

    def destroy(self):
        if self.comm:
            self.lib.ncclCommDestroy.argtypes = [C.c_void_p]
            self.lib.ncclCommDestroy(self.comm)
            self.comm = None
