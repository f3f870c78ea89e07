"""A raw NCCL communicator for the C ABI's rbis_batch_stats_allreduce, made from a torch.distributed process group.

torch.distributed does not hand out its ncclComm_t, so this creates a second communicator over the same ranks with the
NCCL library torch has already loaded (ctypes): rank 0 draws the unique id, torch.distributed broadcasts its 128 bytes,
every rank calls ncclCommInitRank.  Plumbing only -- a C++ host links NCCL and passes its own communicator
(tests/cpp/stats_allreduce.cpp)."""
import ctypes as C
import os


class _UniqueId(C.Structure):
    _fields_ = [("internal", C.c_char * 128)]


def _load():
    candidates = []
    try:
        import nvidia.nccl as pkg  # the NCCL wheel torch links against

        for d in list(getattr(pkg, "__path__", [])):
            candidates.append(os.path.join(d, "lib", "libnccl.so.2"))
    except Exception:
        pass
    candidates += ["libnccl.so.2", "libnccl.so"]
    for c in candidates:
        try:
            return C.CDLL(c, mode=C.RTLD_GLOBAL)
        except OSError:
            continue
    raise RuntimeError("no NCCL library found")


class Communicator:
    """comm.handle is the ncclComm_t (void*) of this rank."""

    def __init__(self, rank, world, device):
        import torch
        import torch.distributed as dist

        self.lib = _load()
        self.lib.ncclGetUniqueId.argtypes = [C.POINTER(_UniqueId)]
        self.lib.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, _UniqueId, C.c_int]
        self.lib.ncclCommDestroy.argtypes = [C.c_void_p]
        uid = _UniqueId()
        if rank == 0:
            rc = self.lib.ncclGetUniqueId(C.byref(uid))
            if rc != 0:
                raise RuntimeError(f"ncclGetUniqueId failed: {rc}")
        if world > 1:
            t = torch.frombuffer(bytearray(bytes(uid)), dtype=torch.uint8).clone()
            backend = dist.get_backend()
            t = t.to(torch.device("cuda", device)) if backend == "nccl" else t
            dist.broadcast(t, src=0)
            C.memmove(C.byref(uid), bytes(t.cpu().numpy().tobytes()), 128)
        torch.cuda.set_device(device)
        self.handle = C.c_void_p()
        rc = self.lib.ncclCommInitRank(C.byref(self.handle), int(world), uid, int(rank))
        if rc != 0:
            raise RuntimeError(f"ncclCommInitRank failed: {rc}")
        self.rank, self.world = rank, world

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle.value:
            self.lib.ncclCommDestroy(self.handle)
            self.handle = C.c_void_p()
