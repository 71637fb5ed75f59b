"""Engine-level collectives for the one-process-per-GPU mode (``fbr_comm_*`` in include/fiber_b200.h).

The map needs no data-path collective: rank ``g`` owns ``shard.block_of(n, g, G)``.  What surrounds a map
does exchange data, and those steps belong to the engine, not to the caller's framework:

* ``broadcast``  -- shared arguments resident on one rank (the parzen sample block, ``initargs``);
* ``scatter`` / ``gather`` -- a map's input array / ordered result blocks resident on a root rank: the
  fan-out and fan-in of the reference's master sockets (fiber/pool.py:910-920) as grouped
  ``ncclSend/ncclRecv``;
* ``allgather``  -- every rank's ordered block -> the full ordered result everywhere;
* ``allreduce_i64`` -- scalar folds (the pi count); ``allreduce`` -- ``experimental.Ring``'s collective.

Bootstrap: rank 0 makes the 128-byte id (``unique_id()``) and publishes it -- a ``Ring`` puts it in its member
table instead of the reference's ip/port (fiber/experimental/ring.py:44-55); under ``torchrun`` any key-value
store does (``Comm.from_store``).  NCCL is the transport (NVLink 5 / NVSwitch); there is no CPU fallback.
"""
import ctypes
import os

import numpy as np

from . import _abi

U8, I32, I64, F32, F64 = range(5)
SUM, PROD, MAX, MIN = range(4)
ID_BYTES = 128
_NP_DTYPE = {np.dtype(np.uint8): U8, np.dtype(np.int32): I32, np.dtype(np.int64): I64, np.dtype(np.float32): F32, np.dtype(np.float64): F64}


def _check(status):
    if status != _abi.FBR_OK:
        raise _abi.EngineError(status, _abi.load().fbr_comm_last_error().decode("utf-8", "replace"))
    return status


def _nccl_candidates():
    """The NCCL build bundled with torch's wheels (what torch.distributed itself loads), if present."""
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        if spec and spec.submodule_search_locations:
            for loc in spec.submodule_search_locations:
                cand = os.path.join(loc, "lib", "libnccl.so.2")
                if os.path.exists(cand):
                    return cand
    except Exception:       # noqa: BLE001
        pass
    return None


_loaded = None


def load_nccl():
    """Bind NCCL (dlopen) and return its version number, e.g. 22809."""
    global _loaded
    if _loaded is None:
        v = ctypes.c_int(0)
        path = os.environ.get("FBR_NCCL_LIB") or _nccl_candidates()
        _check(_abi.load().fbr_comm_load(path.encode() if path else None, ctypes.byref(v)))
        _loaded = v.value
    return _loaded


def unique_id():
    """``ncclGetUniqueId``: the bootstrap handle rank 0 publishes (needs no GPU)."""
    load_nccl()
    buf = (ctypes.c_char * ID_BYTES)()
    _check(_abi.load().fbr_comm_unique_id(buf))
    return bytes(buf)


class DeviceBuffer:
    """A device allocation owned by a communicator (ring nodes have no pool to allocate from)."""

    def __init__(self, comm, nbytes):
        self.comm, self.nbytes = comm, int(nbytes)
        p = ctypes.c_void_p()
        _check(comm.lib.fbr_comm_device_alloc(comm.handle, self.nbytes, ctypes.byref(p)))
        self.ptr = p.value

    def upload(self, array):
        a = np.ascontiguousarray(array)
        assert a.nbytes <= self.nbytes
        _check(self.comm.lib.fbr_comm_memcpy_h2d(self.comm.handle, ctypes.c_void_p(self.ptr), a.ctypes.data, a.nbytes))
        return self

    def download(self, dtype=np.uint8, count=None):
        dtype = np.dtype(dtype)
        n = self.nbytes // dtype.itemsize if count is None else count
        out = np.empty(n, dtype=dtype)
        _check(self.comm.lib.fbr_comm_memcpy_d2h(self.comm.handle, out.ctypes.data, ctypes.c_void_p(self.ptr), out.nbytes))
        return out

    def free(self):
        if self.ptr and self.comm.handle:
            self.comm.lib.fbr_comm_device_free(self.comm.handle, ctypes.c_void_p(self.ptr))
        self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:       # noqa: BLE001
            pass


def _ptr(x):
    if isinstance(x, DeviceBuffer):
        return ctypes.c_void_p(x.ptr)
    if isinstance(x, ctypes.c_void_p):
        return x
    if hasattr(x, "data_ptr"):              # a torch CUDA tensor: plumbing, only its address is used
        return ctypes.c_void_p(x.data_ptr())
    return ctypes.c_void_p(int(x))


class Comm:
    """One rank of an NCCL communicator bound to one CUDA device (``fbr_comm_t``)."""

    def __init__(self, device, nranks, rank, id_bytes):
        load_nccl()
        self.lib = _abi.load()
        assert len(id_bytes) == ID_BYTES
        h = ctypes.c_void_p()
        _check(self.lib.fbr_comm_create(int(device), int(nranks), int(rank), id_bytes, ctypes.byref(h)))
        self.handle, self.device, self.rank, self.nranks = h, int(device), int(rank), int(nranks)

    @classmethod
    def from_store(cls, store, device, nranks, rank, key="fbr_comm_id"):
        """Bootstrap through any key-value store with ``set(key, bytes)`` / ``get(key)`` (e.g. the c10d TCPStore
        ``torchrun`` already provides): rank 0 publishes the id, the others read it."""
        if rank == 0:
            store.set(key, unique_id())
        raw = store.get(key)
        return cls(device, nranks, rank, bytes(raw))

    def alloc(self, nbytes):
        return DeviceBuffer(self, nbytes)

    def sync(self):
        _check(self.lib.fbr_comm_sync(self.handle))

    def broadcast(self, buf, nbytes, root=0):
        _check(self.lib.fbr_comm_broadcast(self.handle, _ptr(buf), int(nbytes), int(root)))

    def allgather(self, send, recv, bytes_per_rank):
        _check(self.lib.fbr_comm_allgather(self.handle, _ptr(send), _ptr(recv), int(bytes_per_rank)))

    def gather(self, send, recv_on_root, bytes_per_rank, root=0):
        _check(self.lib.fbr_comm_gather(self.handle, _ptr(send), _ptr(recv_on_root) if recv_on_root is not None else None,
                                        int(bytes_per_rank), int(root)))

    def scatter(self, send_on_root, recv, bytes_per_rank, root=0):
        _check(self.lib.fbr_comm_scatter(self.handle, _ptr(send_on_root) if send_on_root is not None else None, _ptr(recv),
                                         int(bytes_per_rank), int(root)))

    def allreduce(self, send, recv, count, dtype=F32, op=SUM):
        _check(self.lib.fbr_comm_allreduce(self.handle, _ptr(send), _ptr(recv), int(count), int(dtype), int(op)))

    def allreduce_timed(self, buf, count, dtype=F32, op=SUM, iters=10):
        """``iters`` in-place all-reduces, CUDA-event time per call in ms (on the communicator's stream)."""
        ms = ctypes.c_float(0)
        _check(self.lib.fbr_comm_allreduce_timed(self.handle, _ptr(buf), int(count), int(dtype), int(op), int(iters), ctypes.byref(ms)))
        return float(ms.value)

    def allreduce_i64(self, value):
        v = ctypes.c_int64(int(value))
        _check(self.lib.fbr_comm_allreduce_i64(self.handle, ctypes.byref(v)))
        return int(v.value)

    def allreduce_i64_begin(self, value):
        """Enqueue the scalar fold and return: it overlaps whatever the caller does next (the next map)."""
        _check(self.lib.fbr_comm_allreduce_i64_begin(self.handle, int(value)))

    def allreduce_i64_end(self):
        v = ctypes.c_int64(0)
        _check(self.lib.fbr_comm_allreduce_i64_end(self.handle, ctypes.byref(v)))
        return int(v.value)

    def destroy(self):
        h, self.handle = self.handle, None
        if h:
            self.lib.fbr_comm_destroy(h)

    def __del__(self):
        try:
            self.destroy()
        except Exception:       # noqa: BLE001
            pass


def allreduce_bench(comm, n_elements, steps=10, warmup=3):
    """BASELINE.json config 5 behind the C ABI: all-reduce (SUM) of an fp32 buffer across the communicator.
    Returns ``(ok, algbw_GBps, busbw_GBps, ms)``: values are small integers (rank + 1) so the fp32 sum is exact
    and the check is bit-exact; timing with CUDA events on the communicator's stream, max over ranks."""
    buf = comm.alloc(n_elements * 4).upload(np.full(n_elements, float(comm.rank + 1), dtype=np.float32))
    comm.allreduce(buf, buf, n_elements, F32, SUM)
    comm.sync()
    want = float(comm.nranks * (comm.nranks + 1) // 2)
    got = buf.download(np.float32)
    ok = bool((got == want).all())
    if warmup:
        comm.allreduce_timed(buf, n_elements, F32, SUM, warmup)
    ms = comm.allreduce_timed(buf, n_elements, F32, SUM, steps)
    # the job's time is the slowest rank's: max over ranks through the communicator itself (ms in ns as int64)
    ms = -comm_min_i64(comm, -int(ms * 1e6)) / 1e6
    buf.free()
    nbytes = n_elements * 4
    algbw = nbytes / (ms * 1e-3) / 1e9
    return ok, algbw, algbw * 2 * (comm.nranks - 1) / comm.nranks, ms


def comm_min_i64(comm, value):
    """Global minimum of one int64 per rank (used for max-over-ranks timing)."""
    b = comm.alloc(8).upload(np.array([value], dtype=np.int64))
    comm.allreduce(b, b, 1, I64, MIN)
    comm.sync()
    out = int(b.download(np.int64)[0])
    b.free()
    return out
