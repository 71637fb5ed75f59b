"""``fiber_b200.experimental.Ring`` -- the reference's ring bootstrap helper
(fiber/experimental/ring.py:44-129) for one 8xB200 box, with NCCL over NVLink as the transport.

Same shape as the reference: ``Ring(processes, func, initializer).run()`` starts ``processes`` ring
nodes, each of which publishes itself in ``ring.members`` (``RingNode``: rank, connected, ip, port),
runs ``initializer(ring)`` and then ``func(rank, size)``.  Differences, by construction:

* a node is one process bound to one GPU (``LOCAL_RANK``), started with the ``spawn`` context
  (the reference starts rank 0 that way too, ring.py:113-116) instead of job-backed
  ``fiber.Process``es; under ``torchrun`` (``WORLD_SIZE`` already set) ``run()`` executes the local
  rank in place;
* the rendezvous a node publishes is the ``MASTER_ADDR:MASTER_PORT`` NCCL bootstraps from (the
  reference publishes ip/port for gloo, examples/ring.py:163-171);
* ``engine_ring_init`` is the B200 initializer: the member table carries the 128-byte NCCL bootstrap id
  (``RingNode.comm_id``, made by ``Ring.run`` before the nodes start) and every node builds its communicator
  through the C ABI (``fbr_comm_create`` = ``ncclCommInitRank``); the collective of the reference demo
  (``dist.all_reduce(param.grad.data, SUM)``, examples/ring.py:81-86) is then ``ring_comm().allreduce`` =
  ``ncclAllReduce`` over NVLink/NVSwitch, with no torch on the path;
* ``torch_ring_init`` keeps the reference's own route for hosts without a GPU (gloo; the CPU tests) and for
  code written against ``torch.distributed``: ``init_process_group("nccl")`` on GPUs.
"""
import multiprocessing as mp
import os
import socket
import time

__all__ = ["Ring", "RingNode", "torch_ring_init", "engine_ring_init", "ring_comm", "allreduce_bench"]

_comm = None          # this node's engine communicator (engine_ring_init)


def ring_comm():
    """The node's ``fiber_b200.comm.Comm`` once ``engine_ring_init`` has run."""
    if _comm is None:
        raise RuntimeError("this ring node has no engine communicator (initializer was not engine_ring_init)")
    return _comm


class RingNode:
    """A node in the ``Ring`` (fiber/experimental/ring.py:44-55)."""

    def __init__(self, rank):
        self.rank = rank
        self.connected = False
        self.ip = None
        self.port = None
        self.comm_id = None       # NCCL bootstrap id (128 bytes) published by rank 0: the B200 ring's "address"


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


class Ring:
    def __init__(self, processes, func, initializer, initargs=None):
        self.size = processes
        self.initializer = initializer
        self.initargs = initargs
        self.func = func
        self.rank = 0
        if getattr(func, "__fiber_meta__", None):
            self.__fiber_meta__ = func.__fiber_meta__        # ring.py:80-84
        self.members = [RingNode(i) for i in range(self.size)]
        self._master = ("127.0.0.1", None)

    def _target(self):
        rank = self.rank
        node = self.members[rank]
        node.connected = True
        node.ip, node.port = self._master
        self.members[0].connected = True                      # rank 0's rendezvous is fixed up front
        self.members[0].ip, self.members[0].port = self._master
        self.initializer(self)
        self.func(rank, self.size)

    def _child(self, rank):
        os.environ.update({"RANK": str(rank), "LOCAL_RANK": str(rank), "WORLD_SIZE": str(self.size),
                           "MASTER_ADDR": self._master[0], "MASTER_PORT": str(self._master[1])})
        self.rank = rank
        self._target()

    def run(self):
        """Start the ring (ring.py:103-129) and wait for every node to finish."""
        if self.size <= 0:
            return
        if int(os.environ.get("WORLD_SIZE", "0")) == self.size and "RANK" in os.environ:
            # already launched one-process-per-GPU (torchrun): run the local node in place
            self._master = (os.environ.get("MASTER_ADDR", "127.0.0.1"), int(os.environ.get("MASTER_PORT", "29500")))
            self.rank = int(os.environ["RANK"])
            self._target()
            return
        self._master = ("127.0.0.1", _free_port())
        if self.initializer is engine_ring_init:
            from .. import comm as _c
            self.members[0].comm_id = _c.unique_id()      # travels to every node with the pickled member table
        ctx = mp.get_context("spawn")
        procs = [ctx.Process(target=self._child, args=(i,)) for i in range(self.size)]
        for p in procs:
            p.start()
        for p in procs:
            p.join()
        bad = [p.exitcode for p in procs if p.exitcode != 0]
        if bad:
            raise RuntimeError("ring nodes failed with exit codes %s" % bad)


def torch_ring_init(ring):
    """Stock initializer (the role of ``pytorch_ring_init``, examples/ring.py:139-171): wait for the
    master's rendezvous, then join the process group -- NCCL when the node has a GPU."""
    import torch
    import torch.distributed as dist

    master = ring.members[0]
    wait = 0.1
    while master.connected is False:
        time.sleep(wait)
        wait *= 2
    os.environ["MASTER_ADDR"] = str(master.ip)
    os.environ["MASTER_PORT"] = str(master.port)
    if dist.is_initialized():
        return
    if torch.cuda.is_available():
        local = int(os.environ.get("LOCAL_RANK", ring.rank)) % torch.cuda.device_count()
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", rank=ring.rank, world_size=ring.size, device_id=torch.device("cuda", local))
    else:
        dist.init_process_group("gloo", rank=ring.rank, world_size=ring.size)


def engine_ring_init(ring):
    """B200 initializer: join the NCCL communicator whose bootstrap id rank 0 published in the member table
    (``fbr_comm_create`` -> ``ncclCommInitRank``), one node per GPU.  Under ``torchrun`` (no parent that could
    have filled the table) the id travels through a c10d TCPStore next to the rendezvous port."""
    global _comm
    from .. import comm as _c
    local = int(os.environ.get("LOCAL_RANK", ring.rank))
    cid = ring.members[0].comm_id
    if cid is None:
        import datetime
        import torch.distributed as dist
        store = dist.TCPStore(os.environ.get("MASTER_ADDR", "127.0.0.1"), int(os.environ.get("MASTER_PORT", "29500")) + 17,
                              ring.size, is_master=(ring.rank == 0), timeout=datetime.timedelta(seconds=120))
        if ring.rank == 0:
            store.set("fbr_ring_id", _c.unique_id())
        cid = bytes(store.get("fbr_ring_id"))
        ring.members[0].comm_id = cid
    ndev = _device_count()
    _comm = _c.Comm(local % max(1, ndev), ring.size, ring.rank, cid)
    ring.members[ring.rank].connected = True


def _device_count():
    import ctypes
    from .. import _abi
    n = ctypes.c_int(0)
    _abi.check(_abi.load().fbr_device_count(ctypes.byref(n)))
    return n.value


def allreduce_bench(n_elements, steps=10, warmup=3, device=None):
    """BASELINE.json config 5: all-reduce (SUM) of an fp32 buffer across the ring.  Returns
    ``(ok, algbw_GBps, busbw_GBps, ms)``: values are small integers (rank+1) so the fp32 sum is exact
    and the check is bit-exact; timing with CUDA events on the collective's stream, max over ranks."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(), dist.get_world_size()
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    buf = torch.full((n_elements,), float(rank + 1), dtype=torch.float32, device=device)
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    want = float(world * (world + 1) // 2)
    ok = bool((buf == want).all().item())
    for _ in range(warmup):
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    if device.type == "cuda":
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            dist.all_reduce(buf, op=dist.ReduceOp.SUM)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
    else:
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            dist.all_reduce(buf, op=dist.ReduceOp.SUM)
        ms = 1e3 * (time.perf_counter() - t0) / steps
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    nbytes = n_elements * 4
    algbw = nbytes / (ms * 1e-3) / 1e9
    return ok, algbw, algbw * 2 * (world - 1) / world, ms
