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
* ``torch_ring_init`` is the stock initializer: ``init_process_group("nccl")`` on GPUs (gloo on
  CPU-only hosts, used by the CPU tests); the collective of the reference demo
  (``dist.all_reduce(param.grad.data, SUM)``, examples/ring.py:81-86) then runs as
  ``ncclAllReduce`` over NVLink/NVSwitch.
"""
import multiprocessing as mp
import os
import socket
import time

__all__ = ["Ring", "RingNode", "torch_ring_init", "allreduce_bench"]


class RingNode:
    """A node in the ``Ring`` (fiber/experimental/ring.py:44-55)."""

    def __init__(self, rank):
        self.rank = rank
        self.connected = False
        self.ip = None
        self.port = None


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
