"""Sharding of one map over ranks (one process per GPU, launched by torchrun).

The path shards naturally: tasks are independent, the reference already hands chunks to workers
with no inter-worker communication (fiber/pool.py:1179-1181, 819-824).  Rank ``g`` of ``G`` owns the
contiguous task block ``block_of(n, g, G)`` (== PUSH round-robin with chunk = block);
no data-path collective is needed.  The only exchange steps are

* a scalar all-reduce when the caller wants a global fold (the pi count), and
* an all-gather when the ordered results of every block must end up together,

both issued by the engine's own communicator on GPUs (``fiber_b200.comm.Comm``: ``fbr_comm_allreduce_i64`` /
``fbr_comm_allgather``, NCCL behind the C ABI); the CPU tests pass a ``torch.distributed`` module instead
(gloo), which exercises the same block arithmetic.
"""


def block_of(n_tasks, rank, world, align=1):
    """Contiguous block ``[lo, hi)`` of rank ``rank``: block sizes are multiples of ``align``
    (claim-unit alignment) except the last, and differ by less than ``2 * align`` (the tail unit may be partial)."""
    if not (0 <= rank < world) or n_tasks < 0 or align < 1:
        raise ValueError("bad shard request")
    units = (n_tasks + align - 1) // align
    base, extra = divmod(units, world)
    lo_u = rank * base + min(rank, extra)
    hi_u = lo_u + base + (1 if rank < extra else 0)
    return min(n_tasks, lo_u * align), min(n_tasks, hi_u * align)


def blocks(n_tasks, world, align=1):
    return [block_of(n_tasks, r, world, align) for r in range(world)]


def _is_engine_comm(group):
    from .comm import Comm
    return isinstance(group, Comm)


def all_reduce_sum_i64(dist_module, value, device=None):
    """Global int64 sum of one scalar per rank (``ncclAllReduce`` on GPUs)."""
    if _is_engine_comm(dist_module):
        return dist_module.allreduce_i64(value)
    import torch
    t = torch.tensor([int(value)], dtype=torch.int64, device=device)
    dist_module.all_reduce(t, op=dist_module.ReduceOp.SUM)
    return int(t.item())


def all_gather_blocks(dist_module, local, n_total, world, align=1):
    """Concatenate every rank's ordered block (a 1-D torch tensor, block sizes from ``block_of``)
    into the full ordered result on every rank.  Blocks may differ in size by one unit, so the
    exchange pads to the largest block and trims.

    With an engine communicator, ``local`` is a device buffer / pointer holding this rank's block padded to the
    largest block, and the result is a ``DeviceBuffer`` of ``world * width`` items laid out block after block
    (``gathered_block_offsets`` says where each block starts): nothing leaves the device."""
    if _is_engine_comm(dist_module):
        comm, item = dist_module, int(getattr(local, "itemsize", 1))
        sizes = [hi - lo for lo, hi in blocks(n_total, world, align)]
        width = max(sizes) if sizes else 0
        out = comm.alloc(max(1, width * world * item))
        comm.allgather(local, out, width * item)
        comm.sync()
        return out
    import torch
    sizes = [hi - lo for lo, hi in blocks(n_total, world, align)]
    width = max(sizes) if sizes else 0
    padded = torch.zeros(width, dtype=local.dtype, device=local.device)
    padded[: local.numel()] = local
    out = torch.empty(width * world, dtype=local.dtype, device=local.device)
    dist_module.all_gather_into_tensor(out, padded)
    return torch.cat([out[r * width: r * width + sizes[r]] for r in range(world)])


def gathered_block_offsets(n_total, world, align=1):
    """Item offset of every rank's block inside the buffer ``all_gather_blocks`` fills on an engine communicator
    (blocks are padded to the largest one)."""
    sizes = [hi - lo for lo, hi in blocks(n_total, world, align)]
    width = max(sizes) if sizes else 0
    return [(r * width, sizes[r]) for r in range(world)]
