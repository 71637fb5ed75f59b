"""Ring node function for tests/test_ring.py (must be importable by spawned processes)."""
import json
import os


def allreduce_node(rank, size):
    import torch
    import torch.distributed as dist
    from fiber_b200.experimental import allreduce_bench
    assert dist.get_rank() == rank and dist.get_world_size() == size
    # the reference demo's message shapes (examples/ring.py:89-96): one all-reduce per parameter
    shapes = [500, 20, 25000, 50, 400000, 500, 5000, 10]
    ok = True
    dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    for n in shapes:
        g = torch.full((n,), float(rank + 1), device=dev)
        dist.all_reduce(g, op=dist.ReduceOp.SUM)
        ok &= bool((g == size * (size + 1) / 2).all().item())
    ok2, algbw, busbw, ms = allreduce_bench(int(os.environ.get("FBR_RING_ELEMS", "262144")), steps=2, warmup=1)
    # N(0,1) payload, rtol 1e-5 against an fp64 sum of the same per-rank streams
    gens = [torch.Generator().manual_seed(1234 + r) for r in range(size)]
    parts = [torch.randn(4096, generator=g_, dtype=torch.float32) for g_ in gens]
    mine = parts[rank].clone().to(dev)
    dist.all_reduce(mine, op=dist.ReduceOp.SUM)
    ref = sum(p.double() for p in parts)
    ok3 = bool(torch.allclose(mine.cpu().double(), ref, rtol=1e-5, atol=1e-6))
    out = os.environ.get("FBR_RING_OUT")
    if out:
        with open("%s.%d" % (out, rank), "w") as fh:
            json.dump({"rank": rank, "ok": ok and ok2 and ok3, "oks": [ok, ok2, ok3], "busbw": busbw, "backend": dist.get_backend()}, fh)
    dist.destroy_process_group()


def engine_allreduce_node(rank, size):
    """The same node function on the engine's own communicator (fbr_comm_*): no torch on the path."""
    import numpy as np
    from fiber_b200 import comm as C
    from fiber_b200.experimental import ring_comm
    c = ring_comm()
    assert (c.rank, c.nranks) == (rank, size)
    shapes = [500, 20, 25000, 50, 400000, 500, 5000, 10]         # examples/ring.py:89-96: one all-reduce per parameter
    ok = True
    for n in shapes:
        b = c.alloc(n * 4).upload(np.full(n, float(rank + 1), dtype=np.float32))
        c.allreduce(b, b, n, C.F32, C.SUM)
        c.sync()
        ok &= bool((b.download(np.float32) == size * (size + 1) / 2).all())
        b.free()
    ok2, algbw, busbw, ms = C.allreduce_bench(c, int(os.environ.get("FBR_RING_ELEMS", "262144")), steps=2, warmup=1)
    # N(0,1) payload, rtol 1e-5 against an fp64 sum of the same per-rank streams
    parts = [np.random.default_rng(1234 + r).standard_normal(4096).astype(np.float32) for r in range(size)]
    b = c.alloc(4096 * 4).upload(parts[rank])
    c.allreduce(b, b, 4096, C.F32, C.SUM)
    c.sync()
    ref = sum(p.astype(np.float64) for p in parts)
    ok3 = bool(np.allclose(b.download(np.float32).astype(np.float64), ref, rtol=1e-5, atol=1e-6))
    # the other collectives of the one-process-per-GPU mode: broadcast, allgather, gather/scatter, scalar fold
    blk = 1 << 16
    src = np.arange(blk, dtype=np.uint8) if rank == 0 else np.zeros(blk, dtype=np.uint8)
    b0 = c.alloc(blk).upload(src)
    c.broadcast(b0, blk, root=0)
    mine = c.alloc(blk).upload(np.full(blk, rank + 1, dtype=np.uint8))
    allb = c.alloc(blk * size)
    c.allgather(mine, allb, blk)
    root_buf = c.alloc(blk * size) if rank == 0 else None
    c.gather(mine, root_buf, blk, root=0)
    back = c.alloc(blk)
    c.scatter(allb if rank == 0 else None, back, blk, root=0)
    c.sync()
    want_all = np.repeat(np.arange(1, size + 1, dtype=np.uint8), blk)
    ok4 = bool((b0.download() == np.arange(blk, dtype=np.uint8)).all()) and bool((allb.download() == want_all).all()) \
        and bool((back.download() == rank + 1).all()) and (rank != 0 or bool((root_buf.download() == want_all).all())) \
        and c.allreduce_i64(rank + 1) == size * (size + 1) // 2
    out = os.environ.get("FBR_RING_OUT")
    if out:
        with open("%s.%d" % (out, rank), "w") as fh:
            json.dump({"rank": rank, "ok": ok and ok2 and ok3 and ok4, "oks": [ok, ok2, ok3, ok4], "busbw": busbw, "backend": "fbr_comm/nccl %d" % C.load_nccl()}, fh)
    c.destroy()
