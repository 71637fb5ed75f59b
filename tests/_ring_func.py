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
