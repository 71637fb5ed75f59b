"""World-size-2 worker for tests/test_shard_gloo.py: exercises the N>1 host path of bench.py
(rank blocks, barrier + max-over-ranks timing, count all-reduce, block all-gather) over gloo on CPU.
The per-rank "device step" is stood in for by the C oracle -- this test is about the sharding and
exchange logic, not the kernels."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as td  # noqa: E402

import bench  # noqa: E402
from fiber_b200 import shard  # noqa: E402
from oracle import cref  # noqa: E402


def main():
    n_total = int(sys.argv[1])
    dist = bench.Dist(int(os.environ["WORLD_SIZE"]), backend="gloo")
    lo, hi = shard.block_of(n_total, dist.rank, dist.world, align=4096)
    local, count = cref.pi_inside_range(lo, hi - lo)

    def step():
        cref.pi_inside_range(lo, min(hi - lo, 1000), want_array=False)

    t = bench.timed_steps(dist, 3, 1, step)
    total = dist.sum_i64(count)
    assert total == shard.all_reduce_sum_i64(td, count, "cpu")
    full = shard.all_gather_blocks(td, torch.from_numpy(local), n_total, dist.world, align=4096)
    ref, ref_count = cref.pi_inside_range(0, n_total)
    ok = bool((full.numpy() == ref).all()) and total == ref_count
    tmax = dist.max(float(dist.rank + 1))
    if dist.rank == 0:
        print(json.dumps({"ok": ok, "total": total, "ref": ref_count, "t": t, "tmax": tmax, "block": [lo, hi]}), flush=True)
    dist.finish()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
