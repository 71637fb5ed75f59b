"""Wall-clock e2e of Pool.map(is_inside, range(1e8)) for each results= mode (quick A/B)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fiber_b200  # noqa: E402
from examples import workloads as W  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for mode in ("host", "bytes", "device"):      # host: bool results one bit each, stored zero-copy; bytes: one byte each
    pool = fiber_b200.Pool(1, results=mode)
    r = range(10 ** 8)
    for _ in range(3):
        c = pool.map(W.is_inside, r).sum()
    t0 = time.perf_counter()
    for _ in range(steps):
        c = pool.map(W.is_inside, r).sum()
    dt = (time.perf_counter() - t0) / steps
    print("results=%-6s %.4f ms/step  %.3e tasks/s  count %d  waves %d" % (mode, dt * 1e3, 1e8 / dt, c, pool.stats()["dispatch_launches"] // (steps + 3)), flush=True)
    pool.terminate()
    pool.join()
