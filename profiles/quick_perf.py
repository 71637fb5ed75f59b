"""Event-timed kernel times of the bench workloads (no ncu): quick A/B while tuning kernels."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from fiber_b200 import _abi  # noqa: E402

PEAK = bench.load_peaks()[0]["hbm_gbs"]


def report(tag, st):
    d_ms = st["dispatch_ms"] / max(1, st["dispatch_launches"])
    g_ms = st["gather_ms"] / max(1, st["gather_launches"])
    d_b = st["dispatch_bytes"] / max(1, st["dispatch_launches"])
    g_b = st["gather_bytes"] / max(1, st["gather_launches"])
    line = "%-10s dispatch %.4f ms (%.0f GB/s, %.1f%%)" % (tag, d_ms, d_b / d_ms / 1e6, 100 * d_b / d_ms / 1e6 / PEAK)
    if st["gather_launches"]:      # direct-placement maps launch no gather
        line += "   gather %.4f ms (%.0f GB/s, %.1f%%)" % (g_ms, g_b / g_ms / 1e6, 100 * g_b / g_ms / 1e6 / PEAK)
    print(line, flush=True)


steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
import time
for flags, tag in ((_abi.FBR_POOL_OVERLAP, "overlap"), (0, "serial")):
    e2 = bench.RawEngine(0, 320 << 20, flags)
    o2 = e2.dalloc(bench.PI_TASKS)
    for _ in range(3):
        e2.wait(e2.submit("pi_inside_det", bench.PI_TASKS, o2))
    t0 = time.perf_counter()
    seqs = [e2.submit("pi_inside_det", bench.PI_TASKS, o2) for _ in range(steps)]
    cnt = [e2.wait(q)[0] for q in seqs]
    dt = (time.perf_counter() - t0) / steps
    print("pi %-8s %.4f ms/step (wall, %d pipelined steps) count %d waves %d" % (tag, dt * 1e3, steps, cnt[-1], e2.stats()["dispatch_launches"] // (steps + 3)), flush=True)
    e2.dfree(o2); e2.close()
eng = bench.RawEngine(0, 160 << 20)
out = eng.dalloc(bench.PI_TASKS)
for i in range(steps + 3):
    if i == 3:
        eng.stats(reset=True)
    eng.wait(eng.submit("pi_inside_det", bench.PI_TASKS, out))
report("pi", eng.stats())
eng.dfree(out)
eng.close()
n = bench.PAYLOAD_TASKS
eng = bench.RawEngine(0, n * 4096 + (1 << 20))
a, b = eng.dalloc(n * 4096), eng.dalloc(n * 4096)
_abi.check(eng.lib.fbr_payload_fill_device(eng.h, 0, a, 0, n))
for body, sumflag in (("payload_map_4k", False), ("payload_checksum_4k", True)):
    for i in range(steps + 3):
        if i == 3:
            eng.stats(reset=True)
        eng.wait(eng.submit(body, n, b, args_dev=a, arg_stride=4096, want_sum=sumflag))
    report(body[:10], eng.stats())
eng.close()
