"""Short profiling target: two device-resident steps of each bench workload (pi 1e8, payload4k 1e6).

    ncu ... python profiles/prof_target.py [pi|payload|all]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from fiber_b200 import _abi  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
if which in ("pi", "all"):
    eng = bench.RawEngine(0, 160 << 20)
    out = eng.dalloc(bench.PI_TASKS)
    for _ in range(steps):
        print("pi", eng.wait(eng.submit("pi_inside_det", bench.PI_TASKS, out)))
    eng.dfree(out)
    eng.close()
if which in ("payload", "all"):
    n = bench.PAYLOAD_TASKS
    eng = bench.RawEngine(0, n * 4096 + (1 << 20))
    a, b = eng.dalloc(n * 4096), eng.dalloc(n * 4096)
    _abi.check(eng.lib.fbr_payload_fill_device(eng.h, 0, a, 0, n))
    for _ in range(steps):
        print("payload", eng.wait(eng.submit("payload_map_4k", n, b, args_dev=a, arg_stride=4096, want_sum=False)))
    for _ in range(steps):
        print("checksum", eng.wait(eng.submit("payload_checksum_4k", n, b, args_dev=a, arg_stride=4096)))
    eng.dfree(a)
    eng.dfree(b)
    eng.close()
