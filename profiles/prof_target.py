"""Short profiling target: device-resident steps of each bench workload (pi 1e8, payload4k 1e6, parzen).

    ncu ... python profiles/prof_target.py [pi|payload|parzen|all] [steps]

pi / payload run `steps` direct-placement steps (the dispatch kernel stores at the final index) and then
`steps` steps through task records + ring + gather_ordered (shuffled arrival / FBR_VIA_RING), so one
capture holds every hot kernel.
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
        print("pi direct", eng.wait(eng.submit("pi_inside_det", bench.PI_TASKS, out)))
    for k in range(steps):
        print("pi shuffled", eng.wait(eng.submit("pi_inside_det", bench.PI_TASKS, out, extra_flags=_abi.FBR_SHUFFLE, seed=k + 1)))
    eng.dfree(out)
    eng.close()
if which in ("payload", "all"):
    n = bench.PAYLOAD_TASKS
    eng = bench.RawEngine(0, n * 4096 + (1 << 20))
    a, b = eng.dalloc(n * 4096), eng.dalloc(n * 4096)
    _abi.check(eng.lib.fbr_payload_fill_device(eng.h, 0, a, 0, n))
    for _ in range(steps):
        print("payload direct", eng.wait(eng.submit("payload_map_4k", n, b, args_dev=a, arg_stride=4096, want_sum=False)))
    for _ in range(steps):
        print("payload via ring", eng.wait(eng.submit("payload_map_4k", n, b, args_dev=a, arg_stride=4096, want_sum=False,
                                                      extra_flags=_abi.FBR_VIA_RING)))
    for _ in range(steps):
        print("checksum", eng.wait(eng.submit("payload_checksum_4k", n, b, args_dev=a, arg_stride=4096)))
    eng.dfree(a)
    eng.dfree(b)
    eng.close()
if which in ("parzen", "all"):
    import fiber_b200
    from examples import workloads as W
    from oracle import bodies as B
    xs, px, widths = B.parzen_example_inputs()
    pool = fiber_b200.Pool(1)
    for _ in range(steps + 1):
        r = pool.starmap(W.parzen_estimation_f32, [(xs, px, w) for w in widths], 1)
    print("parzen", r[0], r[-1])
    pool.terminate()
    pool.join()
