"""Host-side cost of one submission vs the device time of the step (is the submit thread the bottleneck?).

    python profiles/submit_perf.py [n_gpus]

For Pool(1) and the in-process Pool(n_gpus): wall time of fbr_map_submit alone, of submit + wait, per map of
1e8 index tasks per GPU (bit-packed bool results into the pinned segment)."""
import ctypes
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fiber_b200  # noqa: E402
from fiber_b200 import _abi, registry  # noqa: E402

ng = int(sys.argv[1]) if len(sys.argv) > 1 else fiber_b200.cpu_count()
for n_workers in sorted({1, ng}):
    pool = fiber_b200.Pool(n_workers)
    pool.start_workers()
    eng, lib = pool._engine, pool._engine.lib
    n_items = n_workers * 10 ** 8
    spec = registry.spec("pi_inside_bits8")
    d = _abi.MapDesc()
    d.func_id, d.flags, d.n_tasks, d.chunksize = spec.func_id, _abi.FBR_WANT_SUM, n_items // 8, 4
    d.index_start, d.index_step, d.n_items = 0, 1, n_items
    res = _abi.Result()
    t_sub, t_all = [], []
    for it in range(13):
        seq = ctypes.c_uint64()
        t0 = time.perf_counter()
        _abi.check(lib.fbr_map_submit(eng.handle, ctypes.byref(d), ctypes.byref(seq)))
        t1 = time.perf_counter()
        _abi.check(lib.fbr_result_wait(eng.handle, seq.value, -1, ctypes.byref(res)))
        t2 = time.perf_counter()
        _abi.check(lib.fbr_result_release(eng.handle, seq.value))
        if it >= 3:
            t_sub.append(t1 - t0)
            t_all.append(t2 - t0)
    print("Pool(%d): submit %.3f ms  submit+wait %.3f ms  (%d waves)  -> %.3e tasks/s" %
          (n_workers, 1e3 * sum(t_sub) / len(t_sub), 1e3 * sum(t_all) / len(t_all), res.n_waves, n_items / (sum(t_all) / len(t_all))), flush=True)
    pool.terminate()
    pool.join()
