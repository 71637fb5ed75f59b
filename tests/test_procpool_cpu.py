"""CPU: host logic of the process-isolated resilient pool (fiber_b200/procpool.py) with a stand-in worker
(tests/_fake_worker.py): blocks are pulled by idle workers, results are placed by index in the shared segment, a
worker that dies has its block re-queued with attempt + 1 and is replaced (fiber/pool.py:1612-1659, 1009-1057)."""
import numpy as np
import pytest

from fiber_b200.procpool import BLOCK_ALIGN, ProcessPool, WorkerDied

from ._fake_worker import fake_worker_main


class _Spec:
    """Minimal body description the master needs (name, record sizes, flags, result dtype)."""

    def __init__(self, name):
        self.name, self.result_bytes, self.flags = name, 8, 0x4

    def result_dtype(self):
        return np.dtype(np.int64), ()

    def to_python(self, row):
        return row.item()

    def rows_to_list(self, arr):
        return arr.tolist()


def _pool(n, **kw):
    return ProcessPool(n, devices=list(range(n)), results="bytes", worker_main=fake_worker_main, **kw)


def test_blocks_are_pulled_and_placed_by_index():
    pool = _pool(3, block_tasks=BLOCK_ALIGN)
    try:
        n = 5 * BLOCK_ALIGN + 1234
        r = pool.submit(_Spec("identity_i64"), None, "map", range(n), 32).get(60)
        assert np.array_equal(np.asarray(r), np.arange(n)) and r.sum() == n * (n - 1) // 2
        xs = list(range(100, 0, -1))
        assert pool.submit(_Spec("square_i64"), None, "map", xs, 32).get(60) == [x * x for x in xs]
        assert pool.submit(_Spec("identity_i64"), None, "starmap", [(i,) for i in range(50)], 1).get(60) == list(range(50))
        assert len(pool.submit(_Spec("identity_i64"), None, "map", range(0), 32).get(5)) == 0
        assert pool.stats["blocks_dispatched"] >= 6 + 1 + 1 and pool.stats["workers_lost"] == 0
        with pytest.raises(OverflowError):
            pool.submit(_Spec("square_i64"), None, "map", [1, 2, 3037000500], 32).get(60)
        assert pool.submit(_Spec("square_i64"), None, "map", [3], 32).get(60) == [9]        # the pool keeps serving
        # a map that fails in its first block while its other blocks are still running: their late reports must not
        # leak into the next map's accounting
        bad = np.concatenate([[3037000500], np.arange(4 * BLOCK_ALIGN)]).astype(np.int64)
        h_bad = pool.submit(_Spec("square_i64"), None, "map", bad, 32)
        h_next = pool.submit(_Spec("identity_i64"), None, "map", range(3 * BLOCK_ALIGN + 5), 32)
        with pytest.raises(OverflowError):
            h_bad.get(60)
        r = h_next.get(60)
        m = 3 * BLOCK_ALIGN + 5
        assert np.array_equal(np.asarray(r), np.arange(m)) and r.sum() == m * (m - 1) // 2
    finally:
        pool.terminate()
        pool.join()


def test_dead_worker_block_is_requeued_and_worker_replaced():
    pool = _pool(2, block_tasks=BLOCK_ALIGN)
    try:
        n = 40 * BLOCK_ALIGN                      # arguments 0xDEAD + k * 2^20 kill their worker on the first attempt
        r = pool.submit(_Spec("trap_identity_i64"), None, "map", range(n), 32).get(120)
        assert np.array_equal(np.asarray(r), np.arange(n))
        st = pool.stats
        assert st["workers_lost"] == 2 and st["blocks_redispatched"] == 2 and st["workers_started"] == 4, st
        # a worker that reports its own death (sticky CUDA error) is handled the same way
        r = pool.submit(_Spec("fault_report_i64"), None, "map", range(100), 32).get(120)
        assert r == list(range(100)) and pool.stats["workers_lost"] == 3
        assert pool.submit(_Spec("identity_i64"), None, "map", range(10), 32).get(60) == list(range(10))
    finally:
        pool.terminate()
        pool.join()


def test_without_error_handling_a_death_fails_the_map_but_not_the_pool():
    pool = _pool(2, block_tasks=BLOCK_ALIGN, redispatch=False)
    try:
        with pytest.raises(WorkerDied, match="without error_handling"):
            pool.submit(_Spec("trap_identity_i64"), None, "map", range(4 * BLOCK_ALIGN), 32).get(120)
        assert pool.submit(_Spec("identity_i64"), None, "map", range(10), 32).get(60) == list(range(10))
        assert pool.stats["workers_lost"] == 1
    finally:
        pool.terminate()
        pool.join()
