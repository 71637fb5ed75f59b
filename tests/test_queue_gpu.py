"""GPU: tests/test_queue.py of the reference restated with device processes (resident one-warp
kernels) as the workers: pipes, queues shared between the host and GPU processes, exact round-robin
balance, terminate / exitcode / watchdog."""
import collections
import time

import pytest

import fiber_b200

from . import workloads as W

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def check_leak():                                             # tests/test_queue.py:72-81
    assert fiber_b200.active_children() == []
    yield
    deadline = time.time() + 5
    while fiber_b200.active_children() and time.time() < deadline:
        time.sleep(0.01)
    assert fiber_b200.active_children() == []


def test_subprocess_with_pipe():                              # :100-106
    reader, writer = fiber_b200.Pipe()
    p = fiber_b200.Process(target=W.write_pipe, args=(writer, b"fiber pipe"))
    p.start()
    msg = reader.recv(10)
    p.join()
    assert msg == b"fiber pipe" and p.exitcode == 0 and not p.is_alive()


def test_pipe_duplex_over_fiber_process():                    # :122-129
    conn1, conn2 = fiber_b200.Pipe(duplex=True)
    p = fiber_b200.Process(target=W.pipe_worker, args=(conn2,))
    p.start()
    conn1.send(b"hello")
    data = conn1.recv(10)
    p.join()
    assert data == b"ack" and p.exitcode == 0


def test_simple_queue_fiber():                                # :150-156
    q = fiber_b200.SimpleQueue()
    p = fiber_b200.Process(target=W.put_queue, args=(q, 10))
    p.start()
    p.join()
    assert q.get(10) == 10


def test_simple_queue_fiber2():                               # :158-174
    q = fiber_b200.SimpleQueue()
    procs = [fiber_b200.Process(target=W.put_queue, args=(q, 10)) for _ in range(3)]
    for p in procs:
        p.start()
    for p in procs:
        p.join()
    assert [q.get(10) for _ in procs] == [10, 10, 10]


def test_simple_queue_fiber_multi():                          # :176-185
    n = 10
    q = fiber_b200.SimpleQueue()
    p = fiber_b200.Process(target=W.put_queue, args=(q, [i for i in range(n)]))
    p.start()
    p.join()
    assert [q.get(10) for _ in range(n)] == list(range(n))


def test_simple_queue_read_write_from_different_proc():       # :187-201
    n = 10
    q, q_out = fiber_b200.SimpleQueue(), fiber_b200.SimpleQueue()
    p1 = fiber_b200.Process(target=W.put_queue, args=(q, [i for i in range(n)]))
    p2 = fiber_b200.Process(target=W.get_queue, args=(q, q_out, n))
    p1.start()
    p2.start()
    assert [q_out.get(10) for _ in range(n)] == list(range(n))
    p1.join()
    p2.join()
    assert p1.exitcode == 0 and p2.exitcode == 0 and p2.handled() == n


def test_queue_balance():                                     # :218-250
    inqueue, outqueue = fiber_b200.SimpleQueue(), fiber_b200.SimpleQueue()
    num_workers, multiplier = 4, 600
    workers = [fiber_b200.Process(target=W.worker, args=(inqueue, outqueue, i), daemon=True) for i in range(num_workers)]
    for w in workers:
        w.start()
    assert len(fiber_b200.active_children()) == num_workers
    for _ in range(num_workers * multiplier):
        inqueue.put("work")
    results = [outqueue.get(20) for _ in range(num_workers * multiplier)]
    stats = collections.Counter(results)
    for _ in range(num_workers * multiplier):
        inqueue.put("quit")
    for w in workers:
        w.join()
    for i in range(num_workers):
        assert stats[i] == 600                                 # data is fairly queued
    assert all(w.exitcode == 0 for w in workers) and sum(w.handled() for w in workers) == 2400


def test_terminate_and_watchdog():                            # fiber/process.py terminate / exitcode
    q_in, q_out = fiber_b200.SimpleQueue(), fiber_b200.SimpleQueue()
    p = fiber_b200.Process(target=W.worker, args=(q_in, q_out, 7))
    p.start()
    q_in.put("work")
    assert q_out.get(10) == 7 and p.is_alive() and p.pid is not None
    p.terminate()
    p.join(10)
    assert not p.is_alive() and p.exitcode == -15              # SIGTERM-like
    idle = fiber_b200.Process(target=W.worker, args=(q_in, q_out, 8), idle_timeout=0.3)
    idle.start()
    idle.join(10)
    assert idle.exitcode == 3                                  # idle watchdog: never hangs the GPU


def test_device_process_alongside_pool_maps():
    """A resident device process must not block Pool maps (or be blocked by their allocations)."""
    q_in, q_out = fiber_b200.SimpleQueue(), fiber_b200.SimpleQueue()
    p = fiber_b200.Process(target=W.worker, args=(q_in, q_out, 1), idle_timeout=20)
    p.start()
    pool = fiber_b200.Pool(1)
    for _ in range(5):
        assert pool.map(W.f, range(1000)) == [i * i for i in range(1000)]
        q_in.put("work")
        assert q_out.get(10) == 1
    xs = __import__("oracle.bodies", fromlist=["x"]).parzen_example_inputs()
    assert len(pool.starmap(W.parzen_estimation, [(xs[0], xs[1], w) for w in xs[2][:5]], 1)) == 5
    q_in.put("quit")
    p.join(10)
    assert p.exitcode == 0
    pool.terminate()
    pool.join()
