"""CPU: host-side behaviour of SimpleQueue / Pipe (fiber/queues.py:262-352) -- the forwarder's
fair-queue-in / round-robin-out contract, message encoding, and the no-GPU failure mode of device
processes.  Restates the parts of tests/test_queue.py that need no worker process."""
import collections
import ctypes
import threading

import pytest

import fiber_b200
from fiber_b200 import _abi
from fiber_b200.queues import Connection, decode, encode

from . import workloads as W


def test_pipe():                                             # tests/test_queue.py:83-88
    reader, writer = fiber_b200.Pipe()
    writer.send(b"hello")
    assert reader.recv(5) == b"hello"


def test_pipe_duplex():                                      # tests/test_queue.py:108-120
    conn1, conn2 = fiber_b200.Pipe(duplex=True)
    conn1.send(b"hello")
    assert conn2.recv(5) == b"hello"
    conn2.send(b"hi")
    assert conn1.recv(5) == b"hi"
    assert conn1.poll(0.01) is False


def test_pipe_simplex_directions():                          # fiber/queues.py:276-281
    reader, writer = fiber_b200.Pipe(duplex=False)
    assert reader.readable and not reader.writable and writer.writable and not writer.readable
    writer.send(1)
    assert reader.recv(5) == 1
    with pytest.raises(OSError):
        reader.send(1)
    with pytest.raises(OSError):
        writer.recv(0.01)
    with pytest.raises(ValueError):
        Connection()


def test_simple_queue_fifo_and_types():                      # tests/test_queue.py:141-176
    q = fiber_b200.SimpleQueue()
    items = [10, -3, 2 ** 62, 0.5, "work", "quit", b"fiber pipe", None, "x" * 56]
    for it in items:
        q.put(it)
    assert [q.get(5) for _ in items] == items
    for i in range(10):
        q.put(i)
    assert [q.get(5) for _ in range(10)] == list(range(10))
    with pytest.raises(TimeoutError):
        q.get(0.01)
    for bad in ([1, 2], {"a": 1}, object(), True, "x" * 57, 2 ** 63):
        with pytest.raises((TypeError, ValueError, OverflowError)):
            q.put(bad)
    for v in (None, 7, -1.25, b"\x00\xff", "héllo"):
        assert decode(encode(v)) == v
    assert ctypes.sizeof(_abi.Record) == 64


def test_queue_balance_host_readers():
    """Round-robin out: 2400 messages over 4 connected readers => exactly 600 each
    (tests/test_queue.py:218-250).  Readers here are host endpoints on threads."""
    q, out = fiber_b200.SimpleQueue(), fiber_b200.SimpleQueue()
    n_workers, mult = 4, 600
    readers = [Connection(recv_queue=q._q) for _ in range(n_workers)]
    for r in readers:
        r.poll(0)                      # connect (opens the lane) before anything is put

    def work(conn, ident):
        w = Connection(send_queue=out._q)
        while True:
            if conn.recv(10) == "quit":
                break
            w.send(ident)
    threads = [threading.Thread(target=work, args=(readers[i], i)) for i in range(n_workers)]
    for t in threads:
        t.start()
    for _ in range(n_workers * mult):
        q.put("work")
    stats = collections.Counter(out.get(10) for _ in range(n_workers * mult))
    for _ in range(n_workers):
        q.put("quit")
    for t in threads:
        t.join(10)
    assert [stats[i] for i in range(n_workers)] == [mult] * n_workers
    assert q.stats()["readers"] == n_workers and q.stats()["forwarded"] == n_workers * mult + n_workers


def test_fair_queue_in_from_many_writers():
    q = fiber_b200.SimpleQueue()
    writers = [Connection(send_queue=q._q) for _ in range(3)]
    for k in range(50):
        for i, w in enumerate(writers):
            w.send(i * 1000 + k)
    got = [q.get(5) for _ in range(150)]
    for i in range(3):                                        # per-writer FIFO order is preserved
        assert [g % 1000 for g in got if g // 1000 == i] == list(range(50))


def _no_gpu():
    n = ctypes.c_int(0)
    return not (_abi.load().fbr_device_count(ctypes.byref(n)) == 0 and n.value > 0)


def test_process_binding_and_no_cpu_fallback():
    assert W.worker.__fbr_process__ == "queue_worker" and W.worker.__fiber_meta__ == {"gpu": 1}
    with pytest.raises(KeyError):
        fiber_b200.device_process("no_such_process_body")
    p = fiber_b200.Process(target=print, args=(1,))
    assert p.pid is None and not p.is_alive() and p.exitcode is None
    with pytest.raises(TypeError, match="no CPU fallback"):
        p.start()
    if _no_gpu():
        q = fiber_b200.SimpleQueue()
        p = fiber_b200.Process(target=W.put_queue, args=(q, 10))
        with pytest.raises(_abi.EngineError) as ei:
            p.start()
        assert ei.value.status == _abi.FBR_ENODEV
    assert fiber_b200.active_children() == []
