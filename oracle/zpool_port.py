"""TEST INFRASTRUCTURE ONLY -- CPU port of the reference's ``ZPool`` / ``ResilientZPool`` algorithm.

This is the CPU arm ``bench.py`` times (``cpu_baseline.kind == "port"`` and ``--impl reference``)
and the behavioural oracle for pool semantics.  ``/root/reference`` cannot travel to the GPU box
(it is Python, needs ``nnpy``) so the algorithm is restated here, message for message:

=====================================  ==========================================================
reference (fiber/pool.py)              here
=====================================  ==========================================================
``ZPool.__init__`` :888-943            ``PortPool.__init__``: master PUSH + result PULL bound on
                                       random loopback TCP ports, task-handler thread
``_handle_tasks`` :952-963             ``_pump_tasks``: queue -> ``pickle.dumps`` -> send, with the
                                       same 20000-in-flight throttle (:904, :957-959)
``map_async`` :1139-1184               ``map_async``: default chunksize 32, ``list()`` of
                                       length-less iterables, one task tuple per chunk
                                       ``(seq, batch_start, func, chunk, False)``
``starmap_async`` :1258-1305           items wrapped ``(item,)``, 5th field ``True``
``apply_async`` :1089-1116             one chunk ``[(args, kwds)]``, 5th field ``True``
``zpool_worker_core`` :760-825         ``_worker_main``: recv chunk; per item ``func(arg)`` /
                                       ``func(*args)`` / ``func(*args, **kwds)``; ONE result message
                                       ``(seq, batch, batch+i, res)`` per item
``_res_get`` :968-973                  ``_recv_result``: recv + ``pickle.loads``
``Inventory`` :644-728                 ``Inventory``: placement by index, other seqs banked
``close/terminate/join`` :1337-1403    same state machine (RUN -> CLOSE | TERMINATE)
``ResilientZPool`` :1425-1688          ``ResilientPortPool``: REQ/REP pull dispatch, 8-byte
                                       ``struct "4si"`` hello, pending table keyed ``(seq, batch)``
                                       per worker ident, re-queue of a dead worker's chunks
=====================================  ==========================================================

Transport: pyzmq PUSH/PULL/REQ/REP over ``tcp://127.0.0.1`` -- the same socket patterns the
reference's default nanomsg context uses (fiber/socket.py:328-334).  Workers are real OS processes
(``multiprocessing`` *spawn* context: the reference also starts each worker as a fresh interpreter,
fiber/popen_fiber_spawn.py:233-249).  Validated against the real reference in
``tests/golden/make_golden.py``'s vectors by ``tests/test_oracle.py``.
"""
import math
import multiprocessing as mp
import os
import pickle
import queue
import secrets
import struct
import threading
import time

import zmq

RUN, CLOSE, TERMINATE = 0, 1, 2
DEFAULT_CHUNKSIZE = 32          # fiber/pool.py:1169-1170
MAX_PROCESSING_TASKS = 20000    # fiber/pool.py:904


class Inventory:
    """fiber/pool.py:644-728 -- per-``seq`` result arrays filled by index as messages arrive."""

    def __init__(self, recv):
        self._recv = recv
        self._next_seq = 0
        self._slots = {}
        self._left = {}
        self._cursor = {}

    def add(self, ntasks):
        self._next_seq += 1
        s = self._next_seq
        self._slots[s] = [None] * ntasks
        self._left[s] = ntasks
        self._cursor[s] = 0
        return s

    def _bank(self):
        seq, _batch, idx, value = self._recv()
        self._slots[seq][idx] = value
        return seq, idx

    def get(self, want):
        while self._left[want] != 0:
            seq, _ = self._bank()
            self._left[seq] -= 1
        out = self._slots[want]
        self._slots[want] = None
        return out

    def iget_unordered(self, want):
        while self._left[want] != 0:
            seq, idx = self._bank()
            self._left[seq] -= 1
            if seq == want:
                value = self._slots[want][idx]
                self._slots[want][idx] = None
                yield value

    def iget_ordered(self, want):
        # NB the reference treats a ``None`` result as "not arrived yet" (pool.py:702); restated.
        cur = self._cursor[want]
        total = len(self._slots[want])
        while cur != total:
            if self._slots[want][cur] is not None:
                value = self._slots[want][cur]
                self._slots[want][cur] = None
                cur += 1
                self._left[want] = cur
                yield value
                continue
            seq, idx = self._bank()
            if seq == want and idx == cur:
                value = self._slots[want][idx]
                self._slots[want][idx] = None
                cur += 1
                self._left[want] = cur
                yield value


class MapResult:
    def __init__(self, seq, inventory):
        self._seq, self._inv = seq, inventory

    def get(self):
        return self._inv.get(self._seq)

    def iget_ordered(self):
        return self._inv.iget_ordered(self._seq)

    def iget_unordered(self):
        return self._inv.iget_unordered(self._seq)


class ApplyResult(MapResult):
    def get(self):
        return self._inv.get(self._seq)[0]


def _run_chunk(task, send, ident=None):
    """fiber/pool.py:795-824: execute one chunk, one result message per item."""
    seq, batch, func, arg_list, starmap = task
    for i, item in enumerate(arg_list):
        if starmap:
            if len(item) == 2:
                args, kwds = item
                res = func(*args, **kwds)
            elif len(item) == 1:
                res = func(*item[0])
            else:
                raise ValueError("Bad number of args, %s %s", len(item), item)
        else:
            res = func(item)
        msg = (seq, batch, batch + i, res)
        if ident is not None:
            msg += (ident,)
        send(pickle.dumps(msg))


def _worker_main(master_addr, result_addr, initializer, initargs, req):
    """fiber/pool.py:832-878 (``zpool_worker``) + :760-825 (``zpool_worker_core``)."""
    if initializer is not None:
        initializer(*initargs)
    ctx = zmq.Context()
    master = ctx.socket(zmq.REQ if req else zmq.PULL)
    master.connect(master_addr)
    result = ctx.socket(zmq.PUSH)
    result.connect(result_addr)
    ident = secrets.token_bytes(4) if req else None
    try:
        while True:
            if req:
                master.send(struct.pack("4si", ident, os.getpid()))
            task = pickle.loads(master.recv())
            if task is None:
                break
            if len(task[3]) == 0:
                continue
            _run_chunk(task, result.send, ident)
    finally:
        result.close(linger=2000)
        master.close(linger=0)
        ctx.term()


class PortPool:
    """CPU port of ``fiber.pool.ZPool`` (push dispatch, no error handling)."""

    _req = False

    def __init__(self, processes=None, initializer=None, initargs=(), maxtasksperchild=None):
        self._processes = processes if processes is not None else 1   # pool.py:894
        self._initializer, self._initargs = initializer, initargs
        self._state = RUN
        self._taskq = queue.Queue()
        self.sent_tasks = 0
        self.recv_tasks = 0
        self._ctx = zmq.Context()
        self._master = self._ctx.socket(zmq.REP if self._req else zmq.PUSH)
        port = self._master.bind_to_random_port("tcp://127.0.0.1", min_port=40000, max_port=65535)
        self._master_addr = "tcp://127.0.0.1:%d" % port
        self._result = self._ctx.socket(zmq.PULL)
        port = self._result.bind_to_random_port("tcp://127.0.0.1", min_port=40000, max_port=65535)
        self._result_addr = "tcp://127.0.0.1:%d" % port
        self._inventory = Inventory(self._recv_result)
        self._workers = []
        self._workers_started = False
        self._pump = threading.Thread(target=self._pump_tasks, daemon=True)
        self._pump.start()

    # -- workers (lazy start, pool.py:1118-1137) -------------------------------------------------
    def _spawn_worker(self):
        p = mp.get_context("spawn").Process(
            target=_worker_main,
            args=(self._master_addr, self._result_addr, self._initializer, self._initargs, self._req),
            daemon=True)
        p.start()
        return p

    def start_workers(self):
        if not self._workers_started:
            self._workers_started = True
            self._workers = [self._spawn_worker() for _ in range(self._processes)]

    def wait_until_workers_up(self):
        self.start_workers()
        time.sleep(0.5)

    # -- hot loops ------------------------------------------------------------------------------
    def _pump_tasks(self):
        while True:
            if self.sent_tasks - self.recv_tasks > MAX_PROCESSING_TASKS:
                time.sleep(0.2)
                continue
            task = self._taskq.get()
            if task is StopIteration:
                return
            self._master.send(pickle.dumps(task))
            self.sent_tasks += 1

    def _recv_result(self):
        payload = self._result.recv()
        self.recv_tasks += 1
        return pickle.loads(payload)

    # -- API ------------------------------------------------------------------------------------
    def _check_running(self):
        if self._state != RUN:
            raise ValueError("Pool is not running")

    @staticmethod
    def _chunks(seq, size):
        for i in range(0, len(seq), size):
            yield seq[i:i + size]

    def _submit(self, func, items, chunksize, starmap):
        if chunksize is None:
            chunksize = DEFAULT_CHUNKSIZE
        self.start_workers()
        seq = self._inventory.add(len(items))
        for b, chunk in enumerate(self._chunks(items, chunksize)):
            self._taskq.put((seq, b * chunksize, func, chunk, starmap))
        return MapResult(seq, self._inventory)

    def map_async(self, func, iterable, chunksize=None, callback=None, error_callback=None):
        if error_callback:
            raise NotImplementedError
        self._check_running()
        if not hasattr(iterable, "__len__"):
            iterable = list(iterable)
        return self._submit(func, iterable, chunksize, False)

    def map(self, func, iterable, chunksize=None):
        return self.map_async(func, iterable, chunksize).get()

    def starmap_async(self, func, iterable, chunksize=None, callback=None, error_callback=None):
        self._check_running()
        if not hasattr(iterable, "__len__"):
            iterable = list(iterable)
        return self._submit(func, [(item,) for item in iterable], chunksize, True)

    def starmap(self, func, iterable, chunksize=None):
        return self.starmap_async(func, iterable, chunksize).get()

    def apply_async(self, func, args=(), kwds={}, callback=None, error_callback=None):
        self._check_running()
        self.start_workers()
        seq = self._inventory.add(1)
        self._taskq.put((seq, 0, func, [(args, kwds)], True))
        return ApplyResult(seq, self._inventory)

    def apply(self, func, args=(), kwds={}):
        return self.apply_async(func, args, kwds).get()

    def imap(self, func, iterable, chunksize=1):
        return self.map_async(func, iterable, chunksize).iget_ordered()

    def imap_unordered(self, func, iterable, chunksize=1):
        return self.map_async(func, iterable, chunksize).iget_unordered()

    # -- shutdown (pool.py:1332-1403) -------------------------------------------------------------
    def close(self):
        if self._state == RUN:
            self._state = CLOSE
            for _ in range(self._processes):
                self._taskq.put(None)

    def terminate(self):
        self._state = TERMINATE
        for p in self._workers:
            if p.is_alive():
                p.terminate()

    def join(self):
        assert self._state in (CLOSE, TERMINATE)
        for p in self._workers:
            p.join()
        self._taskq.put(StopIteration)
        self._pump.join(timeout=2)
        self._master.close(linger=0)
        self._result.close(linger=0)
        self._ctx.term()


class ResilientPortPool(PortPool):
    """CPU port of ``fiber.pool.ResilientZPool`` (pool.py:1425-1688): workers PULL work with a
    REQ hello; the master remembers which chunk each worker holds and re-queues the chunks of a
    worker that died."""

    _req = True

    def __init__(self, processes=None, initializer=None, initargs=(), maxtasksperchild=None):
        self._pending = {}        # ident -> {(seq, batch): task}
        self._pid_to_ident = {}
        self._lock = threading.Lock()
        super().__init__(processes, initializer, initargs, maxtasksperchild)
        self._watch = threading.Thread(target=self._watch_workers, daemon=True)

    def start_workers(self):
        first = not self._workers_started
        super().start_workers()
        if first:
            self._watch.start()

    def _pump_tasks(self):
        while True:
            task = self._taskq.get()
            if task is StopIteration:
                return
            ident, pid = struct.unpack("4si", self._master.recv())   # pool.py:1526-1527
            with self._lock:
                self._pending.setdefault(ident, {})
                self._pid_to_ident[pid] = ident
                if task is not None:
                    self._pending[ident][(task[0], task[1])] = task  # pool.py:1540
            self._master.send(pickle.dumps(task))
            self.sent_tasks += 1

    def _recv_result(self):
        seq, batch, idx, value, ident = pickle.loads(self._result.recv())
        self.recv_tasks += 1
        with self._lock:
            task = self._pending.get(ident, {}).get((seq, batch))
            if task is not None and idx == batch + len(task[3]) - 1:   # pool.py:1499-1505
                del self._pending[ident][(seq, batch)]
        return seq, batch, idx, value

    def _watch_workers(self):
        """pool.py:1612-1659: poll every 0.5 s, restart dead workers, re-queue their chunks."""
        while self._state == RUN:
            for i, p in enumerate(self._workers):
                if self._state != RUN:
                    break
                if p.exitcode is not None:
                    p.join()
                    self._workers[i] = self._spawn_worker()
                    with self._lock:
                        ident = self._pid_to_ident.pop(p.pid, None)
                        lost = self._pending.pop(ident, {}) if ident is not None else {}
                    for task in lost.values():
                        self._taskq.put(task)
            time.sleep(0.5)


def chunk_plan(n, chunksize=None):
    """``[(batch_start, count), ...]`` exactly as ``map_async`` cuts ``n`` items
    (fiber/pool.py:1084-1087, 1169-1181).  Host-logic oracle for the engine's chunk descriptors."""
    if chunksize is None:
        chunksize = DEFAULT_CHUNKSIZE
    return [(s, min(chunksize, n - s)) for s in range(0, n, chunksize)]


def n_jobs(processes, cpu_per_job=1):
    """Number of job-backed worker processes (fiber/pool.py:1405-1408)."""
    return math.ceil(float(processes) / cpu_per_job)
