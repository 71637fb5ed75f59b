"""``fiber_b200.Process`` -- the ``fiber.Process`` surface (fiber/process.py:83-323: ``start``,
``join``, ``terminate``, ``is_alive``, ``exitcode``, ``pid``, ``name``, ``daemon``) for GPU-resident
device processes.

The reference's ``Process`` runs an arbitrary Python callable in a job-backed OS process
(``Popen._launch`` -> ``backend.create_job``, fiber/popen_fiber_spawn.py:356-512).  Here a process is
a resident one-warp kernel on one GPU (``fbr_process_start``) that executes a compiled-in *process
body* against queue / pipe lanes; the Python ``target`` only selects the body
(``@device_process(name)``), exactly as ``@device_body`` does for ``Pool``.  Unbound targets raise
``TypeError``: nothing here runs a target on the CPU.
"""
import ctypes
import itertools
import threading

from . import _abi
from .backend import get_backend
from .core import DeviceCommand, JobSpec, ProcessStatus
from .queues import Connection, SimpleQueuePush, encode

__all__ = ["Process", "device_process", "active_children"]

_BODIES = {
    "queue_worker": _abi.FBR_PROC_QUEUE_WORKER,   # worker(q_in, q_out, ident)   tests/test_queue.py:44-50
    "put_queue": _abi.FBR_PROC_PUT_QUEUE,         # put_queue(q, data)           tests/test_queue.py:23-33
    "get_queue": _abi.FBR_PROC_GET_QUEUE,         # get_queue(q_in, q_out, n)    tests/test_queue.py:36-42
    "write_pipe": _abi.FBR_PROC_WRITE_PIPE,       # write_pipe(pipe, msg)        tests/test_queue.py:19-20
    "pipe_worker": _abi.FBR_PROC_PIPE_WORKER,     # pipe_worker(conn)            tests/test_queue.py:53-57
}
_counter = itertools.count(1)
_children = []
_children_lock = threading.Lock()


def device_process(name, **meta):
    """Bind a Python function to the device process body ``name`` (sets ``__fbr_process__`` and
    ``__fiber_meta__``, the attribute ``Popen`` reads for resource hints,
    fiber/popen_fiber_spawn.py:265-273)."""
    if name not in _BODIES:
        raise KeyError("no device process body named %r (have: %s)" % (name, ", ".join(sorted(_BODIES))))
    md = {"gpu": 1}
    md.update(meta)

    def decorator(func):
        func.__fbr_process__ = name
        func.__fiber_meta__ = md
        return func
    return decorator


def active_children():
    """fiber/process.py active_children(): live device processes started from this host process."""
    with _children_lock:
        _children[:] = [p for p in _children if p.is_alive()]
        return list(_children)


def _reader_lane(obj):
    if isinstance(obj, SimpleQueuePush):
        return obj._q.open_reader(), obj._q
    if isinstance(obj, Connection):
        if obj._rq is None:
            raise OSError("connection is write-only")
        return obj._rq.open_reader(), obj._rq
    raise TypeError("expected a SimpleQueue or Connection, got %r" % type(obj).__name__)


def _writer_lane(obj):
    if isinstance(obj, SimpleQueuePush):
        return obj._q.open_writer(), obj._q
    if isinstance(obj, Connection):
        if obj._sq is None:
            raise OSError("connection is read-only")
        return obj._sq.open_writer(), obj._sq
    raise TypeError("expected a SimpleQueue or Connection, got %r" % type(obj).__name__)


class Process:
    def __init__(self, group=None, target=None, name=None, args=(), kwargs={}, *, daemon=None, device=0,
                 idle_timeout=30.0):
        assert group is None, "group argument must be None for now"
        self._target, self._args, self._kwargs = target, tuple(args), dict(kwargs)
        self._idx = next(_counter)
        self.name = name or "Process-%d" % self._idx
        self.daemon = bool(daemon)
        self._device = device
        self._idle_timeout = idle_timeout
        self._handle = None
        self._keep = []            # queues whose lanes the device process uses
        self._exitcode = None
        self._backend = self._job = None

    def __repr__(self):
        status = "initial" if self._handle is None else ("started" if self.is_alive() else "stopped[%s]" % self.exitcode)
        return "<%s(%s, %s%s)>" % (type(self).__name__, self.name, status, ", daemon" if self.daemon else "")

    @property
    def pid(self):
        """Job-derived id in the reference (fiber/popen_fiber_spawn.py:153-156); here the launch index."""
        return None if self._handle is None else self._idx

    ident = pid

    def start(self):
        assert self._handle is None, "cannot start a process twice"
        body = getattr(self._target, "__fbr_process__", None)
        if body is None:
            raise TypeError("fiber_b200.Process: target %r is not bound to a device process body "
                            "(@fiber_b200.device_process(name); available: %s). There is no CPU fallback."
                            % (self._target, ", ".join(sorted(_BODIES))))
        if self._kwargs:
            raise TypeError("device process targets take positional arguments only")
        a = self._args
        lane_in = lane_out = None
        ident, msg, lst = 0, None, None
        if body == "queue_worker":
            q_in, q_out, ident = a
            (lane_in, k1), (lane_out, k2) = _reader_lane(q_in), _writer_lane(q_out)
            self._keep += [k1, k2]
        elif body == "put_queue":
            q, data = a
            lane_out, k = _writer_lane(q)
            self._keep.append(k)
            if type(data) is list:
                lst = (_abi.Record * max(1, len(data)))(*[encode(d) for d in data])
                ident = len(data)
            else:
                msg = encode(data)
        elif body == "get_queue":
            q_in, q_out, ident = a
            (lane_in, k1), (lane_out, k2) = _reader_lane(q_in), _writer_lane(q_out)
            self._keep += [k1, k2]
        elif body == "write_pipe":
            conn, m = a
            lane_out, k = _writer_lane(conn)
            self._keep.append(k)
            msg = encode(m)
        elif body == "pipe_worker":
            (conn,) = a
            (lane_in, k1), (lane_out, k2) = _reader_lane(conn), _writer_lane(conn)
            self._keep += [k1, k2]
        # Process -> JobSpec -> backend.create_job, the reference's layering
        # (fiber/process.py:187-215 -> popen_fiber_spawn.py:257-284 -> backend.create_job)
        cmd = DeviceCommand(body=_BODIES[body], lane_in=lane_in, lane_out=lane_out, ident=int(ident), msg=msg,
                            records=lst, idle_timeout=self._idle_timeout, keepalive=self._keep)
        meta = getattr(self._target, "__fiber_meta__", {}) or {}
        spec = JobSpec(command=cmd, name=self.name, cpu=meta.get("cpu"), mem=meta.get("mem"), gpu=self._device)
        self._backend = get_backend()
        self._job = self._backend.create_job(spec)
        self._handle = self._job.data
        with _children_lock:
            _children.append(self)

    def _poll(self):
        alive = self._backend.get_job_status(self._job) == ProcessStatus.STARTED
        if not alive:
            self._exitcode = self._job.exitcode
        return alive

    def is_alive(self):
        if self._handle is None or self._exitcode is not None:
            return False
        return self._poll()

    @property
    def exitcode(self):
        if self._handle is not None and self._exitcode is None:
            self._poll()
        return self._exitcode

    def join(self, timeout=None):
        assert self._handle is not None, "can only join a started process"
        if self._exitcode is not None:
            return
        code = self._backend.wait_for_job(self._job, timeout)
        if code is not None:
            self._exitcode = code

    def terminate(self):
        if self._handle is not None:
            self._backend.terminate_job(self._job)

    def handled(self):
        return self._backend.handled(self._job)

    def __del__(self):
        h, self._handle = getattr(self, "_handle", None), None
        if h and getattr(self, "_backend", None) is not None:
            self._backend.release_job(self._job)
