"""``SimpleQueue`` / ``Pipe`` of the reference (fiber/queues.py:262-352) on pinned, device-mapped
rings: every endpoint owns an SPSC lane of fixed-layout 64-byte records, a forwarder fair-queues
writers into readers with strict round-robin (``nn_device``, fiber/socket.py:297-320).  Endpoints
can be the host (these classes) or GPU-resident device processes (``fiber_b200.Process``).

Messages are what a GPU endpoint can read: ``None``, ``int`` (int64), ``float``, and ``bytes`` /
``str`` of at most 56 bytes.  Anything else raises ``TypeError`` (the reference pickles arbitrary
objects, queues.py:164-181; there is no CPU pickling path here).
"""
import ctypes
import struct

from . import _abi

__all__ = ["SimpleQueue", "SimpleQueuePush", "Pipe", "Connection"]


def encode(obj):
    r = _abi.Record()
    if obj is None:
        r.tag, r.len = _abi.FBR_REC_NONE, 0
        return r
    if isinstance(obj, bool) or not isinstance(obj, (int, float, bytes, bytearray, str)):
        raise TypeError("fiber_b200 queues carry None, int, float, bytes or str (<= 56 bytes); got %r" % type(obj).__name__)
    if isinstance(obj, int):
        if not -2 ** 63 <= obj <= 2 ** 63 - 1:
            raise OverflowError("queue message does not fit int64")
        r.tag, data = _abi.FBR_REC_INT, struct.pack("<q", obj)
    elif isinstance(obj, float):
        r.tag, data = _abi.FBR_REC_FLOAT, struct.pack("<d", obj)
    elif isinstance(obj, str):
        r.tag, data = _abi.FBR_REC_STR, obj.encode("utf-8")
    else:
        r.tag, data = _abi.FBR_REC_BYTES, bytes(obj)
    if len(data) > 56:
        raise ValueError("queue message payload is %d bytes; the fixed-layout record holds 56" % len(data))
    r.len = len(data)
    ctypes.memmove(r.payload, data, len(data))
    return r


def decode(r):
    data = bytes(r.payload[: r.len])
    if r.tag == _abi.FBR_REC_NONE:
        return None
    if r.tag == _abi.FBR_REC_INT:
        return struct.unpack("<q", data)[0]
    if r.tag == _abi.FBR_REC_FLOAT:
        return struct.unpack("<d", data)[0]
    if r.tag == _abi.FBR_REC_STR:
        return data.decode("utf-8")
    return data


def _ms(timeout):
    return -1 if timeout is None else max(0, int(timeout * 1000))


class _Queue:
    """One forwarder queue (the reference's ProcessDevice("r", "w"))."""

    def __init__(self):
        self.lib = _abi.load()
        h = ctypes.c_void_p()
        _abi.qcheck(self.lib.fbr_queue_create(ctypes.byref(h)))
        self.handle = h

    def open_writer(self):
        lane = ctypes.c_void_p()
        _abi.qcheck(self.lib.fbr_queue_open_writer(self.handle, ctypes.byref(lane)))
        return lane

    def open_reader(self):
        lane = ctypes.c_void_p()
        _abi.qcheck(self.lib.fbr_queue_open_reader(self.handle, ctypes.byref(lane)))
        return lane

    def stats(self):
        fwd, nw, nr = ctypes.c_uint64(), ctypes.c_uint32(), ctypes.c_uint32()
        _abi.qcheck(self.lib.fbr_queue_stats(self.handle, ctypes.byref(fwd), ctypes.byref(nw), ctypes.byref(nr)))
        return {"forwarded": fwd.value, "writers": nw.value, "readers": nr.value}

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            self.lib.fbr_queue_destroy(h)


class Connection:
    """Host endpoint with the ``ZConnection`` surface (fiber/queues.py:86-187): ``send`` / ``recv`` /
    ``poll``; its lanes are opened lazily on first use like ``LazyZConnection`` (queues.py:190-249)."""

    def __init__(self, recv_queue=None, send_queue=None):
        self._rq, self._sq = recv_queue, send_queue
        self._rlane = self._slane = None
        self._closed = False
        if recv_queue is None and send_queue is None:
            raise ValueError("at least one of `readable` and `writable` must be True")

    @property
    def readable(self):
        return self._rq is not None

    @property
    def writable(self):
        return self._sq is not None

    def _check(self):
        if self._closed:
            raise OSError("handle is closed")

    def send(self, obj):
        self._check()
        if self._sq is None:
            raise OSError("connection is read-only")
        if self._slane is None:
            self._slane = self._sq.open_writer()
        rec = encode(obj)
        _abi.qcheck(self._sq.lib.fbr_lane_send(self._slane, ctypes.byref(rec), -1))

    send_bytes = send

    def recv(self, timeout=None):
        self._check()
        if self._rq is None:
            raise OSError("connection is write-only")
        if self._rlane is None:
            self._rlane = self._rq.open_reader()
        rec = _abi.Record()
        rc = self._rq.lib.fbr_lane_recv(self._rlane, ctypes.byref(rec), _ms(timeout))
        if rc == _abi.FBR_ETIMEOUT:
            raise TimeoutError("no message")
        _abi.qcheck(rc)
        return decode(rec)

    recv_bytes = recv

    def poll(self, timeout=0.0):
        import time
        self._check()
        if self._rlane is None:
            self._rlane = self._rq.open_reader()
        deadline = None if timeout is None else time.monotonic() + timeout
        ready = ctypes.c_int(0)
        while True:
            _abi.qcheck(self._rq.lib.fbr_lane_poll(self._rlane, ctypes.byref(ready)))
            if ready.value or (deadline is not None and time.monotonic() >= deadline):
                return bool(ready.value)
            time.sleep(0.0005)

    def close(self):
        self._closed = True


def Pipe(duplex=True):
    """fiber/queues.py:262-281: a pair of connected connections; with ``duplex=False`` the first is
    read-only and the second write-only."""
    if duplex:
        a, b = _Queue(), _Queue()          # a: conn1 -> conn2, b: conn2 -> conn1
        return Connection(recv_queue=b, send_queue=a), Connection(recv_queue=a, send_queue=b)
    q = _Queue()
    return Connection(recv_queue=q), Connection(send_queue=q)


class SimpleQueuePush:
    """fiber/queues.py:284-352: ``put`` / ``get``; ``reader`` / ``writer`` are this process's own
    lazily connected endpoints.  Several device processes may read the same queue; messages are
    dealt to the connected readers round-robin."""

    def __init__(self):
        self._q = _Queue()
        self.reader = Connection(recv_queue=self._q)
        self.writer = Connection(send_queue=self._q)

    def __repr__(self):
        return "SimpleQueuePush<%s>" % (self._q.stats(),)

    def get(self, timeout=None):
        return self.reader.recv(timeout)

    def put(self, obj):
        self.writer.send(obj)

    def stats(self):
        return self._q.stats()


SimpleQueue = SimpleQueuePush
