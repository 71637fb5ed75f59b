"""Socket layer with the reference's plug-in shape (fiber/socket.py:66-82 ``SockContext``,
:379-413 ``Socket``, :416-425 ``ProcessDevice``) on pinned lanes instead of nanomsg sockets.

A bound socket owns a forwarder queue; its *address* is a token other sockets ``connect`` to.  Modes
keep their meaning (fiber/socket.py:328-334): ``w`` PUSH, ``r`` PULL, ``rw`` PAIR, ``req`` / ``rep``
request-reply (two lanes).  Messages are the fixed-layout records of ``fiber_b200.queues`` (``bytes``
of at most 56 bytes, ints, floats, short strings).  ``device(s1_mode, s2_mode)`` returns the forwarder
the reference builds for queues and pipes (``nn_device``): writers are fair-queued in, readers are
served round-robin."""
import itertools
import threading

from .queues import Connection, _Queue

_registry = {}                 # address token -> ((_Queue forward, _Queue backward), upstream): duplex sockets that connect
                               # to an `upstream` address send forward / receive backward, the others the reverse
_registry_lock = threading.Lock()
_tokens = itertools.count(1)
MODES = ("r", "w", "rw", "req", "rep")


def _publish(pair, upstream=False):
    with _registry_lock:
        addr = "lane://%d" % next(_tokens)
        _registry[addr] = (pair, upstream)
        return addr


def _lookup(addr):
    with _registry_lock:
        try:
            return _registry[addr]
        except KeyError:
            raise ConnectionError("no lane endpoint bound at %r" % (addr,)) from None


class SockContext:
    """fiber/socket.py:66-82: the factory a transport implements."""
    default_addr = None

    def new(self, mode):
        raise NotImplementedError

    @staticmethod
    def bind_random(sock, addr):
        raise NotImplementedError

    @staticmethod
    def connect(sock, addr):
        raise NotImplementedError

    @staticmethod
    def close(sock):
        sock.close()


class _LaneSocket:
    def __init__(self, mode):
        self.mode, self.conn, self.addr = mode, None, None

    def _attach(self, pair, binder):
        fwd, back = pair                       # fwd: binder's outgoing direction for "w"; see table below
        m = self.mode
        if m == "w":
            self.conn = Connection(send_queue=fwd)
        elif m == "r":
            self.conn = Connection(recv_queue=fwd)
        else:                                  # rw / req / rep: two directions
            self.conn = Connection(recv_queue=back, send_queue=fwd) if binder else Connection(recv_queue=fwd, send_queue=back)

    def send(self, data):
        self.conn.send(data)

    def recv(self, timeout=None):
        return self.conn.recv(timeout)

    def close(self):
        if self.conn is not None:
            self.conn.close()


class LaneContext(SockContext):
    default_addr = "lane://"

    def new(self, mode):
        return _LaneSocket(mode) if mode in MODES else None

    @staticmethod
    def bind_random(sock, addr):
        pair = (_Queue(), _Queue())
        sock.addr = _publish(pair)
        sock._attach(pair, binder=True)
        return sock.addr

    @staticmethod
    def connect(sock, addr):
        sock.addr = addr
        pair, upstream = _lookup(addr)
        # a duplex socket on the inbound side of a device talks in the same direction a binder would
        sock._attach(pair, binder=upstream)

    def device(self, s1_mode, s2_mode):
        """Forwarder between an inbound and an outbound endpoint.  Returns ``(device, in_addr,
        out_addr)`` like fiber/socket.py:352-366: writers connect to ``in_addr``, readers to
        ``out_addr``; for duplex modes the second direction flows the other way."""
        q_fwd, q_back = _Queue(), _Queue()
        # in_addr side: "w"/"rw" sockets send into q_fwd, "rw" sockets read q_back;
        # out_addr side: "r"/"rw" sockets read q_fwd, "rw" sockets send into q_back (the duplex Pipe of
        # fiber/queues.py:272 is ProcessDevice("rw", "rw") with one socket on each address)
        in_addr = _publish((q_fwd, q_back), upstream=True)
        out_addr = _publish((q_fwd, q_back), upstream=False)
        return _Device(), in_addr, out_addr


class _Device:
    """The forwarder itself is the engine's hub thread (queues.cu); nothing to start."""

    def start(self):
        return None


default_socket_ctx = LaneContext()


def get_ctx():
    return default_socket_ctx


class Socket:
    """fiber/socket.py:379-413."""

    def __init__(self, ctx=None, mode="rw"):
        self._mode = mode
        self._ctx = ctx or get_ctx()
        self._sock = self._ctx.new(mode)
        if self._sock is None:
            raise ValueError('Socket mode "{}" not supported by {}'.format(mode, self._ctx.__class__.__name__))

    def __repr__(self):
        return "{}<{},{}>".format(self.__class__.__name__, self._ctx.__class__.__name__, self._mode)

    def send(self, data):
        self._sock.send(data)

    def recv(self, timeout=None):
        return self._sock.recv(timeout)

    def bind(self):
        return self._ctx.bind_random(self._sock, self._ctx.default_addr)

    def connect(self, addr):
        self._ctx.connect(self._sock, addr)

    def close(self):
        self._ctx.close(self._sock)


class ProcessDevice:
    """fiber/socket.py:416-425."""

    def __init__(self, s1_mode, s2_mode, ctx=None):
        ctx = ctx or get_ctx()
        self.device, self.in_addr, self.out_addr = ctx.device(s1_mode, s2_mode)

    def start(self):
        self.device.start()
