"""Plug-in contract of the reference's job layer (fiber/core.py:21-113), kept so that a backend can
be swapped by name exactly as in fiber: a ``Backend`` creates, watches and stops *jobs*; ``JobSpec``
says what to run and with which resources; ``ProcessStatus`` is what ``get_job_status`` answers.
On this engine a job is a GPU-resident device process (``fiber_b200.gpu_backend``)."""
import enum
from dataclasses import dataclass, field
from typing import Any, Optional


class ProcessStatus(enum.Enum):
    UNKNOWN = 0
    INITIAL = 1
    STARTED = 2
    STOPPED = 3


@dataclass(eq=True)
class JobSpec:
    """What ``create_job`` receives.  ``command`` is the launch description of the job: the reference
    passes a python argv list (fiber/popen_fiber_spawn.py:233-249); the GPU backend expects a
    ``DeviceCommand`` (process body + lanes).  ``gpu`` selects the device."""
    image: Optional[str] = None
    command: Any = None
    name: Optional[str] = None
    cpu: Optional[int] = None
    mem: Optional[int] = None
    volumes: Optional[dict] = None
    gpu: Optional[int] = None

    def __repr__(self):
        return "<JobSpec: {}>".format(vars(self))


@dataclass
class Job:
    """Handle returned by ``create_job``: backend-private ``data`` plus the job id ``jid``."""
    data: Any = None
    jid: Any = None
    host: Optional[str] = None

    def update(self):
        raise NotImplementedError


class Backend:
    """The six calls every backend answers (fiber/core.py:79-113)."""

    @property
    def name(self):
        raise NotImplementedError

    def create_job(self, job_spec):
        raise NotImplementedError

    def get_job_status(self, job):
        raise NotImplementedError

    def get_job_logs(self, job):
        return ""

    def wait_for_job(self, job, timeout):
        """``None`` if still running after ``timeout`` seconds (``None`` = wait forever), else the exit code."""
        raise NotImplementedError

    def terminate_job(self, job):
        raise NotImplementedError

    def get_listen_addr(self):
        raise NotImplementedError


@dataclass
class DeviceCommand:
    """Launch description of a device process: which compiled-in process body runs, on which lanes."""
    body: int                         # fbr_process_kind
    lane_in: Any = None               # reader lane handle (c_void_p) or None
    lane_out: Any = None              # writer lane handle (c_void_p) or None
    ident: int = 0
    msg: Any = None                   # _abi.Record or None
    records: Any = None               # ctypes array of _abi.Record (put_queue(list)) or None
    idle_timeout: float = 30.0
    keepalive: list = field(default_factory=list)
