"""The GPU backend: what ``fiber/local_backend.py:26-72`` is to subprocesses, this is to GPU-resident
device processes.  ``create_job`` launches the resident one-warp kernel described by the job's
``DeviceCommand`` on device ``job_spec.gpu`` (``fbr_process_start``); status / wait / terminate map to
``fbr_process_poll`` / ``fbr_process_join`` / ``fbr_process_terminate``."""
import ctypes

from . import _abi, core
from .core import ProcessStatus


class Backend(core.Backend):
    name = "gpu"

    def __init__(self):
        self._next_jid = 0

    def create_job(self, job_spec):
        cmd = job_spec.command
        if not isinstance(cmd, core.DeviceCommand):
            raise TypeError("the gpu backend runs DeviceCommand jobs (compiled-in process bodies), got %r" % (cmd,))
        lib = _abi.load()
        handle = ctypes.c_void_p()
        _abi.qcheck(lib.fbr_process_start(
            job_spec.gpu or 0, cmd.body, cmd.lane_in, cmd.lane_out, int(cmd.ident),
            ctypes.byref(cmd.msg) if cmd.msg is not None else None,
            cmd.records, len(cmd.records) if cmd.records is not None else 0,
            int(cmd.idle_timeout * 1000), ctypes.byref(handle)))
        self._next_jid += 1
        job = core.Job(data=handle, jid=self._next_jid)
        job.host = "cuda:%d" % (job_spec.gpu or 0)
        job.exitcode = None
        return job

    def _poll(self, job):
        alive, code = ctypes.c_int(1), ctypes.c_int(0)
        _abi.qcheck(_abi.load().fbr_process_poll(job.data, ctypes.byref(alive), ctypes.byref(code)))
        if not alive.value:
            job.exitcode = code.value
        return bool(alive.value)

    def get_job_status(self, job):
        if job.exitcode is not None or not self._poll(job):
            return ProcessStatus.STOPPED
        return ProcessStatus.STARTED

    def wait_for_job(self, job, timeout):
        if job.exitcode is not None:
            return job.exitcode
        if timeout == 0:
            return None if self._poll(job) else job.exitcode
        rc = _abi.load().fbr_process_join(job.data, -1 if timeout is None else int(timeout * 1000))
        if rc == _abi.FBR_ETIMEOUT:
            return None
        _abi.qcheck(rc)
        self._poll(job)
        return job.exitcode

    def terminate_job(self, job):
        _abi.qcheck(_abi.load().fbr_process_terminate(job.data))

    def handled(self, job):
        n = ctypes.c_uint64(0)
        _abi.qcheck(_abi.load().fbr_process_handled(job.data, ctypes.byref(n)))
        return n.value

    def release_job(self, job):
        h, job.data = job.data, None
        if h:
            _abi.load().fbr_process_destroy(h)

    def get_listen_addr(self):
        # the reference returns (ip, port, interface) for TCP rendezvous; lanes need no address
        return "gpu", 0, "nvlink"
