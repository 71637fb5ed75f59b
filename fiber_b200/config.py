"""The two configuration knobs the pool path reads in the reference (fiber/config.py:132,135):

* ``cpu_per_job``    -- worker cores per job; the pool starts ``ceil(processes / cpu_per_job)`` jobs
                        (fiber/pool.py:1014, 1405-1408).  Here a job is a GPU and the value only
                        enters ``Pool.n_jobs``.
* ``use_push_queue`` -- ``SimpleQueue()`` returns the push queue or raises ``NotImplementedError``
                        (fiber/context.py:47-54).

Precedence as in the reference: defaults < ``FIBER_<KEY>`` environment < ``fiber_b200.init(**kw)`` /
direct assignment (fiber/config.py:18-20, 158-163, 221-249).  Unknown keys raise ``ValueError``.
"""
import os

_DEFAULTS = {"cpu_per_job": 1, "use_push_queue": True}
cpu_per_job = 1
use_push_queue = True


def _coerce(key, value):
    if isinstance(_DEFAULTS[key], bool):
        return value if isinstance(value, bool) else str(value).strip().lower() in ("1", "true", "yes", "on")
    return int(value)


def init(**kwargs):
    """Reset to defaults, apply ``FIBER_*`` environment variables, then ``kwargs``."""
    values = dict(_DEFAULTS)
    for key in _DEFAULTS:
        env = os.environ.get("FIBER_" + key.upper())
        if env is not None:
            values[key] = _coerce(key, env)
    for key, value in kwargs.items():
        if key not in _DEFAULTS:
            raise ValueError("invalid config key: {}".format(key))
        values[key] = _coerce(key, value)
    globals().update(values)
    return values


def reset():
    return init()


init()
