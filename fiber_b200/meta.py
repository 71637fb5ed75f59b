"""Resource-hint decorator with the reference's contract (fiber/meta.py:28-58): hints are attached to
the function as ``__fiber_meta__``; the accepted names are ``cpu``, ``memory`` (stored as ``mem``,
megabytes) and ``gpu``.  The pool compares this attribute when it starts its workers
(fiber/pool.py:1122-1137)."""

VALID_META_KEYS = ["cpu", "memory", "gpu"]
_STORED_AS = {"memory": "mem"}


def post_process(metadata):
    """Rename hint keys to the names they are stored under."""
    return {_STORED_AS.get(name, name): value for name, value in metadata.items()}


def meta(**hints):
    unknown = [name for name in hints if name not in VALID_META_KEYS]
    assert not unknown, "Invalid meta argument \"{}\"".format(unknown[0])
    stored = post_process(hints)

    def attach(func):
        func.__fiber_meta__ = dict(stored)
        return func

    return attach
