"""``fiber_b200.meta`` -- the resource-hint decorator of the reference (fiber/meta.py:16-58).

Same keys (``cpu``, ``memory`` -> stored as ``mem``, ``gpu``), same storage attribute
``func.__fiber_meta__``; the pool reads it when it starts its workers (fiber/pool.py:1122-1137).
"""

VALID_META_KEYS = ["cpu", "memory", "gpu"]


def post_process(metadata):
    # memory is given in MB and stored under "mem" (fiber/meta.py:19-25)
    if "memory" in metadata:
        metadata["mem"] = metadata.pop("memory")
    return metadata


def meta(**kwargs):
    for k in kwargs:
        assert k in VALID_META_KEYS, "Invalid meta argument \"{}\"".format(k)

    def decorator(func):
        func.__fiber_meta__ = post_process(dict(kwargs))
        return func

    return decorator
