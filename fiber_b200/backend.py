"""Backend registry with the reference's selection rules (fiber/backend.py:56-76): backends are
looked up by name and imported lazily from ``fiber_b200.<name>_backend``; an unknown name raises
``multiprocessing.ProcessError``; instances are cached per name.  Tests swap instances in
``_backends`` for fault injection exactly as the reference's tests do
(tests/test_process.py:27-39, 180-190)."""
import importlib
import multiprocessing as mp

_backends = {}
available_backend = ["gpu"]
default_backend = "gpu"


def get_backend(name=None, **kwargs):
    if name is None:
        name = default_backend
    elif name not in available_backend:
        raise mp.ProcessError("Invalid backend: {}".format(name))
    inst = _backends.get(name)
    if inst is None:
        inst = importlib.import_module("fiber_b200.{}_backend".format(name)).Backend(**kwargs)
        _backends[name] = inst
    return inst
