"""experimental.Ring (fiber/experimental/ring.py:44-129; BASELINE.json config 5)."""
import json
import os
import tempfile

import pytest

from fiber_b200.experimental import Ring, RingNode, engine_ring_init, torch_ring_init

from . import _ring_func


def _run_ring(size, elems, func=_ring_func.allreduce_node, init=torch_ring_init):
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "ring")
        os.environ["FBR_RING_OUT"] = out
        os.environ["FBR_RING_ELEMS"] = str(elems)
        try:
            ring = Ring(size, func, init)
            assert [m.rank for m in ring.members] == list(range(size)) and isinstance(ring.members[0], RingNode)
            ring.run()
        finally:
            del os.environ["FBR_RING_OUT"], os.environ["FBR_RING_ELEMS"]
        return [json.load(open("%s.%d" % (out, r))) for r in range(size)]


def test_ring_world_size_2_gloo_cpu():
    """Bootstrap + all-reduce with world_size 2 on CPU (gloo), exact small-integer sums."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU variant")
    res = _run_ring(2, 65536)
    assert [r["rank"] for r in res] == [0, 1] and all(r["ok"] for r in res) and res[0]["backend"] == "gloo"


def test_ring_size_zero_is_a_noop():
    Ring(0, _ring_func.allreduce_node, torch_ring_init).run()      # ring.py:108-109


@pytest.mark.gpu
def test_ring_allreduce_nccl():
    """All GPUs of the box: ncclAllReduce of a 64 Mi-element fp32 buffer (256 MiB), bit-exact."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    res = _run_ring(n, 64 * 1024 * 1024)
    assert all(r["ok"] for r in res) and res[0]["backend"] == "nccl"


@pytest.mark.gpu
def test_ring_allreduce_engine_comm():
    """The same ring on the engine's own communicator (fbr_comm_* = NCCL behind the C ABI): the bootstrap id
    travels in the member table, every collective of the one-process-per-GPU mode is checked bit-exact."""
    import ctypes
    from fiber_b200 import _abi
    n = ctypes.c_int(0)
    _abi.check(_abi.load().fbr_device_count(ctypes.byref(n)))
    if n.value < 2:
        pytest.skip("needs >= 2 GPUs")
    res = _run_ring(n.value, 64 * 1024 * 1024, _ring_func.engine_allreduce_node, engine_ring_init)
    assert all(r["ok"] for r in res), [r["oks"] for r in res]
    assert res[0]["backend"].startswith("fbr_comm/nccl")


def test_comm_bootstrap_id_needs_no_gpu():
    """ncclGetUniqueId through the C ABI works on a host without a GPU (the ring parent makes it before the
    nodes start); building a communicator there fails loudly."""
    from fiber_b200 import _abi, comm
    assert comm.load_nccl() >= 21800
    a, b = comm.unique_id(), comm.unique_id()
    assert len(a) == comm.ID_BYTES and a != b
    import ctypes
    n = ctypes.c_int(0)
    if _abi.load().fbr_device_count(ctypes.byref(n)) != 0 or n.value == 0:
        with pytest.raises(_abi.EngineError) as ei:
            comm.Comm(0, 1, 0, a)
        assert ei.value.status == _abi.FBR_ENODEV
