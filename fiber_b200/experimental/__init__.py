from .ring import Ring, RingNode, allreduce_bench, engine_ring_init, ring_comm, torch_ring_init  # noqa: F401
