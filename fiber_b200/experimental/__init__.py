from .ring import Ring, RingNode, allreduce_bench, torch_ring_init  # noqa: F401
