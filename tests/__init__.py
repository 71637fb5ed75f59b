"""Parity and host-logic tests of the fiber_b200 Pool.map engine (CPU suite: -m "not gpu"; B200 suite: -m gpu)."""
