"""User-side programs for fiber_b200: the reference examples' workloads bound to device bodies."""
