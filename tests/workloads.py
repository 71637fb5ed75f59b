"""Re-export of the user-side function definitions (examples/workloads.py)."""
from examples.workloads import *  # noqa: F401,F403
from examples.workloads import f, f2, fy, identity, is_inside, parzen_estimation, parzen_estimation_f32, \
    payload_checksum, payload_map, random_error_worker, sleep_worker, unbound  # noqa: F401
