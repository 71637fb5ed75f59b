"""Monte-Carlo pi on the B200 pool: the workload of the reference's pi example
(examples/pi_estimation.py there) with the deterministic `is_inside` body.

    python -m examples.monte_carlo_pi [samples] [gpus]
"""
import sys
import time

import fiber_b200

from .workloads import is_inside


def estimate(samples, gpus):
    with_count = fiber_b200.Pool(processes=gpus).map(is_inside, range(samples))
    return 4.0 * with_count.sum() / samples


if __name__ == "__main__":
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10 ** 8
    g = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    t0 = time.perf_counter()
    pi = estimate(n, g)
    print("pi ~= %.8f from %d samples on %d GPU(s) in %.3f s" % (pi, n, g, time.perf_counter() - t0))
