"""Parzen-window density sweep on the B200 pool: one apply_async per window width, as the
reference's parzen example does, with the 10 000 x 2 sample block uploaded once.

    python -m examples.parzen_window [f32]
"""
import sys

import numpy as np

import fiber_b200

from .workloads import parzen_estimation, parzen_estimation_f32


def sweep(samples, point, widths, body):
    pool = fiber_b200.Pool(processes=4)
    pending = [pool.apply_async(body, args=(samples, point, h)) for h in widths]
    return sorted(r.get() for r in pending)


if __name__ == "__main__":
    rng_state = np.random.seed(123)
    cloud = np.random.multivariate_normal(np.zeros(2), np.eye(2), 10000)
    hs = np.arange(0.1, 10.3, 0.1)
    origin = np.array([[0], [0]])
    body = parzen_estimation_f32 if "f32" in sys.argv[1:] else parzen_estimation
    for h, density in sweep(cloud, origin, hs, body):
        print("h = %s, p(x) = %s" % (h, density))
