"""examples/parzen_estimation.py of the reference, on the B200 engine."""
import numpy as np

from fiber_b200 import Pool

from examples.workloads import parzen_estimation


def multiprocess(processes, samples, x, widths):
    pool = Pool(processes=processes)
    results = [pool.apply_async(parzen_estimation, args=(samples, x, w)) for w in widths]
    results = [p.get() for p in results]
    results.sort()  # to sort the results by input window width
    return results


def main():
    np.random.seed(123)
    mu_vec = np.array([0, 0])
    cov_mat = np.array([[1, 0], [0, 1]])
    x_2Dgauss = np.random.multivariate_normal(mu_vec, cov_mat, 10000)
    widths = np.arange(0.1, 10.3, 0.1)
    point_x = np.array([[0], [0]])
    for r in multiprocess(4, x_2Dgauss, point_x, widths):
        print('h = %s, p(x) = %s' % (r[0], r[1]))


if __name__ == '__main__':
    main()
