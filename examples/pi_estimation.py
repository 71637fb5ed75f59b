"""examples/pi_estimation.py of the reference, on the B200 engine (same shape, 1e8 samples)."""
import sys

from fiber_b200 import Pool

from examples.workloads import is_inside

NUM_SAMPLES = int(float(sys.argv[1])) if len(sys.argv) > 1 else int(1e8)


def main():
    pool = Pool(processes=4)
    pi = 4.0 * pool.map(is_inside, range(0, NUM_SAMPLES)).sum() / NUM_SAMPLES
    print("Pi is roughly {}".format(pi))


if __name__ == '__main__':
    main()
