"""The mapped functions of the reference's own tests/examples, written the way a user writes them
and bound to their device bodies.  The Python bodies document the semantics (and are what the CPU
oracle pool executes); fiber_b200.Pool never calls them."""
import fiber_b200


@fiber_b200.device_body("square_i64")
def f(x):                      # tests/test_pool.py:18-19
    return x * x


@fiber_b200.device_body("mul2_i64")
def f2(x, y):                  # tests/test_pool.py:21-22
    return x * y


@fiber_b200.device_body("square_scale_i64")
def fy(x, y=1):                # tests/test_pool.py:24-25
    return x * x * y


@fiber_b200.device_body("identity_i64")
def identity(i):
    return i


@fiber_b200.device_body("fault_identity_i64")
def random_error_worker(i):    # tests/test_pool.py:60-68 (faults injected on the device)
    return i


@fiber_b200.device_body("sleep_f64")
def sleep_worker(duration):    # tests/test_pool.py:56-57
    import time
    time.sleep(duration)


@fiber_b200.device_body("pi_inside_det")
def is_inside(p):              # examples/pi_estimation.py:9-11, deterministic restatement
    from oracle.bodies import pi_inside_det
    return pi_inside_det(p)


@fiber_b200.device_body("parzen_f64")
def parzen_estimation(x_samples, point_x, h):   # examples/parzen_estimation.py:6-15
    from oracle.bodies import parzen_estimation as ref
    return ref(x_samples, point_x, h)


@fiber_b200.device_body("parzen_f32")
def parzen_estimation_f32(x_samples, point_x, h):
    from oracle.bodies import parzen_estimation as ref
    return ref(x_samples, point_x, h)


@fiber_b200.device_body("payload_map_4k")
def payload_map(t, rec):
    from oracle.bodies import payload_map as ref
    return ref(t, rec)


@fiber_b200.device_body("payload_checksum_4k")
def payload_checksum(t, rec):
    from oracle.bodies import payload_checksum as ref
    return ref(t, rec)


def unbound(x):
    return x + 1
