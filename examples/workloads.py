"""User-side function definitions of the reference's tests/examples, bound to their device bodies.

This is what a fiber user's module looks like after switching to fiber_b200: the functions keep
their names and signatures, and a decorator says which compiled-in device body runs them.
``fiber_b200.Pool`` never calls the Python bodies; for the bodies that are not one-liners the
Python body only documents the semantics and refuses to run (there is no CPU path).
"""
import fiber_b200


def _device_only(name):
    raise RuntimeError("%s is bound to a device body; fiber_b200 never executes it on the CPU" % name)


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
def random_error_worker(i):    # tests/test_pool.py:60-68: ~5 % of attempts kill their worker
    _device_only("random_error_worker")


@fiber_b200.device_body("sleep_f64")
def sleep_worker(duration):    # tests/test_pool.py:56-57
    _device_only("sleep_worker")


@fiber_b200.device_body("pi_inside_det")
def is_inside(p):
    """examples/pi_estimation.py:9-11 made a pure function of ``p``: ``x, y`` come from one
    Philox4x32-10 block (key 0xF1BE5EED, counter p) via CPython's 53-bit double construction, then
    ``x * x + y * y < 1`` in float64."""
    _device_only("is_inside")


@fiber_b200.device_body("parzen_f64")
def parzen_estimation(x_samples, point_x, h):
    """examples/parzen_estimation.py:6-15, float64 window test (bit-exact)."""
    _device_only("parzen_estimation")


@fiber_b200.device_body("parzen_f32")
def parzen_estimation_f32(x_samples, point_x, h):
    """Same with the samples cast once to float32 (the north-star's fp32 path)."""
    _device_only("parzen_estimation_f32")


@fiber_b200.device_initializer("parzen_f64")
def set_parzen_samples(x_samples, point_x):
    """Pool initializer (fiber/pool.py:858-859 runs it once per worker): the sample array every task shares
    becomes the workers' broadcast block, uploaded once per device."""
    _device_only("set_parzen_samples")


@fiber_b200.device_body("parzen_f64")
def parzen_at(h):
    """parzen_estimation with the samples taken from the pool's initializer block."""
    _device_only("parzen_at")


@fiber_b200.device_body("trap_identity_i64")
def trap_identity(i):
    """Identity whose first attempt at an argument with low 20 bits 0xDEAD executes `trap` on the GPU: the
    worker's CUDA context dies for real (the reference's worker process killed mid-chunk)."""
    _device_only("trap_identity")


@fiber_b200.device_body("payload_map_4k")
def payload_map(t, rec):
    """BASELINE.json config 4: ``[(w * 2654435761 + t) & 0xFFFFFFFF for w in rec]``, rec = 1024 u32."""
    _device_only("payload_map")


@fiber_b200.device_body("payload_checksum_4k")
def payload_checksum(t, rec):
    """``sum(rec) & 0xFFFFFFFF``."""
    _device_only("payload_checksum")


def unbound(x):
    return x + 1


# ---- process targets of the reference's tests/test_queue.py, bound to device process bodies ---------
@fiber_b200.device_process("queue_worker")
def worker(q_in, q_out, ident):            # tests/test_queue.py:44-50
    _device_only("worker")


@fiber_b200.device_process("put_queue")
def put_queue(q, data):                    # tests/test_queue.py:23-33
    _device_only("put_queue")


@fiber_b200.device_process("get_queue")
def get_queue(q_in, q_out, n):             # tests/test_queue.py:36-42
    _device_only("get_queue")


@fiber_b200.device_process("write_pipe")
def write_pipe(pipe, msg):                 # tests/test_queue.py:19-20
    _device_only("write_pipe")


@fiber_b200.device_process("pipe_worker")
def pipe_worker(conn):                     # tests/test_queue.py:53-57
    _device_only("pipe_worker")
