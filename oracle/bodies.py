"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the workload bodies on the Pool.map hot path.

Every function cites the reference line it restates (paths relative to /root/reference).
Two flavours per body: a scalar pure-Python one (the callable that is actually pushed through the
reference pool / the oracle pool port, so it must be picklable by module name) and a vectorised
NumPy one for million-task parity checks.

Bodies
------
* ``square`` / ``mul2`` / ``square_scale``     tests/test_pool.py:18-25 (``f``, ``f2``, ``fy``)
* ``pi_inside_det``                             examples/pi_estimation.py:9-11 made deterministic
                                                (SURVEY.md section 8(d) C2): Philox4x32-10 keyed by
                                                the task index replaces ``random.random()``
* ``parzen_estimation``                         examples/parzen_estimation.py:6-15
* ``payload_map`` / ``payload_checksum``        BASELINE.json config 4 (synthetic 4 KB payload map,
                                                SURVEY.md section 8(d) C4)
"""
import numpy as np

# ----------------------------------------------------------------------------------------------
# tests/test_pool.py:18-25
# ----------------------------------------------------------------------------------------------


def square(x):
    """``f`` in tests/test_pool.py:18-19."""
    return x * x


def mul2(x, y):
    """``f2`` in tests/test_pool.py:21-22."""
    return x * y


def square_scale(x, y=1):
    """``fy`` in tests/test_pool.py:24-25."""
    return x * x * y


def identity(i):
    """Return value of ``random_error_worker`` (tests/test_pool.py:60-68) without the fault."""
    return i


# ----------------------------------------------------------------------------------------------
# Philox4x32-10 (Salmon et al., SC'11; Random123 v1.09 philox.h).  Not in the reference: it is the
# counter-based RNG SURVEY.md 8(d) picks so that ``is_inside`` becomes a pure function of the task.
# Pinned by the Random123 known-answer vectors in tests/test_oracle.py.
# ----------------------------------------------------------------------------------------------
PHILOX_M0 = 0xD2511F53
PHILOX_M1 = 0xCD9E8D57
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85
PI_KEY = (0xF1BE5EED, 0x00000000)
_M32 = 0xFFFFFFFF


def philox4x32_10(ctr, key):
    """Scalar Philox4x32 with 10 rounds. ``ctr`` 4 x u32, ``key`` 2 x u32 -> 4 x u32."""
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for rnd in range(10):
        p0 = PHILOX_M0 * c0
        p1 = PHILOX_M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & _M32, p1 & _M32, ((p0 >> 32) ^ c3 ^ k1) & _M32, p0 & _M32
        k0 = (k0 + PHILOX_W0) & _M32
        k1 = (k1 + PHILOX_W1) & _M32
    return c0, c1, c2, c3


def philox4x32_10_np(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10 over uint64 arrays holding u32 values."""
    c0 = c0.astype(np.uint64)
    c1 = c1.astype(np.uint64)
    c2 = c2.astype(np.uint64)
    c3 = c3.astype(np.uint64)
    m32 = np.uint64(_M32)
    s32 = np.uint64(32)
    k0 = np.uint64(k0)
    k1 = np.uint64(k1)
    for rnd in range(10):
        p0 = np.uint64(PHILOX_M0) * c0
        p1 = np.uint64(PHILOX_M1) * c2
        n0 = ((p1 >> s32) ^ c1 ^ k0) & m32
        n2 = ((p0 >> s32) ^ c3 ^ k1) & m32
        c1 = p1 & m32
        c3 = p0 & m32
        c0, c2 = n0, n2
        k0 = (k0 + np.uint64(PHILOX_W0)) & m32
        k1 = (k1 + np.uint64(PHILOX_W1)) & m32
    return c0, c1, c2, c3


def pi_uniforms(p):
    """The two doubles task ``p`` draws: CPython's ``random.random()`` construction
    (Modules/_randommodule.c: ``(a>>5)*67108864.0 + (b>>6)) / 9007199254740992.0``) fed from one
    Philox block with counter ``(p_lo, p_hi, 0, 0)`` and key ``PI_KEY``."""
    p &= 0xFFFFFFFFFFFFFFFF  # int64 two's complement view of negative indices
    a, b, c, d = philox4x32_10((p & _M32, p >> 32, 0, 0), PI_KEY)
    x = ((a >> 5) * 67108864.0 + (b >> 6)) / 9007199254740992.0
    y = ((c >> 5) * 67108864.0 + (d >> 6)) / 9007199254740992.0
    return x, y


def pi_inside_det(p):
    """Deterministic ``is_inside`` (examples/pi_estimation.py:9-11, tests/test_pool.py:70-72):
    same arithmetic ``x * x + y * y < 1`` (three separately rounded float64 ops), but ``x, y`` are a
    pure function of the task argument instead of the per-process Mersenne Twister."""
    x, y = pi_uniforms(p)
    return x * x + y * y < 1


def pi_inside_det_np(start, stop, step=1):
    """``[pi_inside_det(p) for p in range(start, stop, step)]`` as a uint8 array."""
    p = np.arange(start, stop, step, dtype=np.int64).view(np.uint64)
    lo = p & np.uint64(_M32)
    hi = p >> np.uint64(32)
    z = np.zeros_like(p)
    a, b, c, d = philox4x32_10_np(lo, hi, z, z, *PI_KEY)
    x = ((a >> np.uint64(5)).astype(np.float64) * 67108864.0 + (b >> np.uint64(6)).astype(np.float64)) / 9007199254740992.0
    y = ((c >> np.uint64(5)).astype(np.float64) * 67108864.0 + (d >> np.uint64(6)).astype(np.float64)) / 9007199254740992.0
    return (x * x + y * y < 1.0).astype(np.uint8)


# ----------------------------------------------------------------------------------------------
# examples/parzen_estimation.py:6-15
# ----------------------------------------------------------------------------------------------


def pi_inside_bits8(g, start=0, step=1):
    """Body ``pi_inside_bits8`` (include/fiber_b200.h, FBR_RES_BITS8): the bool results of the 8 range()
    indices ``8g .. 8g+7`` of ``range(start, ..., step)`` in one byte, bit k (LSB first) = index 8g+k.
    Same per-index function as ``pi_inside_det`` (examples/pi_estimation.py:9-11): only the result layout
    differs, the reference's list of bools is ``unpackbits(bytes, bitorder="little")[:n]``."""
    b = 0
    for k in range(8):
        b |= int(pi_inside_det(start + (8 * g + k) * step)) << k
    return b


def pi_inside_bits_np(start, n, step=1):
    """The ``ceil(n/8)`` result bytes of a bit-packed map over ``range(start, start + n*step, step)``, bits past
    ``n`` cleared (what ``fiber_b200.Pool(results="bits")`` hands back as ``ResultArray.packed``)."""
    return np.packbits(pi_inside_det_np(start, start + n * step, step), bitorder="little")


def parzen_estimation(x_samples, point_x, h):
    """Restatement of examples/parzen_estimation.py:6-15 (hypercube Parzen window).

    A sample counts when, for every dimension d, ``abs((point_x[d] - x[d]) / h) <= 1/2`` (the
    reference breaks out of the row loop on ``> 1/2``, lines 9-13).  The normaliser is
    ``h ** point_x.shape[1]`` -- with the example's ``point_x.shape == (2, 1)`` that is ``h**1``
    (line 15), a reference quirk that parity must reproduce.
    """
    k_n = 0
    for sample in x_samples:
        inside = True
        for d in range(len(sample)):
            if abs((point_x[d][0] - sample[d]) / h) > (1 / 2):
                inside = False
                break
        if inside:
            k_n += 1
    return (h, (k_n / len(x_samples)) / (h ** point_x.shape[1]))


def parzen_count_np(x_samples, point_x, h, dtype=np.float64):
    """``k_n`` of the above, vectorised, evaluated in ``dtype`` arithmetic (float64 = reference,
    float32 = what the north-star's fp32 device path computes)."""
    xs = np.asarray(x_samples, dtype=dtype)
    px = np.asarray(point_x, dtype=dtype).reshape(1, -1)
    hh = dtype(h)
    q = np.abs((px - xs) / hh)
    return int(np.count_nonzero(~(q > dtype(0.5)).any(axis=1)))


def parzen_estimation_np(x_samples, point_x, h, dtype=np.float64):
    k_n = parzen_count_np(x_samples, point_x, h, dtype)
    return (h, (k_n / len(x_samples)) / (h ** np.asarray(point_x).shape[1]))


def parzen_boundary_count(x_samples, point_x, h):
    """Samples whose fp32 inside/outside decision may legitimately differ from fp64: those with
    ``|q - 0.5| <= 2^-22 * max(1, q)`` for some d (SURVEY.md 8(d) C3 tolerance statement)."""
    xs = np.asarray(x_samples, dtype=np.float64)
    px = np.asarray(point_x, dtype=np.float64).reshape(1, -1)
    q = np.abs((px - xs) / np.float64(h))
    near = np.abs(q - 0.5) <= (2.0 ** -22) * np.maximum(1.0, q)
    return int(np.count_nonzero(near.any(axis=1)))


def parzen_example_inputs():
    """Inputs exactly as examples/parzen_estimation.py:32-40 builds them."""
    np.random.seed(123)
    mu_vec = np.array([0, 0])
    cov_mat = np.array([[1, 0], [0, 1]])
    x_2Dgauss = np.random.multivariate_normal(mu_vec, cov_mat, 10000)
    widths = np.arange(0.1, 10.3, 0.1)
    point_x = np.array([[0], [0]])
    return x_2Dgauss, point_x, widths


# ----------------------------------------------------------------------------------------------
# Synthetic 4 KB payload map (BASELINE.json config 4).  SplitMix64 finaliser
# (Steele/Lea/Flood 2014; java.util.SplittableRandom) -- pinned by its known first output
# 0xE220A8397B1DCDAF for state 0.
# ----------------------------------------------------------------------------------------------
PAYLOAD_WORDS = 1024
PAYLOAD_SEED = 0xF1BE5
PAYLOAD_MUL = 2654435761
_M64 = 0xFFFFFFFFFFFFFFFF


def splitmix64(x):
    """One SplitMix64 output for state ``x`` (state is advanced by the golden gamma first)."""
    z = (x + 0x9E3779B97F4A7C15) & _M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


def splitmix64_np(x):
    with np.errstate(over="ignore"):
        z = x.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def payload_record(t):
    """Input record of task ``t``: 1024 x u32, ``w[j] = low32(splitmix64(SEED ^ (t*1024 + j)))``."""
    return [splitmix64(PAYLOAD_SEED ^ (t * PAYLOAD_WORDS + j)) & _M32 for j in range(PAYLOAD_WORDS)]


def payload_records_np(t0, t1):
    """Records of tasks ``[t0, t1)`` as a ``(t1-t0, 1024)`` uint32 array."""
    idx = np.arange(t0 * PAYLOAD_WORDS, t1 * PAYLOAD_WORDS, dtype=np.uint64)
    w = splitmix64_np(np.uint64(PAYLOAD_SEED) ^ idx) & np.uint64(_M32)
    return w.astype(np.uint32).reshape(t1 - t0, PAYLOAD_WORDS)


def payload_map(t, rec):
    """Body of the synthetic map: ``out[j] = rec[j] * 2654435761 + t`` (u32 wrap-around)."""
    return [(w * PAYLOAD_MUL + t) & _M32 for w in rec]


def payload_map_np(t0, recs):
    """Vectorised ``payload_map`` for tasks ``t0 .. t0+len(recs)``."""
    t = (np.arange(t0, t0 + recs.shape[0], dtype=np.uint64) & np.uint64(_M32)).astype(np.uint32)
    with np.errstate(over="ignore"):
        return recs.astype(np.uint32) * np.uint32(PAYLOAD_MUL) + t[:, None]


def payload_checksum(t, rec):
    """Secondary variant: 4 B result ``sum(rec) mod 2^32`` (``t`` unused, kept for the signature)."""
    return sum(rec) & _M32


def payload_checksum_np(recs):
    return (recs.astype(np.uint64).sum(axis=1) & np.uint64(_M32)).astype(np.uint32)
