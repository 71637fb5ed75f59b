"""TEST INFRASTRUCTURE ONLY -- ctypes loader for the plain-C oracle (oracle/fbr_oracle.c)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libfbr_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "fbr_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s", "_build/libfbr_oracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        i64, p = ctypes.c_int64, ctypes.c_void_p
        L.orc_pi_inside_range.restype = i64
        L.orc_pi_inside_range.argtypes = [i64, i64, i64, p]
        L.orc_pi_inside_one.restype = ctypes.c_int
        L.orc_pi_inside_one.argtypes = [i64]
        L.orc_parzen_count_f64.restype = i64
        L.orc_parzen_count_f64.argtypes = [p, i64, ctypes.c_int, p, ctypes.c_double]
        L.orc_parzen_count_f32.restype = i64
        L.orc_parzen_count_f32.argtypes = [p, i64, ctypes.c_int, p, ctypes.c_float]
        L.orc_parzen_density.restype = ctypes.c_double
        L.orc_parzen_density.argtypes = [i64, i64, ctypes.c_double, ctypes.c_int]
        L.orc_splitmix64.restype = ctypes.c_uint64
        L.orc_splitmix64.argtypes = [ctypes.c_uint64]
        L.orc_payload_records.argtypes = [i64, i64, p]
        L.orc_payload_map.argtypes = [i64, i64, p, p]
        L.orc_payload_checksum.argtypes = [i64, p, p]
        L.orc_place_by_index.argtypes = [p, p, i64, i64, i64, i64, p]
        L.orc_philox4x32_10.argtypes = [p, p, p]
        L.orc_square_i64.argtypes = [p, i64, p, p]
        _lib = L
    return _lib


def pi_inside_range(start, n, step=1, want_array=True):
    out = np.empty(n, dtype=np.uint8) if want_array else None
    cnt = lib().orc_pi_inside_range(start, step, n, out.ctypes.data if want_array else None)
    return out, int(cnt)


def pi_inside_det_c(p):
    """Picklable-by-name callable: the deterministic pi body at C speed (for the CPU pool arm)."""
    return bool(lib().orc_pi_inside_one(p))


def parzen_count(xs, px, h, dtype=np.float64):
    xs = np.ascontiguousarray(xs, dtype=dtype)
    px = np.ascontiguousarray(np.asarray(px).reshape(-1), dtype=dtype)
    fn = lib().orc_parzen_count_f64 if dtype == np.float64 else lib().orc_parzen_count_f32
    return int(fn(xs.ctypes.data, xs.shape[0], xs.shape[1], px.ctypes.data, float(h)))


def payload_records(t0, n):
    out = np.empty((n, 1024), dtype=np.uint32)
    lib().orc_payload_records(t0, n, out.ctypes.data)
    return out


def payload_map(t0, recs):
    recs = np.ascontiguousarray(recs, dtype=np.uint32)
    out = np.empty_like(recs)
    lib().orc_payload_map(t0, recs.shape[0], recs.ctypes.data, out.ctypes.data)
    return out


def payload_checksum(recs):
    recs = np.ascontiguousarray(recs, dtype=np.uint32)
    out = np.empty(recs.shape[0], dtype=np.uint32)
    lib().orc_payload_checksum(recs.shape[0], recs.ctypes.data, out.ctypes.data)
    return out


def place_by_index(ring, order, chunk, n, result_bytes):
    ring = np.ascontiguousarray(ring, dtype=np.uint8)
    order = np.ascontiguousarray(order, dtype=np.int64)
    out = np.zeros(n * result_bytes, dtype=np.uint8)
    lib().orc_place_by_index(ring.ctypes.data, order.ctypes.data, len(order), chunk, n, result_bytes, out.ctypes.data)
    return out
