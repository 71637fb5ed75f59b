"""GPU: randomised parity sweep.  Many (body, n, chunksize, ring size, flags) combinations against
the oracle -- claim-unit sizing, wave cutting, tail units, unaligned slots, shuffled arrival, the
flat vs row gather paths and multi-threaded submission all have to agree bit for bit."""
import ctypes
import threading

import numpy as np
import pytest

import fiber_b200
from fiber_b200 import _abi, registry
from oracle import cref

from . import workloads as W

pytestmark = pytest.mark.gpu


def _raw(lib, h, body, n, flags, chunksize, start=0, step=1, args=None, stride=0, seed=1):
    spec = registry.spec(body)
    d = _abi.MapDesc()
    d.func_id, d.flags, d.n_tasks, d.chunksize = spec.func_id, flags, n, chunksize
    d.index_start, d.index_step, d.shuffle_seed = start, step, seed
    if args is not None:
        d.args, d.arg_stride = args.ctypes.data, stride
    seq = ctypes.c_uint64()
    _abi.check(lib.fbr_map_submit(h, ctypes.byref(d), ctypes.byref(seq)))
    res = _abi.Result()
    _abi.check(lib.fbr_result_wait(h, seq.value, -1, ctypes.byref(res)))
    out = np.frombuffer((ctypes.c_char * max(1, n * spec.result_bytes)).from_address(res.data), dtype=np.uint8)[: n * spec.result_bytes].copy()
    s, waves = int(res.sum), int(res.n_waves)
    _abi.check(lib.fbr_result_release(h, seq.value))
    return out, s, waves


@pytest.mark.parametrize("ring_kib", [64, 1024, 65536])
def test_random_maps_against_oracle(ring_kib):
    rng = np.random.default_rng(1234 + ring_kib)
    lib = _abi.load()
    ids = (ctypes.c_int * 1)(0)
    h = ctypes.c_void_p()
    _abi.check(lib.fbr_pool_create(1, ids, ring_kib << 10, 0, ctypes.byref(h)))
    try:
        for trial in range(40):
            n = int(rng.choice([1, 2, 15, 16, 17, 31, 33, 255, 4095, 4097, 65537, int(rng.integers(1, 300000))]))
            cs = int(rng.choice([0, 1, 3, 7, 32, 100, 4096, 100000]))
            flags = _abi.FBR_WANT_SUM
            if rng.random() < 0.4:
                flags |= _abi.FBR_SHUFFLE
            if rng.random() < 0.3:
                flags |= _abi.FBR_FULL_WINDOW
            if rng.random() < 0.4:
                flags |= _abi.FBR_VIA_RING          # task records + result ring + gather_ordered instead of direct placement
            if rng.random() < 0.2:
                flags |= _abi.FBR_RESILIENT
                flags &= ~_abi.FBR_SHUFFLE
            start = int(rng.integers(-10 ** 6, 10 ** 6))
            step = int(rng.choice([1, 1, 3, -2]))
            kind = trial % 3
            if kind == 0:        # pi over a range: uint8 results + count
                out, s, _ = _raw(lib, h, "pi_inside_det", n, flags, cs, start, step)
                ref, count = cref.pi_inside_range(start, n, step)
                assert np.array_equal(out, ref) and s == count, (n, cs, flags, start, step)
            elif kind == 1:      # square over explicit int64 records
                xs = rng.integers(-3 * 10 ** 9, 3 * 10 ** 9, size=n, dtype=np.int64)
                out, s, _ = _raw(lib, h, "square_i64", n, flags, cs, args=xs, stride=8)
                assert np.array_equal(out.view(np.int64), xs * xs) and s == int((xs * xs).sum()), (n, cs, flags)
            else:                # 4 KB records, map + checksum (odd chunk sizes -> unaligned flat gather)
                m = min(n, 3000)
                recs = cref.payload_records(start % 1000, m)
                out, _, _ = _raw(lib, h, "payload_map_4k", m, flags & ~_abi.FBR_WANT_SUM, cs, args=recs, stride=4096)
                assert np.array_equal(out.view(np.uint32).reshape(m, 1024), cref.payload_map(0, recs)), (m, cs, flags)
                out, s, _ = _raw(lib, h, "payload_checksum_4k", m, flags, cs, args=recs, stride=4096)
                ck = cref.payload_checksum(recs)
                assert np.array_equal(out.view(np.uint32), ck) and s == int(ck.astype(np.int64).sum()), (m, cs, flags)
    finally:
        lib.fbr_pool_destroy(h)


def test_concurrent_submitters_share_one_pool():
    """Several host threads submit and collect maps on one pool at the same time (the reference's
    sockets are not thread-safe, fiber/pool.py:224-227; the engine serialises on its own lock)."""
    pool = fiber_b200.Pool(1, ring_bytes=4 << 20)
    pool.start_workers()
    errors = []

    def run(tid):
        try:
            rng = np.random.default_rng(tid)
            for _ in range(25):
                n = int(rng.integers(1, 200000))
                lo = int(rng.integers(0, 10 ** 6))
                res = pool.map(W.is_inside, range(lo, lo + n))
                ref, count = cref.pi_inside_range(lo, n)
                assert res.sum() == count and np.array_equal(np.asarray(res).view(np.uint8), ref)
                assert pool.apply(W.f, (tid,)) == tid * tid
        except Exception as e:     # noqa: BLE001
            errors.append(e)
    threads = [threading.Thread(target=run, args=(t,)) for t in range(6)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:1]
    pool.terminate()
    pool.join()


@pytest.mark.parametrize("ring_kib", [64, 4096])
def test_random_pi_ranges_bytes_and_bits(ring_kib):
    """Random range(start, stop, step) maps of the pi body through both result layouts (one byte / one bit per
    task) against the plain-C oracle: starts around 0, 2^32 and 2^40, negative and large steps, lengths that
    leave partial vectors, partial bytes and partial claim units, small rings (many waves)."""
    rng = np.random.default_rng(99 + ring_kib)
    pools = [fiber_b200.Pool(1, ring_bytes=ring_kib << 10, results="bytes"), fiber_b200.Pool(1, ring_bytes=ring_kib << 10)]
    try:
        for trial in range(30):
            n = int(rng.choice([1, 7, 8, 9, 15, 16, 17, 127, 4095, 4096, 4097, 32769, int(rng.integers(1, 200000))]))
            anchor = int(rng.choice([0, 2 ** 32, 2 ** 40, -2 ** 33]))
            start = anchor + int(rng.integers(-70000, 70000))
            step = int(rng.choice([1, 1, 2, 5, -1, -3, 2 ** 31 + 7, 2 ** 20]))
            cs = int(rng.choice([1, 8, 32, 100, 5000]))
            ref, count = cref.pi_inside_range(start, n, step)
            r = range(start, start + n * step, step)
            for pool in pools:
                res = pool.map(W.is_inside, r, cs)
                assert len(res) == n and res.sum() == count, (start, n, step, cs)
                assert np.array_equal(np.asarray(res).view(np.uint8), ref), (start, n, step, cs)
            assert np.array_equal(res.packed, np.packbits(ref, bitorder="little")), (start, n, step, cs)
            # the same tasks as explicit argument records (a list, not a range): 8 int64 items per byte-task
            if n <= 70000:
                items = [start + i * step for i in range(n)]
                for pool in pools:
                    res = pool.map(W.is_inside, items, cs)
                    assert len(res) == n and res.sum() == count, (start, n, step, cs)
                    assert np.array_equal(np.asarray(res).view(np.uint8), ref), (start, n, step, cs)
                assert np.array_equal(res.packed, np.packbits(ref, bitorder="little")), (start, n, step, cs)
    finally:
        for pool in pools:
            pool.terminate()
            pool.join()
