"""Small, fast exercise of every kernel for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import fiber_b200  # noqa: E402
from examples import workloads as W  # noqa: E402
from oracle import bodies as B, cref  # noqa: E402

import ctypes  # noqa: E402

from fiber_b200 import _abi, registry  # noqa: E402
from tests import device_bodies as D  # noqa: E402


def raw(pool, body, n, flags, chunksize=0, args=None, stride=0):
    """Map through the C ABI with explicit flags (shuffled arrival / via-ring / full window)."""
    spec, eng = registry.spec(body), pool._engine
    d = _abi.MapDesc()
    d.func_id, d.flags, d.n_tasks, d.chunksize, d.index_step, d.shuffle_seed = spec.func_id, flags, n, chunksize, 1, 7
    if args is not None:
        d.args, d.arg_stride = args.ctypes.data, stride
    seq, res = ctypes.c_uint64(), _abi.Result()
    _abi.check(eng.lib.fbr_map_submit(eng.handle, ctypes.byref(d), ctypes.byref(seq)))
    _abi.check(eng.lib.fbr_result_wait(eng.handle, seq.value, -1, ctypes.byref(res)))
    out = np.frombuffer((ctypes.c_char * (n * spec.result_bytes)).from_address(res.data), dtype=np.uint8).copy()
    _abi.check(eng.lib.fbr_result_release(eng.handle, seq.value))
    return out


pool = fiber_b200.Pool(1, ring_bytes=1 << 20, express=False, results="bytes")
assert pool.map(W.f, range(1000)) == [i * i for i in range(1000)]
assert pool.map(W.f, list(range(-50, 51)), 7) == [i * i for i in range(-50, 51)]
n = 300_017
res = pool.map(W.is_inside, range(n))
ref, count = cref.pi_inside_range(0, n)
assert res.sum() == count and np.array_equal(np.asarray(res).view(np.uint8), ref)
# ring path: records + result ring + gather_ordered (rows / flat), shuffled arrival and via-ring
assert np.array_equal(raw(pool, "pi_inside_det", n, _abi.FBR_SHUFFLE), ref)
assert np.array_equal(raw(pool, "pi_inside_det", n, _abi.FBR_VIA_RING | _abi.FBR_FULL_WINDOW), ref)
assert np.array_equal(raw(pool, "square_i64", 10007, _abi.FBR_SHUFFLE, 7).view(np.int64), np.arange(10007, dtype=np.int64) ** 2)
bp = fiber_b200.Pool(1, ring_bytes=1 << 20)                                                            # default: pi_inside_bits8, zero-copy stores
rb = bp.map(W.is_inside, range(n))
assert rb.sum() == count and np.array_equal(rb.packed, np.packbits(ref, bitorder="little"))
rb = bp.map(W.is_inside, list(range(n)))                                                               # explicit items: ballot kernel
assert rb.sum() == count and np.array_equal(rb.packed, np.packbits(ref, bitorder="little"))
assert list(bp.imap(W.is_inside, range(50003))) == ref[:50003].astype(bool).tolist()                  # staged waves (no zero copy)
assert np.array_equal(np.asarray(bp.map(D.collatz_steps, range(1, 20001))), D.collatz_steps_np(np.arange(1, 20001)))   # out-of-tree body
ob = bp.map(D.odd_bits, range(-500, 70001))                                                            # out-of-tree bits twin (index)
assert ob.packed is not None and np.array_equal(np.asarray(ob), D.odd_bits_np(np.arange(-500, 70001)))
ob = bp.map(D.odd_bits, list(range(-500, 7001)))                                                       # ... explicit items
assert np.array_equal(np.asarray(ob), D.odd_bits_np(np.arange(-500, 7001)))
r32, c32 = cref.pi_inside_range(2 ** 32 - 1000, 5003, 1)                                                # 2^32 crossing: scalar path
assert np.array_equal(np.asarray(pool.map(W.is_inside, range(2 ** 32 - 1000, 2 ** 32 + 4003))).view(np.uint8), r32)
assert np.array_equal(bp.map(W.is_inside, range(2 ** 32 - 1000, 2 ** 32 + 4003)).packed, np.packbits(r32, bitorder="little"))
recs = cref.payload_records(0, 700)
assert np.array_equal(np.asarray(pool.map(W.payload_map, recs)), cref.payload_map(0, recs))           # TMA dispatch + TMA gather
assert np.array_equal(np.asarray(pool.map(W.payload_map, recs, 7)), cref.payload_map(0, recs))        # odd units
big = fiber_b200.Pool(1, ring_bytes=16 << 20)
big.start_workers()
assert np.array_equal(raw(big, "payload_map_4k", 700, _abi.FBR_VIA_RING, 0, recs, 4096).view(np.uint32).reshape(700, 1024), cref.payload_map(0, recs))   # TMA bulk gather
assert np.array_equal(raw(big, "payload_map_4k", 700, _abi.FBR_SHUFFLE, 3, recs, 4096).view(np.uint32).reshape(700, 1024), cref.payload_map(0, recs))    # flat gather, shuffled
assert np.array_equal(np.asarray(pool.map(W.payload_checksum, recs)), cref.payload_checksum(recs))   # flat gather
xs, px, widths = B.parzen_example_inputs()
assert len(pool.starmap(W.parzen_estimation_f32, [(xs, px, w) for w in widths[:8]], 1)) == 8
assert pool.starmap(W.f2, [(x, x) for x in range(100)], 10) == [x * x for x in range(100)]
rp = fiber_b200.Pool(1, error_handling=True, ring_bytes=1 << 20)
assert rp.map(W.random_error_worker, range(5000)) == list(range(5000))
ep = fiber_b200.Pool(1)
assert [ep.apply(W.f, (i,)) for i in range(50)] == [i * i for i in range(50)]                          # express lane
q_in, q_out = fiber_b200.SimpleQueue(), fiber_b200.SimpleQueue()
p = fiber_b200.Process(target=W.worker, args=(q_in, q_out, 3), idle_timeout=20)
p.start()
q_in.put("work")
assert q_out.get(20) == 3
q_in.put("quit")
p.join(20)
print("sanitize target ok", p.exitcode)
