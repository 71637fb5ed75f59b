"""Small, fast exercise of every kernel for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import fiber_b200  # noqa: E402
from examples import workloads as W  # noqa: E402
from oracle import bodies as B, cref  # noqa: E402

pool = fiber_b200.Pool(1, ring_bytes=1 << 20, express=False)
assert pool.map(W.f, range(1000)) == [i * i for i in range(1000)]
assert pool.map(W.f, list(range(-50, 51)), 7) == [i * i for i in range(-50, 51)]
n = 300_017
res = pool.map(W.is_inside, range(n))
ref, count = cref.pi_inside_range(0, n)
assert res.sum() == count and np.array_equal(np.asarray(res).view(np.uint8), ref)
bp = fiber_b200.Pool(1, results="bits", ring_bytes=1 << 20)                                          # pi_inside_bits8
rb = bp.map(W.is_inside, range(n))
assert rb.sum() == count and np.array_equal(rb.packed, np.packbits(ref, bitorder="little"))
r32, c32 = cref.pi_inside_range(2 ** 32 - 1000, 5003, 1)                                                # 2^32 crossing: scalar path
assert np.array_equal(np.asarray(pool.map(W.is_inside, range(2 ** 32 - 1000, 2 ** 32 + 4003))).view(np.uint8), r32)
assert np.array_equal(bp.map(W.is_inside, range(2 ** 32 - 1000, 2 ** 32 + 4003)).packed, np.packbits(r32, bitorder="little"))
recs = cref.payload_records(0, 700)
assert np.array_equal(np.asarray(pool.map(W.payload_map, recs)), cref.payload_map(0, recs))           # TMA dispatch + TMA gather
assert np.array_equal(np.asarray(pool.map(W.payload_map, recs, 7)), cref.payload_map(0, recs))        # odd units
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
