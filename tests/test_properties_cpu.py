"""CPU: property tests (hypothesis) of the host-side logic: record codec, block partition, claim-unit
planning, int64 argument encoding."""
import ctypes

import numpy as np
from hypothesis import given, settings, strategies as st

from fiber_b200 import _abi, registry, shard
from fiber_b200.queues import decode, encode

scalars = st.one_of(
    st.none(), st.integers(-2 ** 63, 2 ** 63 - 1), st.floats(allow_nan=False),
    st.binary(max_size=56), st.text(max_size=14))


@settings(max_examples=300, deadline=None)
@given(scalars)
def test_record_codec_round_trip(v):
    r = encode(v)
    assert ctypes.sizeof(r) == 64 and r.len <= 56
    out = decode(r)
    assert out == v and type(out) is type(v)


@settings(max_examples=200, deadline=None)
@given(st.integers(0, 10 ** 9), st.integers(1, 16), st.sampled_from([1, 16, 32, 4096]))
def test_block_partition_invariants(n, world, align):
    blocks = shard.blocks(n, world, align)
    assert blocks[0][0] == 0 and blocks[-1][1] == n
    assert all(lo <= hi for lo, hi in blocks)
    assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
    sizes = [hi - lo for lo, hi in blocks]
    assert max(sizes) - min(sizes) < 2 * align


@settings(max_examples=300, deadline=None)
@given(st.sampled_from(["pi_inside_det", "pi_inside_bits8", "square_i64", "mul2_i64", "payload_map_4k", "payload_checksum_4k", "parzen_f64"]),
       st.integers(1, 10 ** 9), st.integers(0, 100000), st.sampled_from([64 << 10, 1 << 20, 256 << 20, 4 << 30]),
       st.integers(1, 8))
def test_claim_unit_invariants(body, n, cs, ring, nw):
    spec = registry.spec(body)
    lib = _abi.load()
    total, prev_end = 0, 0
    for w in range(nw):
        p = _abi.Plan()
        _abi.check(lib.fbr_plan_query(spec.func_id, n, cs, ring, nw, w, 148, ctypes.byref(p)))
        assert p.block_first == prev_end                      # contiguous blocks
        prev_end = p.block_first + p.block_count
        total += p.block_count
        if p.block_count:
            assert p.unit_tasks >= 1 and p.slot_stride % 16 == 0
            assert p.slot_stride >= p.unit_tasks * spec.result_bytes
            assert p.unit_tasks * max(spec.result_bytes, spec.arg_bytes) <= max(ring, max(spec.result_bytes, spec.arg_bytes))
            if spec.result_bytes < 16 and p.unit_tasks > 1:
                assert (p.unit_tasks * spec.result_bytes) % 16 == 0   # full slots are 16 B aligned
            assert p.n_units == -(-p.block_count // p.unit_tasks)
    assert total == n


@settings(max_examples=100, deadline=None)
@given(st.lists(st.integers(-2 ** 63, 2 ** 63 - 1), max_size=50))
def test_int64_encoders_preserve_values(xs):
    s = registry.spec("square_i64")
    e = s.encode_map(xs)
    assert e.n == len(xs) and (e.args.tolist() if e.n else []) == xs
    assert s.pack_apply((xs[0],), {}) == np.int64(xs[0]).tobytes() if xs else True
    star = s.encode_starmap([(x,) for x in xs])
    assert (star.args.tolist() if star.n else []) == xs
