"""CPU: pin the oracle (oracle/bodies.py, oracle/fbr_oracle.c, oracle/zpool_port.py) against
(a) published known-answer vectors and (b) the golden vectors produced by the REAL reference pool
(tests/golden/make_golden.py)."""
import hashlib

import numpy as np
import pytest

from oracle import bodies as B
from oracle import cref
from oracle import zpool_port as Z


# ---- published KATs --------------------------------------------------------------------------------
def test_philox_random123_kat():
    # Random123 kat_vectors: philox4x32 10
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, want in kat:
        assert B.philox4x32_10(ctr, key) == want
        c = np.array(ctr, dtype=np.uint32)
        k = np.array(key, dtype=np.uint32)
        o = np.zeros(4, dtype=np.uint32)
        cref.lib().orc_philox4x32_10(c.ctypes.data, k.ctypes.data, o.ctypes.data)
        assert tuple(int(v) for v in o) == want
        got = B.philox4x32_10_np(*[np.array([v], dtype=np.uint64) for v in ctr], *key)
        assert tuple(int(v[0]) for v in got) == want


def test_splitmix64_kat():
    # first outputs of SplitMix64 seeded with 0 (java.util.SplittableRandom / xoshiro seeding)
    assert B.splitmix64(0) == 0xE220A8397B1DCDAF
    assert B.splitmix64(0x9E3779B97F4A7C15) == 0x6E789E6AA1B965F4
    assert cref.lib().orc_splitmix64(0) == 0xE220A8397B1DCDAF
    assert int(B.splitmix64_np(np.array([0], dtype=np.uint64))[0]) == 0xE220A8397B1DCDAF


# ---- golden vectors from the real reference pool -----------------------------------------------------
def test_pi_against_reference_pool(golden):
    g = golden("pi_inside_det")
    n = g["n"]
    arr, count = cref.pi_inside_range(0, n)
    assert count == g["count"]
    assert hashlib.sha256(arr.tobytes()).hexdigest() == g["sha256_uint8"]
    assert arr[:256].tolist() == g["head_256"]
    assert [int(arr[i:i + 65536].sum()) for i in range(0, n, 65536)] == g["block_65536_counts"]
    assert np.array_equal(B.pi_inside_det_np(0, 200000), arr[:200000])
    assert [int(B.pi_inside_det(p)) for p in g["special_args"]] == g["special_results"]
    assert [cref.lib().orc_pi_inside_one(p) for p in g["special_args"]] == g["special_results"]
    assert [v.hex() for v in B.pi_uniforms(0)] == g["uniforms_p0_hex"]
    assert [v.hex() for v in B.pi_uniforms(12345)] == g["uniforms_p12345_hex"]
    for k, v in g["prefix_counts"].items():
        assert int(arr[:int(k)].sum()) == v
    # strided / negative ranges agree between the three restatements
    a, _ = cref.pi_inside_range(-17, 1000, step=3)
    assert np.array_equal(a, B.pi_inside_det_np(-17, -17 + 3000, 3))


def test_parzen_against_reference_pool(golden):
    g = golden("parzen_102")
    xs, px, widths = B.parzen_example_inputs()
    if hashlib.sha256(np.ascontiguousarray(xs).tobytes()).hexdigest() != g["samples_sha256_f64"]:
        pytest.skip("numpy RNG stream differs from the golden run (numpy %s)" % g["meta"]["numpy"])
    assert [float(w).hex() for w in widths] == g["widths_hex"]
    want = [(float.fromhex(h), float.fromhex(d)) for h, d in g["results_hex"]]
    got = [B.parzen_estimation_np(xs, px, w) for w in widths]
    assert [(float(h), float(d)) for h, d in got] == want
    assert [cref.parzen_count(xs, px, w) for w in widths] == g["k_n"]
    for w, k in list(zip(widths, g["k_n"]))[::17]:
        assert cref.lib().orc_parzen_density(k, len(xs), float(w), 1) == (k / len(xs)) / w
    # pure-Python restatement (slow): three widths
    for i in (0, 50, 101):
        h, d = B.parzen_estimation(xs, px, widths[i])
        assert (float(h), float(d)) == want[i]
    # fp32 restatements agree with each other and stay within the boundary-sample bound
    for w, k in list(zip(widths, g["k_n"]))[::9]:
        k32 = cref.parzen_count(xs, px, w, np.float32)
        assert k32 == B.parzen_count_np(xs, px, w, np.float32)
        assert abs(k32 - k) <= B.parzen_boundary_count(xs, px, w)


def test_payload_against_reference_pool(golden):
    g = golden("payload_map")
    nt = g["n_tasks"]
    recs = cref.payload_records(0, nt)
    assert np.array_equal(recs, B.payload_records_np(0, nt))
    assert recs[3].tolist() == B.payload_record(3)
    assert hashlib.sha256(recs.tobytes()).hexdigest() == g["input_sha256_u32le"]
    out = cref.payload_map(0, recs)
    assert np.array_equal(out, B.payload_map_np(0, recs))
    assert hashlib.sha256(out.tobytes()).hexdigest() == g["output_sha256_u32le"]
    assert out[:2, :8].tolist() == g["output_head"] and out[-1, -8:].tolist() == g["output_tail"]
    assert cref.payload_checksum(recs).tolist() == g["checksums"] == B.payload_checksum_np(recs).tolist()
    assert B.payload_map(5, B.payload_record(5)) == out[5].tolist()


def test_known_answers_functions(golden):
    g = golden("pool_known_answers")
    assert [B.square(x) for x in [1, 2, 3]] == g["map_basic"]
    assert [B.square(i) for i in range(1000)] == g["map_1000"]
    assert B.square(42) == g["apply_async_42"] and B.square_scale(36, y=2) == g["apply_kwds_36_y2"]
    assert [B.mul2(x, x) for x in range(100)] == g["starmap_2arg_cs10"]
    assert [B.square(x) for x in range(5, 500, 7)] == g["map_range_step"]
    x = np.array([3037000499, -3037000499, 2 ** 31, -(2 ** 31)], dtype=np.int64)
    out = np.zeros(4, dtype=np.int64)
    ovf = np.zeros(1, dtype=np.int32)
    cref.lib().orc_square_i64(x.ctypes.data, 4, out.ctypes.data, ovf.ctypes.data)
    assert out.tolist() == g["map_bigint"] and ovf[0] == 0


# ---- the pool port behaves like the reference pool ------------------------------------------------------
@pytest.fixture(scope="module")
def port_pool():
    p = Z.PortPool(2)
    yield p
    p.terminate()
    p.join()


def test_port_pool_matches_reference_answers(port_pool, golden):
    g = golden("pool_known_answers")
    p = port_pool
    assert p.map(B.square, [1, 2, 3]) == g["map_basic"]
    assert p.map(B.square, [i for i in range(1000)]) == g["map_1000"]
    assert p.apply_async(B.square, (42,)).get() == g["apply_async_42"]
    assert p.apply(B.square_scale, (36,), {"y": 2}) == g["apply_kwds_36_y2"]
    assert list(p.imap(B.square, [x for x in range(100)], 1)) == g["imap_100"]
    assert sorted(p.imap_unordered(B.square, [x for x in range(100)], 1)) == g["imap_unordered_100_sorted"]
    assert p.starmap(B.square, [(x,) for x in range(100)], 1) == g["starmap_1arg_cs1"]
    assert p.starmap(B.mul2, [(x, x) for x in range(100)], 10) == g["starmap_2arg_cs10"]
    assert p.map(B.square, []) == g["map_empty"]
    assert p.map(B.square, list(range(-50, 51)), 7) == g["map_negative_cs7"]
    assert p.map(B.square, (i for i in range(33))) == g["map_generator"]
    r1 = p.map_async(B.square, range(0, 200))
    r2 = p.map_async(B.square_scale, range(0, 100))
    assert r2.get() == g["two_inflight_second"] and r1.get() == g["two_inflight_first"]
    n = 20000
    res = p.map(cref.pi_inside_det_c, range(n))
    assert all(type(r) is bool for r in res[:10])
    assert np.array_equal(np.array(res, dtype=np.uint8), cref.pi_inside_range(0, n)[0])


def test_port_pool_close_semantics():
    p = Z.PortPool(1)
    assert p.map(B.square, [1, 2, 3]) == [1, 4, 9]
    p.close()
    with pytest.raises(ValueError):
        p.map(B.square, [1, 2, 3])
    p.join()
    with pytest.raises(NotImplementedError):
        Z.PortPool(1).map_async(B.square, [1], error_callback=print)


def test_resilient_port_pool(golden):
    g = golden("pool_known_answers")
    p = Z.ResilientPortPool(2)
    assert p.map(B.identity, [i for i in range(300)], chunksize=1) == g["resilient_map_300_cs1"]
    assert sorted(p.imap_unordered(B.identity, [i for i in range(300)], chunksize=1)) == g["resilient_imap_unordered_300_sorted"]
    p.terminate()
    p.join()


def test_chunk_plan_and_placement():
    assert Z.chunk_plan(100) == [(0, 32), (32, 32), (64, 32), (96, 4)]       # default chunksize 32
    assert Z.chunk_plan(5, 10) == [(0, 5)] and Z.chunk_plan(0) == []
    assert Z.n_jobs(9, 8) == 2 and Z.n_jobs(4, 1) == 4
    # Inventory places by index whatever the arrival order
    rng = np.random.default_rng(7)
    n, chunk, rb = 1000, 32, 8
    vals = np.arange(n, dtype=np.int64) * 3
    order = rng.permutation((n + chunk - 1) // chunk)
    ring = np.concatenate([vals[c * chunk:(c + 1) * chunk] for c in order]).view(np.uint8)
    placed = cref.place_by_index(ring, order, chunk, n, rb)
    assert np.array_equal(placed.view(np.int64), vals)
    msgs = [(1, (i // chunk) * chunk, i, int(vals[i])) for c in order for i in range(c * chunk, min(n, (c + 1) * chunk))]
    it = iter(msgs)
    inv = Z.Inventory(lambda: next(it))
    assert inv.add(n) == 1
    assert inv.get(1) == vals.tolist()


def test_pi_bits8_restatement_matches_the_byte_body(golden):
    """pi_inside_bits8 is pi_inside_det in another result layout: pinned by the same golden vector."""
    g = golden("pi_inside_det")
    head = np.array(g["head_256"], dtype=np.uint8)
    assert [B.pi_inside_bits8(j) for j in range(32)] == np.packbits(head, bitorder="little").tolist()
    arr, _ = cref.pi_inside_range(-17, 1003, step=3)
    packed = B.pi_inside_bits_np(-17, 1003, 3)
    assert packed.nbytes == 126 and np.array_equal(np.unpackbits(packed, bitorder="little")[:1003], arr)
    assert [B.pi_inside_bits8(j, -17, 3) for j in range(125)] == packed[:125].tolist()     # full bytes
    assert packed[125] == B.pi_inside_bits8(125, -17, 3) & 0b111                             # 1003 = 125*8 + 3


def test_pi_range_cases_from_the_reference_pool(golden):
    """Ranges crossing 2^32, negative starts, backward / large steps: the C oracle and the NumPy restatement
    against what the real reference pool returned (tests/golden/make_golden.py), bytes and packed bits."""
    import hashlib
    g = golden("pi_inside_det")
    assert len(g["range_cases"]) >= 11
    for c in g["range_cases"]:
        arr, count = cref.pi_inside_range(c["start"], c["n"], c["step"])
        assert count == c["count"] and hashlib.sha256(arr.tobytes()).hexdigest() == c["sha256_uint8"], c
        assert hashlib.sha256(np.packbits(arr, bitorder="little").tobytes()).hexdigest() == c["sha256_bits_le"], c
        if c["n"] <= 5000:
            assert np.array_equal(B.pi_inside_det_np(c["start"], c["start"] + c["n"] * c["step"], c["step"]), arr), c
            assert np.array_equal(B.pi_inside_bits_np(c["start"], c["n"], c["step"]), np.packbits(arr, bitorder="little")), c
    full, _ = cref.pi_inside_range(0, g["n"])
    assert hashlib.sha256(np.packbits(full, bitorder="little").tobytes()).hexdigest() == g["sha256_bits_le"]
