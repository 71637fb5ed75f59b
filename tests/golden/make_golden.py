#!/usr/bin/env python
"""Generate the committed golden vectors by running the REAL reference pool.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

What it does
------------
1. copies ``/root/reference/fiber`` to a fresh temp dir OUTSIDE the repo (reference sources are
   never copied into the repo) and flips the one transport constant ``fiber/socket.py:27``
   ``socket_lib = "nanomsg"`` -> ``"zmq"`` (``nnpy`` is not installable offline; pyzmq is present;
   SURVEY.md section 8(c)).  No numeric code lives in the transport, results are ``func(arg)`` placed
   by index (``fiber/pool.py:666-679``), so the vectors are transport independent;
2. re-executes itself with that dir on ``PYTHONPATH`` and pushes the deterministic workload bodies of
   ``oracle/bodies.py`` through ``fiber.Pool`` (``ZPool``, ``fiber/context.py:38-45``) and through
   ``fiber.Pool(error_handling=True)`` (``ResilientZPool``) with the local backend;
3. writes ``tests/golden/*.json``.

The JSON files are what the CPU and GPU parity tests compare against.
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REFERENCE = "/root/reference"


def _sha(arr_bytes):
    return hashlib.sha256(arr_bytes).hexdigest()


def child():
    import numpy as np
    import fiber  # the scratch copy of the reference
    import fiber.pool
    from oracle import bodies as B

    assert os.path.realpath(fiber.__file__).startswith(os.path.realpath(os.environ["FIBER_REF_SCRATCH"]))
    meta = {
        "reference_version": fiber.__version__,
        "transport": "zmq (fiber/socket.py:27 flipped from nanomsg; nnpy absent)",
        "numpy": np.__version__,
        "python": sys.version.split()[0],
    }

    pool = fiber.Pool(processes=4)
    assert type(pool) is fiber.pool.ZPool
    known = {"meta": meta}

    # ---- tests/test_pool.py known answers, restated with the same functions/values -------------
    known["map_basic"] = pool.map(B.square, [1, 2, 3])                                   # :86-91
    known["map_1000"] = pool.map(B.square, [i for i in range(1000)])                     # :93-104
    known["apply_async_42"] = pool.apply_async(B.square, (42,)).get()                    # :106-110
    known["apply_36"] = pool.apply(B.square, (36,))                                      # :112-113
    known["apply_kwds_36_y2"] = pool.apply(B.square_scale, (36,), {"y": 2})              # :115-116
    known["imap_100"] = list(pool.imap(B.square, [x for x in range(100)], 1))            # :121-125
    known["imap_unordered_100_sorted"] = sorted(pool.imap_unordered(B.square, [x for x in range(100)], 1))
    known["starmap_1arg_cs1"] = pool.starmap(B.square, [(x,) for x in range(100)], 1)    # :136-139
    known["starmap_async_1arg_cs1"] = pool.starmap_async(B.square, [(x,) for x in range(100)], 1).get()
    known["starmap_2arg_cs10"] = pool.starmap(B.mul2, [(x, x) for x in range(100)], 10)  # :148-151
    known["starmap_async_1arg_cs10"] = pool.starmap_async(B.square, [(x,) for x in range(100)], 10).get()
    # edge cases the reference semantics define (fiber/pool.py:1084-1087, 1169-1181, 659-679)
    known["map_empty"] = pool.map(B.square, [])
    known["map_negative_cs7"] = pool.map(B.square, list(range(-50, 51)), 7)
    known["map_cs_larger_than_n"] = pool.map(B.square, list(range(10)), 1000)
    known["map_range_step"] = pool.map(B.square, range(5, 500, 7))
    known["map_generator"] = pool.map(B.square, (i for i in range(33)))                  # no __len__ -> list()
    known["map_bigint"] = pool.map(B.square, [3037000499, -3037000499, 2 ** 31, -(2 ** 31)])
    # two maps in flight: results of the other seq are banked while waiting (fiber/pool.py:669-675)
    r1 = pool.map_async(B.square, range(0, 200))
    r2 = pool.map_async(B.square_scale, range(0, 100))
    known["two_inflight_second"] = r2.get()
    known["two_inflight_first"] = r1.get()

    # ---- pi_estimation, deterministic body (examples/pi_estimation.py:9-16) ---------------------
    pi = {"meta": meta, "key": list(B.PI_KEY)}
    n = 10 ** 6
    res = pool.map(B.pi_inside_det, range(0, n))
    assert all(type(r) is bool for r in res[:100])
    arr = np.array(res, dtype=np.uint8)
    pi["n"] = n
    pi["count"] = int(arr.sum())
    pi["pi"] = 4.0 * sum(res) / n
    pi["sha256_uint8"] = _sha(arr.tobytes())
    pi["head_256"] = arr[:256].tolist()
    pi["prefix_counts"] = {str(k): int(arr[:k].sum()) for k in (1, 10, 100, 1000, 10 ** 4, 10 ** 5)}
    # chunk-granular fingerprints (sha of every 65536-task block) so a mismatch can be localised
    pi["block_65536_counts"] = [int(arr[i:i + 65536].sum()) for i in range(0, n, 65536)]
    # indices that exercise the high counter word and negative (two's complement) task ids
    special = ([2 ** 32 - 2 + i for i in range(5)] + [2 ** 40 + i for i in range(5)]
               + [-3, -2, -1] + [2 ** 63 - 1, -(2 ** 63)])
    pi["special_args"] = special
    pi["special_results"] = [int(r) for r in pool.map(B.pi_inside_det, special)]
    # ranges the GPU's vectorised Philox path treats specially (16 consecutive arguments share one round-2
    # product while their indices share the high 32-bit word): 2^32 crossings, negative starts, backward and
    # large steps, lengths that leave partial vectors / partial bytes.  sha256 of the uint8 results and of the
    # same results packed 8 to a byte (LSB first; the layout of Pool(results="bits")).
    cases = [(0, 1, 1), (5, 7, 1), (3, 1001, 7), (-5000, 4097, 3), (2 ** 32 - 100, 333, 1), (2 ** 32 + 50, 97, -3),
             (10, 65537, 1), (2 ** 33 - 7, 4096 + 15, 1), (-3, 40, 1), (2 ** 40, 5000, 2 ** 31 + 1), (7, 130, -1)]
    pi["range_cases"] = []
    for start, m, step in cases:
        r = np.array(pool.map(B.pi_inside_det, range(start, start + m * step, step)), dtype=np.uint8)
        assert len(r) == m
        pi["range_cases"].append({"start": start, "n": m, "step": step, "count": int(r.sum()),
                                  "sha256_uint8": _sha(r.tobytes()),
                                  "sha256_bits_le": _sha(np.packbits(r, bitorder="little").tobytes())})
    pi["sha256_bits_le"] = _sha(np.packbits(arr, bitorder="little").tobytes())
    pi["uniforms_p0_hex"] = [v.hex() for v in B.pi_uniforms(0)]
    pi["uniforms_p12345_hex"] = [v.hex() for v in B.pi_uniforms(12345)]

    # ---- parzen_estimation (examples/parzen_estimation.py:22-40), via apply_async as the example --
    xs, px, widths = B.parzen_example_inputs()
    handles = [pool.apply_async(B.parzen_estimation, args=(xs, px, w)) for w in widths]
    results = [h.get() for h in handles]
    results.sort()
    star = pool.starmap(B.parzen_estimation, [(xs, px, w) for w in widths], 1)
    assert sorted(star) == results
    pz = {
        "meta": meta,
        "n_samples": len(xs),
        "n_widths": len(widths),
        "samples_sha256_f64": _sha(np.ascontiguousarray(xs, dtype=np.float64).tobytes()),
        "samples_head_hex": [[float(v).hex() for v in row] for row in xs[:4]],
        "widths_hex": [float(w).hex() for w in widths],
        "results_hex": [[float(h).hex(), float(d).hex()] for h, d in results],
        "results_repr": [[repr(float(h)), repr(float(d))] for h, d in results[:8]],
        "k_n": [B.parzen_count_np(xs, px, w) for w in widths],
    }
    # cross-check the vectorised oracle against what the reference pool returned
    for (h, d), w in zip(results, widths):
        assert B.parzen_estimation_np(xs, px, w) == (h, d), (h, d)

    # ---- synthetic 4 KB payload map pushed through starmap (BASELINE.json config 4) -------------
    nt = 64
    recs = [B.payload_record(t) for t in range(nt)]
    out = pool.starmap(B.payload_map, [(t, recs[t]) for t in range(nt)], 8)
    out_arr = np.array(out, dtype=np.uint32)
    cks = pool.starmap(B.payload_checksum, [(t, recs[t]) for t in range(nt)], 8)
    pl = {
        "meta": meta,
        "n_tasks": nt,
        "input_sha256_u32le": _sha(np.array(recs, dtype=np.uint32).tobytes()),
        "output_sha256_u32le": _sha(out_arr.tobytes()),
        "output_head": out_arr[:2, :8].tolist(),
        "output_tail": out_arr[-1, -8:].tolist(),
        "checksums": [int(c) for c in cks],
    }
    assert (B.payload_map_np(0, B.payload_records_np(0, nt)) == out_arr).all()

    pool.terminate()
    pool.join()

    # ---- ResilientZPool (fiber/pool.py:1425-1688; tests/test_pool.py:282-315 without the fault) --
    rpool = fiber.Pool(3, error_handling=True)
    assert type(rpool) is fiber.pool.ResilientZPool
    known["resilient_map_300_cs1"] = rpool.map(B.identity, [i for i in range(300)], chunksize=1)
    known["resilient_imap_unordered_300_sorted"] = sorted(rpool.imap_unordered(B.identity, [i for i in range(300)], chunksize=1))
    rpool.terminate()
    rpool.join()

    for name, obj in (("pool_known_answers", known), ("pi_inside_det", pi), ("parzen_102", pz), ("payload_map", pl)):
        with open(os.path.join(HERE, name + ".json"), "w") as fh:
            json.dump(obj, fh, indent=1, sort_keys=True)
            fh.write("\n")
        print("wrote", name + ".json")


def main():
    if not os.path.isdir(REFERENCE):
        sys.exit("make_golden.py needs /root/reference (build container only)")
    scratch = tempfile.mkdtemp(prefix="fiber_ref_")
    try:
        shutil.copytree(os.path.join(REFERENCE, "fiber"), os.path.join(scratch, "fiber"))
        sock = os.path.join(scratch, "fiber", "socket.py")
        src = open(sock).read()
        assert 'socket_lib = "nanomsg"' in src
        open(sock, "w").write(src.replace('socket_lib = "nanomsg"', 'socket_lib = "zmq"', 1))
        env = dict(os.environ)
        env["PYTHONPATH"] = os.pathsep.join([scratch, REPO, env.get("PYTHONPATH", "")])
        env["FIBER_REF_SCRATCH"] = scratch
        subprocess.check_call([sys.executable, os.path.abspath(__file__), "--child"], env=env, cwd=scratch)
    finally:
        shutil.rmtree(scratch, ignore_errors=True)


if __name__ == "__main__":
    if "--child" in sys.argv:
        child()
    else:
        main()
