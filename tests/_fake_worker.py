"""A stand-in worker process for the CPU tests of fiber_b200.procpool (host logic only: block dispatch, placement,
re-queueing, respawn).  It speaks the worker protocol and "computes" identity / squares with NumPy -- test
infrastructure, never used by the product (the real worker is ``procpool.gpu_worker_main``, which runs every block
through the C ABI on a GPU)."""
import os
import pickle
import sys

import numpy as np


def fake_worker_main(device, conn, results, sys_path):
    for p in sys_path:
        if p not in sys.path:
            sys.path.insert(0, p)
    from fiber_b200.procpool import SharedSegment
    segs = {}
    conn.send(("ready", os.getpid()))
    while True:
        msg = conn.recv()
        if msg is None:
            break
        _, job, blk, body, kind, chunksize, payload, shm_name, off, attempt, module = msg
        items = range(*payload[1]) if payload[0] == "range" else pickle.loads(payload[1])
        xs = np.asarray(list(items) if kind != "starmap" else [a[0] for a in items], dtype=np.int64)
        if body == "trap_identity_i64" and attempt == 0 and ((xs & 0xFFFFF) == 0xDEAD).any():
            os._exit(3)                                  # the worker process dies mid-block, without a word
        if body == "fault_report_i64" and attempt == 0 and (xs == 7).any():
            conn.send(("dead", job, blk, "simulated sticky CUDA error"))
            os._exit(3)
        if body == "square_i64" and (np.abs(xs) > 3037000499).any():
            conn.send(("error", job, blk, "OverflowError", "square_i64: result does not fit int64"))
            continue
        out = xs * xs if body == "square_i64" else xs
        seg = segs.get(shm_name) or segs.setdefault(shm_name, SharedSegment(shm_name))
        raw = np.ascontiguousarray(out).view(np.uint8)
        seg.array[off:off + raw.nbytes] = raw
        conn.send(("done", job, blk, int(out.sum())))
    os._exit(0)
