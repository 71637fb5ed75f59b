"""CPU: the N>1 path (rank blocks + torch.distributed exchange) with world_size 2 over gloo."""
import json
import os
import socket
import subprocess
import sys

import pytest

from fiber_b200 import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_block_partition_properties():
    for n, world, align in [(0, 2, 1), (1, 2, 1), (10 ** 6, 8, 4096), (1000003, 3, 32), (5, 8, 1), (10 ** 8, 8, 4096)]:
        bl = shard.blocks(n, world, align)
        assert bl[0][0] == 0 and bl[-1][1] == n
        assert all(bl[i][1] == bl[i + 1][0] for i in range(world - 1))          # contiguous, disjoint, complete
        sizes = [hi - lo for lo, hi in bl]
        assert all(lo % align == 0 for lo, _ in bl if lo < n)                   # claim-unit aligned starts
        assert max(sizes) - min(sizes) < 2 * align                               # balanced (tail unit is partial)
    assert shard.blocks(10 ** 6, 8, 1)[3] == (375000, 500000)                   # 125000 tasks / GPU (config 4)
    with pytest.raises(ValueError):
        shard.block_of(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_world_size_2_gloo():
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), PYTHONPATH=ROOT)
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_gloo_worker.py"), "300017"],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=240) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    line = json.loads(outs[0][0].strip().splitlines()[-1])
    assert line["ok"] and line["total"] == line["ref"] and line["tmax"] == 2.0
    assert line["block"] == [0, 151552]                                           # 37 units of 4096 on rank 0
