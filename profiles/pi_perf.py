"""Event-timed pi dispatch / gather kernels only (quick A/B of FBR_DISPATCH_OCC / FBR_UNIT_TASKS)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
eng = bench.RawEngine(0, 160 << 20)
out = eng.dalloc(bench.PI_TASKS)
cnt = None
for i in range(steps + 3):
    if i == 3:
        eng.stats(reset=True)
    cnt = eng.wait(eng.submit("pi_inside_det", bench.PI_TASKS, out))[0]
st = eng.stats()
print("occ=%s unit=%s  dispatch %.4f ms  gather %.4f ms  count %d" % (
    os.environ.get("FBR_DISPATCH_OCC", "-"), os.environ.get("FBR_UNIT_TASKS", "-"),
    st["dispatch_ms"] / st["dispatch_launches"], st["gather_ms"] / st["gather_launches"], cnt), flush=True)
eng.dfree(out)
eng.close()
