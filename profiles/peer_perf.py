"""Event-timed peer-memory map (args/out on GPU 0, in-process pool over N GPUs)."""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fiber_b200 import _abi, registry
n_gpus = int(sys.argv[1]); n_total = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
lib = _abi.load()
ids = (ctypes.c_int * n_gpus)(*range(n_gpus))
h = ctypes.c_void_p()
ring = int(os.environ.get("PEER_RING_MB", "256")) << 20      # staging halves: the largest wave of the NVLink pipeline
_abi.check(lib.fbr_pool_create(n_gpus, ids, ring, _abi.FBR_POOL_TIMING, ctypes.byref(h)))
din, dout = ctypes.c_void_p(), ctypes.c_void_p()
_abi.check(lib.fbr_device_alloc(h, 0, n_total * 4096, ctypes.byref(din)))
_abi.check(lib.fbr_device_alloc(h, 0, n_total * 4096, ctypes.byref(dout)))
_abi.check(lib.fbr_payload_fill_device(h, 0, din, 0, n_total))
d = _abi.MapDesc(); d.func_id = registry.spec("payload_map_4k").func_id
d.flags = _abi.FBR_ARGS_DEVICE | _abi.FBR_OUT_DEVICE
d.n_tasks, d.arg_stride, d.args, d.out = n_total, 4096, din.value, dout.value
res = _abi.Result()
def step():
    seq = ctypes.c_uint64()
    _abi.check(lib.fbr_map_submit(h, ctypes.byref(d), ctypes.byref(seq)))
    _abi.check(lib.fbr_result_wait(h, seq.value, -1, ctypes.byref(res)))
    _abi.check(lib.fbr_result_release(h, seq.value))
for _ in range(3): step()
lib.fbr_pool_stats_reset(h)
t0 = time.perf_counter()
K = 5
for _ in range(K): step()
dt = (time.perf_counter() - t0) / K
s = _abi.Stats(); lib.fbr_pool_stats(h, ctypes.byref(s))
print("gpus %d: %.3f ms/step; sum over workers per step: dispatch %.3f ms, gather %.3f ms; launches/step %d" % (
    n_gpus, dt * 1e3, s.dispatch_ms / K, s.gather_ms / K, s.dispatch_launches // K), flush=True)
lib.fbr_pool_destroy(h)
