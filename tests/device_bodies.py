"""Device bodies defined OUTSIDE libfiber_b200 (test infrastructure for fbr_register_body).

The CUDA source below is compiled by ``fiber_b200.device_body(name, source=...)`` into a body module under
``fiber_b200/_lib/bodies/`` and registered with the engine at import time; the Python functions are the
definitions the GPU results are compared against.  (The reference can map any callable,
fiber/pool.py:961; this is the route for a callable whose device code is not compiled into the library.)
"""
import numpy as np

import fiber_b200

COLLATZ_SRC = r'''
#include "fiber_b200_body.cuh"

// number of Collatz steps from x down to 1 (capped at 1000); x < 1 is a bad argument
struct CollatzSteps {
    using Arg = int64_t; using Res = int64_t;
    static constexpr bool kIndexArg = true;
    static constexpr bool kVecIndex = false;
    static constexpr bool kCanFault = false;
    __device__ static __forceinline__ Res run(const Arg& a, uint64_t gidx, const fbr::ErrSink& es, uint32_t) {
        if (a < 1) { es.report(fbr::TASK_BADARG, gidx); return 0; }
        uint64_t x = (uint64_t)a;
        int64_t steps = 0;
        while (x != 1 && steps < 1000) { x = (x & 1) ? 3 * x + 1 : x >> 1; ++steps; }
        return steps;
    }
};
FBR_EXPORT_THREAD_BODY(CollatzSteps, "collatz_steps", fbr_body_entry, FBR_RES_I64, FBR_BODY_INDEX_ARG | FBR_BODY_SUMMABLE)
'''

ODD_BITS_SRC = r'''
#include "fiber_b200_body.cuh"

// bool result: does x have an odd number of set bits (two's complement, 64 bits)?
struct OddBits {
    using Arg = int64_t; using Res = uint8_t;
    static constexpr bool kIndexArg = true;
    static constexpr bool kVecIndex = false;
    static constexpr bool kCanFault = false;
    __device__ static __forceinline__ Res run(const Arg& a, uint64_t, const fbr::ErrSink&, uint32_t) {
        return (uint8_t)(__popcll((unsigned long long)a) & 1);
    }
};
FBR_EXPORT_THREAD_BODY(OddBits, "odd_bits", odd_bits_entry, FBR_RES_BOOL, FBR_BODY_INDEX_ARG | FBR_BODY_SUMMABLE)
// ... and its bit-packed twin: 8 items (range() indices or int64 records) per result byte
FBR_EXPORT_BOOL_BODY_BITS(OddBits, "odd_bits_bits8", odd_bits_bits_entry, FBR_BODY_INDEX_ARG | FBR_BODY_SUMMABLE)
'''


@fiber_b200.device_body("collatz_steps", source=COLLATZ_SRC)
def collatz_steps(x):
    if x < 1:
        raise ValueError("collatz_steps: bad argument")
    steps = 0
    while x != 1 and steps < 1000:
        x = 3 * x + 1 if x & 1 else x >> 1
        steps += 1
    return steps


@fiber_b200.device_body("odd_bits", source=ODD_BITS_SRC, entry="odd_bits_entry", bits_entry="odd_bits_bits_entry")
def odd_bits(x):
    return bin(x & (2 ** 64 - 1)).count("1") % 2 == 1


def collatz_steps_np(xs):
    """Vectorised restatement of ``collatz_steps`` (uint64 wrap-around like the device body)."""
    x = np.asarray(xs, dtype=np.uint64).copy()
    steps = np.zeros(x.shape, dtype=np.int64)
    with np.errstate(over="ignore"):
        for _ in range(1000):
            live = x != 1
            if not live.any():
                break
            odd = live & ((x & np.uint64(1)) == 1)
            even = live & ~odd
            x[odd] = x[odd] * np.uint64(3) + np.uint64(1)
            x[even] >>= np.uint64(1)
            steps += live
    return steps


def odd_bits_np(xs):
    x = np.asarray(xs, dtype=np.int64).view(np.uint64)
    return (np.unpackbits(x.view(np.uint8).reshape(-1, 8), axis=1).sum(axis=1) & 1).astype(bool)
