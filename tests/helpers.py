"""Shared test helpers: seeded input patterns modelled on the reference's own unit tests
(test/SadTest.cc:60,108: REF_MAX, SRC_MAX, RANDOM, UNALIGN) and thin ctypes callers for the
checkers (oracle.port = our restatement, oracle.ref = unmodified reference objects)."""
import ctypes as ct

import numpy as np

SEED = 20260923


def rng(extra=0):
    return np.random.default_rng(SEED + extra)


def sad_loop_call(lib, fname, src, src_off, src_stride, ref, ref_off, ref_stride, bh, bw, ref_step, skip, sa_w, sa_h,
                  x_init=0, y_init=0):
    f = getattr(lib, fname)
    f.restype = None
    best = ct.c_uint64(0)
    xs = ct.c_int16(x_init)
    ys = ct.c_int16(y_init)
    f(ct.c_void_p(src.ctypes.data + src_off), ct.c_uint32(src_stride), ct.c_void_p(ref.ctypes.data + ref_off),
      ct.c_uint32(ref_stride), ct.c_uint32(bh), ct.c_uint32(bw), ct.byref(best), ct.byref(xs), ct.byref(ys),
      ct.c_uint32(ref_step), ct.c_uint8(skip), ct.c_int16(sa_w), ct.c_int16(sa_h))
    return int(best.value), int(xs.value), int(ys.value)


def sad_pattern(pattern, r, n_src, n_ref):
    if pattern == "REF_MAX":
        return np.zeros(n_src, np.uint8), np.full(n_ref, 255, np.uint8)
    if pattern == "SRC_MAX":
        return np.full(n_src, 255, np.uint8), np.zeros(n_ref, np.uint8)
    if pattern == "FLAT":  # every position ties -> pins the first-minimum rule
        return np.full(n_src, 77, np.uint8), np.full(n_ref, 80, np.uint8)
    return r.integers(0, 256, n_src, dtype=np.uint8), r.integers(0, 256, n_ref, dtype=np.uint8)
