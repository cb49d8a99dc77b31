"""Quantizer test inputs + ctypes callers for the checkers.  Table construction follows the
reference's svt_aom_invert_quant / svt_av1_build_quantizer arithmetic (inv_transforms.c:3369-3407,
full_loop.c) closely enough to give realistic value ranges; parity does not depend on it."""
import ctypes as ct

import numpy as np


def tables_for(d_dc, d_ac, zbin_factor=84):
    def inv(d):
        l = int(np.floor(np.log2(d)))
        m = 1 + (1 << (16 + l)) // d
        v = (m - (1 << 16)) & 0xffff
        sh = (1 << (16 - l)) & 0xffff
        return np.array([v], np.uint16).astype(np.int16)[0], np.array([sh], np.uint16).astype(np.int16)[0]
    q0, s0 = inv(d_dc)
    q1, s1 = inv(d_ac)
    t = {
        "zbin": np.array([(zbin_factor * d_dc + 64) >> 7, (zbin_factor * d_ac + 64) >> 7], np.int16),
        "round": np.array([(48 * d_dc) >> 7, (48 * d_ac) >> 7], np.int16),
        "quant": np.array([q0, q1], np.int16),
        "quant_shift": np.array([s0, s1], np.int16),
        "dequant": np.array([d_dc, d_ac], np.int16),
    }
    fp = dict(t)
    fp["quant"] = np.array([(1 << 16) // d_dc, (1 << 16) // d_ac], np.uint16).astype(np.int16)
    fp["round"] = np.array([(64 * d_dc) >> 7, (64 * d_ac) >> 7], np.int16)
    return t, fp


def coeffs(r, n, bd, kind):
    lim = 1 << (bd + 7)
    if kind == "zero":
        return np.zeros(n, np.int32)
    if kind == "dc":
        c = np.zeros(n, np.int32)
        c[0] = -lim // 3
        return c
    if kind == "large_neg":
        return np.full(n, -lim + 1, np.int32)
    if kind == "small":  # many inside the dead zone
        return r.integers(-40, 41, n).astype(np.int32)
    c = r.integers(-lim, lim, n).astype(np.int32)
    c[r.random(n) < 0.5] //= 64
    return c


def scan_for(r, n, kind):
    if kind == "identity":
        return np.arange(n, dtype=np.int16)
    return r.permutation(n).astype(np.int16)


def call_ref(lib, fname, coeff, t, scan, extra):
    n = coeff.size
    q = np.full(n, 0x5a5a5a5a, np.int32)
    dq = np.full(n, 0x5a5a5a5a, np.int32)
    eob = ct.c_uint16(0xffff)
    f = getattr(lib, fname)
    f.restype = None
    P = lambda a: ct.c_void_p(a.ctypes.data) if a is not None else ct.c_void_p(0)  # noqa: E731
    f(P(coeff), ct.c_ssize_t(n), P(t["zbin"]), P(t["round"]), P(t["quant"]), P(t["quant_shift"]), P(q), P(dq),
      P(t["dequant"]), ct.byref(eob), P(scan), P(scan), *extra)
    return q, dq, int(eob.value)


def call_port(lib, mode, coeff, t, scan, qm, iqm, ls):
    n = coeff.size
    q = np.full(n, 0x5a5a5a5a, np.int32)
    dq = np.full(n, 0x5a5a5a5a, np.int32)
    eob = ct.c_uint16(0xffff)
    P = lambda a: ct.c_void_p(a.ctypes.data) if a is not None else ct.c_void_p(0)  # noqa: E731
    if mode in ("b_lbd", "b_hbd"):
        f = getattr(lib, "port_quantize_" + mode)
        f.restype = None
        f(P(coeff), ct.c_ssize_t(n), P(t["zbin"]), P(t["round"]), P(t["quant"]), P(t["quant_shift"]), P(q), P(dq),
          P(t["dequant"]), ct.byref(eob), P(scan), P(qm), P(iqm), ls)
    else:
        f = getattr(lib, "port_quantize_" + mode)
        f.restype = None
        f(P(coeff), ct.c_ssize_t(n), P(t["round"]), P(t["quant"]), P(q), P(dq), P(t["dequant"]), ct.byref(eob), P(scan),
          P(qm), P(iqm), ls)
    return q, dq, int(eob.value)


# (b200/reference name suffix, port mode, takes qm, fixed log_scale or None, uses fp tables)
VARIANTS = [
    ("aom_quantize_b", "svt_aom_quantize_b_c_ii", "b_lbd", True, None, False),
    ("aom_highbd_quantize_b", "svt_aom_highbd_quantize_b_c", "b_hbd", True, None, False),
    ("av1_quantize_b_qm", "svt_aom_quantize_b_c_ii", "b_lbd", True, None, False),
    ("av1_highbd_quantize_b_qm", "svt_aom_highbd_quantize_b_c", "b_hbd", True, None, False),
    ("av1_quantize_fp", "svt_av1_quantize_fp_c", "fp_lbd", False, 0, True),
    ("av1_quantize_fp_32x32", "svt_av1_quantize_fp_32x32_c", "fp_lbd", False, 1, True),
    ("av1_quantize_fp_64x64", "svt_av1_quantize_fp_64x64_c", "fp_lbd", False, 2, True),
    ("av1_quantize_fp_qm", "svt_av1_quantize_fp_qm_c", "fp_lbd", True, None, True),
    ("av1_highbd_quantize_fp", "svt_av1_highbd_quantize_fp_c", "fp_hbd", False, None, True),
    ("av1_highbd_quantize_fp_qm", "svt_av1_highbd_quantize_fp_qm_c", "fp_hbd", True, None, True),
]


def cases(r):
    """yield (variant, coeff, tables, scan, qm, iqm, log_scale)"""
    for v in VARIANTS:
        name, refname, mode, has_qm, fixed_ls, use_fp = v
        for (d_dc, d_ac) in [(4, 4), (8, 9), (52, 61), (140, 163), (1336, 1828), (5347, 21387)]:
            tb, tfp = tables_for(d_dc, d_ac)
            t = tfp if use_fp else tb
            for n in (16, 64, 256, 1024):
                for kind in ("random", "small", "dc", "zero", "large_neg"):
                    for ls in ([fixed_ls] if fixed_ls is not None else [0, 1, 2]):
                        bd = 10 if "hbd" in mode else 8
                        c = coeffs(r, n, bd, kind)
                        sc = scan_for(r, n, "perm" if n > 16 else "identity")
                        for use_qm in ([False, True] if has_qm else [False]):
                            qm = r.integers(12, 256, n).astype(np.uint8) if use_qm else None
                            iqm = r.integers(12, 256, n).astype(np.uint8) if use_qm else None
                            yield v, c, t, sc, qm, iqm, ls


def ref_extra(v, qm, iqm, ls):
    name = v[0]
    P = lambda a: ct.c_void_p(a.ctypes.data) if a is not None else ct.c_void_p(0)  # noqa: E731
    if name in ("aom_quantize_b", "aom_highbd_quantize_b", "av1_quantize_b_qm", "av1_highbd_quantize_b_qm"):
        return [P(qm), P(iqm), ct.c_int32(ls)]
    if name in ("av1_quantize_fp_qm", "av1_highbd_quantize_fp_qm"):
        return [P(qm), P(iqm), ct.c_int16(ls)]
    if name == "av1_highbd_quantize_fp":
        return [ct.c_int16(ls)]
    return []
