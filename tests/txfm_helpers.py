"""ctypes callers for the transform checkers (oracle.ref = unmodified reference, oracle.port = restatement)."""
import ctypes as ct

import numpy as np

TX_W = [4, 8, 16, 32, 64, 4, 8, 8, 16, 16, 32, 32, 64, 4, 16, 8, 32, 16, 64]
TX_H = [4, 8, 16, 32, 64, 8, 4, 16, 8, 32, 16, 64, 32, 16, 4, 32, 8, 64, 16]
INV_SIG_A = {0, 1, 2, 3, 4}
INV_SIG_B = {5, 6, 13, 14}


def valid(sz, ty):
    m = max(TX_W[sz], TX_H[sz])
    if m == 64:
        return ty == 0
    if m == 32:
        return ty in (0, 9)
    return True


def ref_fwd(ref, residual, stride, ty, sz, bd=8):
    name = "svt_av1_transform_two_d_%dx%d_c" % (TX_W[sz], TX_H[sz]) if sz <= 4 else "svt_av1_fwd_txfm2d_%dx%d_c" % (
        TX_W[sz], TX_H[sz])
    f = getattr(ref, name)
    f.restype = None
    out = np.zeros(TX_W[sz] * TX_H[sz] + 64, np.int32)
    f(ct.c_void_p(residual.ctypes.data), ct.c_void_p(out.ctypes.data), ct.c_uint32(stride), ct.c_int(ty), ct.c_uint8(bd))
    return out[:TX_W[sz] * TX_H[sz]]


def ref_fwd_partial(ref, residual, stride, ty, sz, level, bd=8):
    """the reference's N2 (level 1) / N4 (level 2) forward transform (transforms.c:5202-6990)"""
    n = "N%d" % (2 * level)
    name = ("svt_aom_transform_two_d_%dx%d_%s_c" if sz <= 4 else "svt_av1_fwd_txfm2d_%dx%d_%s_c") % (TX_W[sz], TX_H[sz], n)
    f = getattr(ref, name)
    f.restype = None
    out = np.full(TX_W[sz] * TX_H[sz] + 64, 0x5a5a5a5a, np.int32)
    f(ct.c_void_p(residual.ctypes.data), ct.c_void_p(out.ctypes.data), ct.c_uint32(stride), ct.c_int(ty), ct.c_uint8(bd))
    return out[:TX_W[sz] * TX_H[sz]]


def port_fwd(port, residual, stride, ty, sz):
    out = np.zeros(TX_W[sz] * TX_H[sz], np.int32)
    port.port_fwd_txfm2d.restype = None
    port.port_fwd_txfm2d(ct.c_void_p(residual.ctypes.data), ct.c_void_p(out.ctypes.data), ct.c_uint32(stride), ty, sz)
    return out


def ref_inv(ref, coeff, pred, stride_r, stride_w, ty, sz, bd):
    f = getattr(ref, "svt_av1_inv_txfm2d_add_%dx%d_c" % (TX_W[sz], TX_H[sz]))
    f.restype = None
    out = np.zeros(TX_H[sz] * stride_w, np.uint16)
    args = [ct.c_void_p(coeff.ctypes.data), ct.c_void_p(pred.ctypes.data), ct.c_int32(stride_r),
            ct.c_void_p(out.ctypes.data), ct.c_int32(stride_w), ct.c_int(ty)]
    if sz in INV_SIG_A:
        args += [ct.c_int32(bd)]
    elif sz in INV_SIG_B:
        args += [ct.c_int(sz), ct.c_int32(bd)]
    else:
        args += [ct.c_int(sz), ct.c_int32(TX_W[sz] * TX_H[sz]), ct.c_int32(bd)]
    f(*args)
    return out


def port_inv(port, coeff, pred, stride_r, stride_w, ty, sz, bd):
    out = np.zeros(TX_H[sz] * stride_w, np.uint16)
    port.port_inv_txfm2d_add.restype = None
    port.port_inv_txfm2d_add(ct.c_void_p(coeff.ctypes.data), ct.c_void_p(pred.ctypes.data), ct.c_int32(stride_r),
                             ct.c_void_p(out.ctypes.data), ct.c_int32(stride_w), ty, sz, bd)
    return out


def mask_written(plane, stride, w, h):
    """only the W columns of each row are defined output"""
    return plane.reshape(h, stride)[:, :w].copy()


def residual_input(r, sz, bd, kind):
    w, h = TX_W[sz], TX_H[sz]
    stride = w + 3
    lim = (1 << bd) - 1
    if kind == "max":
        a = np.full(h * stride, lim, np.int16)
    elif kind == "min":
        a = np.full(h * stride, -lim, np.int16)
    elif kind == "zero":
        a = np.zeros(h * stride, np.int16)
    else:
        a = r.integers(-lim, lim + 1, h * stride).astype(np.int16)
    return a, stride


def coeff_input(r, sz, bd, kind, fwd_fn):
    """realistic dequantised coefficients: forward-transform a random residual, optionally sparsify"""
    w, h = TX_W[sz], TX_H[sz]
    wp, hp = min(w, 32), min(h, 32)
    if kind == "zero":
        return np.zeros(wp * hp, np.int32)
    if kind == "dc":
        c = np.zeros(wp * hp, np.int32)
        c[0] = int(r.integers(-(1 << (bd + 5)), 1 << (bd + 5)))
        return c
    if kind == "extreme":  # exercises the clamps
        return r.integers(-(1 << (bd + 9)), 1 << (bd + 9), wp * hp).astype(np.int32)
    res, stride = residual_input(r, sz, bd, "random")
    full = fwd_fn(res, stride).reshape(h, w)
    c = np.ascontiguousarray(full[:hp, :wp]).reshape(-1).astype(np.int32)
    if kind == "sparse":
        c[r.random(c.size) < 0.8] = 0
    return c
