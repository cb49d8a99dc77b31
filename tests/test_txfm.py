"""GPU parity: forward / inverse 2-D transforms vs the reference C functions for every transform
size x valid type x {8,10}-bit, full-buffer equality (the structure of the reference's
test/FwdTxfm2dAsmTest.cc:313-377 and InvTxfm2dAsmTest.cc)."""
import numpy as np
import pytest

from helpers import rng
from txfm_helpers import (TX_H, TX_W, coeff_input, mask_written, port_fwd, port_inv, ref_fwd, ref_fwd_partial, ref_inv,
                          residual_input, valid)

pytestmark = pytest.mark.gpu


def _fwd_checker(oracle):
    if oracle.ref is not None:
        return lambda res, st, ty, sz, bd: ref_fwd(oracle.ref, res, st, ty, sz, bd)
    return lambda res, st, ty, sz, bd: port_fwd(oracle.port, res, st, ty, sz)


def _inv_checker(oracle):
    if oracle.ref is not None:
        return lambda c, p, sr, sw, ty, sz, bd: ref_inv(oracle.ref, c, p, sr, sw, ty, sz, bd)
    return lambda c, p, sr, sw, ty, sz, bd: port_inv(oracle.port, c, p, sr, sw, ty, sz, bd)


def test_valid_table(b200):
    for sz in range(19):
        for ty in range(16):
            assert b200.txfm_valid(sz, ty) == valid(sz, ty)


@pytest.mark.parametrize("kind", ["random", "max", "min", "zero"])
def test_fwd_txfm_all_sizes_types(b200, oracle, kind):
    chk = _fwd_checker(oracle)
    r = rng(20)
    for sz in range(19):
        for ty in range(16):
            if not valid(sz, ty):
                continue
            for bd in (8, 10):
                res, stride = residual_input(r, sz, bd, kind)
                want = chk(res, stride, ty, sz, bd)
                got = b200.svt_av1_fwd_txfm2d(res, stride, ty, sz, bd, named=(bd == 10))
                assert np.array_equal(got, want), (sz, ty, bd, kind)


@pytest.mark.parametrize("level", [1, 2])
def test_fwd_txfm_partial_n2_n4(b200, oracle, level):
    """svt_av1_fwd_txfm2d_WxH_N2 / _N4: the top-left half / quarter of every dimension, zero elsewhere
    (the reference's own N2/N4 C kernels when oracle/_ref is there, else the masked full transform)."""
    chk = _fwd_checker(oracle)
    r = rng(22 + level)
    for sz in range(19):
        w, h = TX_W[sz], TX_H[sz]
        for ty in range(16):
            if not valid(sz, ty):
                continue
            for kind in ("random", "max"):
                res, stride = residual_input(r, sz, 8, kind)
                if oracle.ref is not None:
                    want = ref_fwd_partial(oracle.ref, res, stride, ty, sz, level)
                else:
                    want = chk(res, stride, ty, sz, 8).reshape(h, w).copy()
                    want[max(h >> level, 1):, :] = 0
                    want[:, max(w >> level, 1):] = 0
                    want = want.reshape(-1)
                got = b200.svt_av1_fwd_txfm2d_partial(res, stride, ty, sz, level, named=(kind == "max"))
                assert np.array_equal(got, want), (sz, ty, level, kind)


@pytest.mark.parametrize("kind", ["real", "sparse", "dc", "extreme", "zero"])
def test_inv_txfm_all_sizes_types(b200, oracle, kind):
    fchk, chk = _fwd_checker(oracle), _inv_checker(oracle)
    r = rng(21)
    for sz in range(19):
        for ty in range(16):
            if not valid(sz, ty):
                continue
            for bd in (8, 10, 12):
                w, h = TX_W[sz], TX_H[sz]
                c = coeff_input(r, sz, bd, kind, lambda res, st: fchk(res, st, ty, sz, bd))
                pred = r.integers(0, 1 << bd, h * (w + 5)).astype(np.uint16)
                want = chk(c, pred, w + 5, w + 2, ty, sz, bd)
                got = b200.svt_av1_inv_txfm2d_add(c, pred, w + 5, w + 2, ty, sz, bd) if bd != 10 else \
                    b200.svt_av1_inv_txfm2d_add_named(c, pred, w + 5, w + 2, ty, sz, bd)
                assert np.array_equal(mask_written(got, w + 2, w, h), mask_written(want, w + 2, w, h)), (sz, ty, bd, kind)


def test_inv_txfm_add_8bit_pixels(b200, oracle):
    """svt_av1_inv_txfm_add (8-bit pixels through a 16-bit temporary, inv_transforms.c:3177)."""
    fchk, chk = _fwd_checker(oracle), _inv_checker(oracle)
    r = rng(22)
    for sz in range(19):
        for ty in (0, 1, 9, 15):
            if not valid(sz, ty):
                continue
            w, h = TX_W[sz], TX_H[sz]
            c = coeff_input(r, sz, 8, "real", lambda res, st: fchk(res, st, ty, sz, 8))
            pred8 = r.integers(0, 256, h * (w + 1)).astype(np.uint8)
            want = chk(c, pred8.astype(np.uint16), w + 1, w, ty, sz, 8).astype(np.uint8)
            got = b200.svt_av1_inv_txfm_add(c, pred8, w + 1, w, ty, sz)
            assert np.array_equal(got, want), (sz, ty)


def test_fwd_txfm_batch_mixed_sizes(b200, oracle):
    """T2: one call, mixed sizes/types over one residual plane == per-block reference calls."""
    chk = _fwd_checker(oracle)
    r = rng(23)
    W, H = 256, 128
    plane = r.integers(-255, 256, W * H).astype(np.int16)
    items, off = [], 0
    for by in range(0, H, 64):
        for bx in range(0, W, 64):
            for sz in (4, 3, 2, 1, 0, 9, 12, 14, 5):
                ty = [t for t in (0, 9, 3, 5) if valid(sz, t)][(bx // 64 + sz) % len([t for t in (0, 9, 3, 5) if valid(sz, t)])]
                items.append((by * W + bx, off, W, sz, ty, 0))
                off += TX_W[sz] * TX_H[sz]
    items = np.array(items, dtype=b200.FWD_ITEM_DTYPE)
    coeff = np.zeros(off, np.int32)
    b200.fwd_txfm_batch_host(plane, coeff, items)
    for it in items:
        sz, ty = int(it["tx_size"]), int(it["tx_type"])
        n = TX_W[sz] * TX_H[sz]
        want = chk(plane[int(it["src_off"]):], W, ty, sz, 8)
        assert np.array_equal(coeff[int(it["dst_off"]):int(it["dst_off"]) + n], want), (sz, ty)


def test_handle_transform_energy_and_repack(b200, oracle):
    """svt_handle_transformWxH / _N2_N4: energy of the dropped coefficients + in-place re-pack, whole buffer compared."""
    import ctypes as ct
    r = rng(31)
    for name, (w, h) in {"16x64": (16, 64), "32x64": (32, 64), "64x16": (64, 16), "64x32": (64, 32), "64x64": (64, 64)}.items():
        for sfx in ("", "_N2_N4"):
            x = r.integers(-(1 << 20), 1 << 20, w * h).astype(np.int32)
            want = x.copy()
            if oracle.ref is not None:
                f = getattr(oracle.ref, "svt_handle_transform%s%s_c" % (name, sfx))
                f.restype = ct.c_uint64
                e_want = f(ct.c_void_p(want.ctypes.data))
            else:
                a = want.reshape(h, w).astype(np.int64)
                wp, hp = min(w, 32), min(h, 32)
                e_want = 0 if sfx else int((a * a).sum() - (a[:hp, :wp] ** 2).sum())
                if w == 64:
                    keep = a[:hp, :32].astype(np.int32).copy()
                    want[32:hp * 32] = keep.reshape(-1)[32:]
            got = x.copy()
            e_got = getattr(b200.lib, "svt_b200_handle_transform%s%s" % (name, sfx))(got.ctypes.data)
            assert e_got == e_want, (name, sfx)
            assert np.array_equal(got, want), (name, sfx)
