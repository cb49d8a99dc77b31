"""GPU parity: Wiener separable filter and statistics vs the reference C functions (fixtures after
test/wiener_convolve_test.cc and test/RestorationPickTest.cc)."""
import ctypes as ct

import numpy as np
import pytest

import rest_helpers as rh
from helpers import rng

pytestmark = pytest.mark.gpu


def test_wiener_convolve_t1(b200, oracle):
    r = rng(90)
    for bd in (8, 10, 12):
        dt = np.uint8 if bd == 8 else np.uint16
        for (w, h) in [(64, 64), (48, 56), (16, 8), (64, 20), (32, 64), (8, 3)]:
            for kind in ("random", "max", "zero"):
                ss = w + 16
                src = {"random": r.integers(0, 1 << bd, ss * (h + 16)), "max": np.full(ss * (h + 16), (1 << bd) - 1),
                       "zero": np.zeros(ss * (h + 16))}[kind].astype(dt)
                fx, fy = rh.wiener_taps(r), rh.wiener_taps(r)
                off = 5 * ss + 6
                if oracle.ref is not None:
                    want = rh.ref_wiener(oracle.ref, src, off, ss, w, h, fx, fy, bd)
                else:
                    want = rh.port_wiener(oracle.port, src.astype(np.uint16), off, ss, w, h, fx, fy, bd, 1 if bd == 8 else 0).astype(dt)
                got = np.zeros(h * w, dt)
                cp = b200.ConvolveParams(); cp.round_0 = 5 if bd == 12 else 3; cp.round_1 = 14 - cp.round_0
                if bd == 8:
                    b200.lib.svt_b200_av1_wiener_convolve_add_src(rh.P(src, off), ss, rh.P(got), w, rh.P(fx), rh.P(fy), w, h, ct.byref(cp))
                else:
                    b200.lib.svt_b200_av1_highbd_wiener_convolve_add_src(rh.P(src, off), ss, rh.P(got), w, rh.P(fx), rh.P(fy), w, h,
                                                                         ct.byref(cp), bd)
                assert np.array_equal(got, want), (bd, w, h, kind)


def test_compute_stats_t1(b200, oracle):
    r = rng(91)
    for bd in (8, 10, 12):
        dt = np.uint8 if bd == 8 else np.uint16
        for win in (7, 5, 3):
            for (W, Hh, hs, he, vs, ve) in [(72, 56, 5, 61, 4, 50), (160, 150, 4, 156, 3, 147), (40, 20, 3, 4, 3, 17)]:
                for kind in ("random", "extreme"):
                    if kind == "random":
                        dgd = r.integers(0, 1 << bd, W * Hh).astype(dt); src = r.integers(0, 1 << bd, W * Hh).astype(dt)
                    else:  # checkerboard of min/max: largest |y| products, exercises the int32 flush bound
                        yy, xx = np.mgrid[0:Hh, 0:W]
                        dgd = (((xx + yy) % 2) * ((1 << bd) - 1)).astype(dt).reshape(-1)
                        src = (((xx + yy + 1) % 2) * ((1 << bd) - 1)).astype(dt).reshape(-1)
                    if oracle.ref is not None:
                        want = rh.ref_stats(oracle.ref, win, dgd, src, hs, he, vs, ve, W, W, bd)
                    else:
                        want = rh.port_stats(oracle.port, win, dgd.astype(np.uint16), src.astype(np.uint16), hs, he, vs, ve, W, W, bd)
                    M = np.zeros(49, np.int64); H = np.zeros(2401, np.int64)
                    if bd == 8:
                        b200.lib.svt_b200_av1_compute_stats(win, rh.P(dgd), rh.P(src), hs, he, vs, ve, W, W, rh.P(M), rh.P(H))
                    else:
                        b200.lib.svt_b200_av1_compute_stats_highbd(win, rh.P(dgd), rh.P(src), hs, he, vs, ve, W, W, rh.P(M), rh.P(H), bd)
                    assert np.array_equal(M[:win * win], want[0]), (bd, win, W, kind)
                    assert np.array_equal(H[:win ** 4], want[1]), (bd, win, W, kind)


def test_compute_stats_large_region_8bit(b200, oracle):
    """A region far larger than a restoration unit: every CTA of the tensor-core kernel passes the
    33025-pixel bound of its int32 totals and must fold them into the int64 partial on the way; the
    half-black / half-white picture gives long runs of same-sign maximal products."""
    r = rng(92)
    W, Hh, hs, he, vs, ve = 840, 726, 3, 837, 3, 723
    yy, xx = np.mgrid[0:Hh, 0:W]
    for kind in ("random", "halves"):
        if kind == "random":
            dgd = r.integers(0, 256, W * Hh).astype(np.uint8); src = r.integers(0, 256, W * Hh).astype(np.uint8)
        else:
            dgd = ((xx > W // 2) * 255).astype(np.uint8).reshape(-1); src = ((yy > Hh // 2) * 255).astype(np.uint8).reshape(-1)
        for win in (7, 5):
            if oracle.ref is not None:
                want = rh.ref_stats(oracle.ref, win, dgd, src, hs, he, vs, ve, W, W, 8)
            else:
                want = rh.port_stats(oracle.port, win, dgd.astype(np.uint16), src.astype(np.uint16), hs, he, vs, ve, W, W, 8)
            M = np.zeros(49, np.int64); H = np.zeros(2401, np.int64)
            b200.lib.svt_b200_av1_compute_stats(win, rh.P(dgd), rh.P(src), hs, he, vs, ve, W, W, rh.P(M), rh.P(H))
            assert np.array_equal(M[:win * win], want[0]), (win, kind)
            assert np.array_equal(H[:win ** 4], want[1]), (win, kind)
