"""Pin the C restatement (oracle/port) against the UNMODIFIED reference objects (oracle/_ref),
on the reference's own test matrices.  CPU only."""
import numpy as np
import pytest

from helpers import rng, sad_loop_call, sad_pattern

BLOCKS = [(16, 16), (32, 32), (64, 64), (16, 8), (64, 32), (24, 24), (31, 7), (4, 4), (48, 64), (5, 11), (128, 128)]
AREAS = [(8, 3), (16, 31), (15, 6), (48, 40), (64, 25)]


@pytest.mark.parametrize("pattern", ["REF_MAX", "SRC_MAX", "RANDOM", "FLAT"])
def test_port_sad_loop_matches_reference(oracle, refc, pattern):
    r = rng(1)
    for (bw, bh) in BLOCKS:
        for (sa_w, sa_h) in AREAS:
            for skip in (0, 1):
                ref_stride = sa_w + bw + 9
                src_stride = bw + 3
                src, ref = sad_pattern(pattern, r, src_stride * bh, ref_stride * (sa_h + bh))
                a = sad_loop_call(oracle.port, "port_sad_loop", src, 0, src_stride, ref, 0, ref_stride, bh, bw, ref_stride,
                                  skip, sa_w, sa_h, -7, -9)
                b = sad_loop_call(refc, "svt_sad_loop_kernel_c", src, 0, src_stride, ref, 0, ref_stride, bh, bw,
                                  ref_stride, skip, sa_w, sa_h, -7, -9)
                assert a == b, (bw, bh, sa_w, sa_h, skip)
