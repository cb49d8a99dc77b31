"""GPU parity: svt_b200_sad_loop_kernel / batch search vs the reference C function
(compute_sad_c.c:58-101) on the reference's own test matrix (test/SadTest.cc:62-107, 432-443)."""
import numpy as np
import pytest

from helpers import rng, sad_loop_call, sad_pattern

pytestmark = pytest.mark.gpu

BLOCKS = [(16, 16), (32, 32), (64, 64), (16, 8), (16, 5), (64, 32), (24, 24), (31, 7), (4, 4), (48, 64), (5, 11),
          (6, 2), (63, 33), (12, 17), (128, 128), (8, 32), (56, 32), (39, 1), (3, 40)]
AREAS = [(8, 3), (8, 15), (16, 31), (12, 31), (15, 6), (48, 40), (64, 25), (1, 1), (7, 2), (96, 24)]
BIG_AREAS = [(192, 75), (240, 200), (640, 400), (336, 192)]


def _checker(oracle):
    if oracle.ref is not None:
        return oracle.ref, "svt_sad_loop_kernel_c"
    return oracle.port, "port_sad_loop"


@pytest.mark.parametrize("pattern", ["REF_MAX", "SRC_MAX", "RANDOM", "FLAT", "UNALIGN"])
def test_sad_loop_matrix(b200, oracle, pattern):
    lib, fn = _checker(oracle)
    r = rng(2)
    for (bw, bh) in BLOCKS:
        for (sa_w, sa_h) in AREAS:
            for skip in (0, 1):
                ref_stride = sa_w + bw + 9
                src_stride = bw + 3
                src_off = ref_off = 0
                if pattern == "UNALIGN":
                    src_off, ref_off = 1, 3
                src, ref = sad_pattern("RANDOM" if pattern == "UNALIGN" else pattern, r, src_stride * bh + 8,
                                       ref_stride * (sa_h + bh) + 8)
                want = sad_loop_call(lib, fn, src, src_off, src_stride, ref, ref_off, ref_stride, bh, bw, ref_stride, skip,
                                     sa_w, sa_h, -7, -9)
                got = b200.svt_sad_loop_kernel(src, src_off, src_stride, ref, ref_off, ref_stride, bh, bw, ref_stride,
                                               skip, sa_w, sa_h, -7, -9)
                assert got == want, (pattern, bw, bh, sa_w, sa_h, skip, got, want)


def test_sad_loop_big_areas(b200, oracle):
    lib, fn = _checker(oracle)
    r = rng(3)
    for (sa_w, sa_h) in BIG_AREAS:
        for (bw, bh) in [(16, 16), (64, 64), (32, 16)]:
            ref_stride = sa_w + bw + 16
            src, ref = sad_pattern("RANDOM", r, bw * bh, ref_stride * (sa_h + bh))
            want = sad_loop_call(lib, fn, src, 0, bw, ref, 0, ref_stride, bh, bw, ref_stride, 0, sa_w, sa_h)
            got = b200.svt_sad_loop_kernel(src, 0, bw, ref, 0, ref_stride, bh, bw, ref_stride, 0, sa_w, sa_h)
            assert got == want, (bw, bh, sa_w, sa_h)


def test_sad_loop_sub_sad_strides(b200, oracle):
    """HME SUB_SAD mode: block rows at twice the plane pitch, search rows at the plane pitch
    (motion_estimation.c:463-481)."""
    lib, fn = _checker(oracle)
    r = rng(4)
    for (bw, bh, sa_w, sa_h) in [(16, 8, 24, 9), (32, 16, 8, 3), (64, 32, 8, 3), (64, 32, 16, 7)]:
        pitch = sa_w + bw + 5
        src, ref = sad_pattern("RANDOM", r, 2 * bw * bh, pitch * (sa_h + 2 * bh))
        want = sad_loop_call(lib, fn, src, 0, 2 * bw, ref, 0, 2 * pitch, bh, bw, pitch, 0, sa_w, sa_h)
        got = b200.svt_sad_loop_kernel(src, 0, 2 * bw, ref, 0, 2 * pitch, bh, bw, pitch, 0, sa_w, sa_h)
        assert got == want
    # a pitch relation that is NOT an integer multiple exercises the per-row staging path
    bw, bh, sa_w, sa_h = 16, 16, 9, 5
    src, ref = sad_pattern("RANDOM", r, bw * bh, 4096)
    want = sad_loop_call(lib, fn, src, 0, bw, ref, 0, 50, bh, bw, 37, 0, sa_w, sa_h)
    got = b200.svt_sad_loop_kernel(src, 0, bw, ref, 0, 50, bh, bw, 37, 0, sa_w, sa_h)
    assert got == want


def test_nxm_sad(b200, oracle):
    r = rng(5)
    oracle.port.port_nxm_sad.restype = np.ctypeslib.ctypes.c_uint32
    for (w, h) in [(4, 4), (8, 8), (16, 16), (64, 64), (128, 128), (31, 7), (5, 11)]:
        src = r.integers(0, 256, (h + 1) * (w + 5), dtype=np.uint8)
        ref = r.integers(0, 256, (h + 1) * (w + 9), dtype=np.uint8)
        want = oracle.port.port_nxm_sad(oracle.p(src), w + 5, oracle.p(ref), w + 9, h, w)
        assert b200.svt_nxm_sad_kernel(src, 0, w + 5, ref, 0, w + 9, h, w) == want


def test_sad_search_batch_picture(b200, oracle):
    """T2: many searches over one padded picture in one launch == per-item reference calls."""
    lib, fn = _checker(oracle)
    r = rng(6)
    W, H, pad = 480, 272, 80
    pitch = W + 2 * pad
    cur = r.integers(0, 256, pitch * (H + 2 * pad), dtype=np.uint8)
    refp = np.roll(cur, 3 * pitch + 5) ^ r.integers(0, 4, cur.size, dtype=np.uint8)
    items = np.zeros(0, dtype=b200.SAD_ITEM_DTYPE)
    lst = []
    for by in range(0, H - 63, 64):
        for bx in range(0, W - 63, 64):
            for (bw, bh, sa_w, sa_h) in [(64, 64, 8, 3), (32, 32, 16, 9), (16, 16, 48, 40)]:
                ox, oy = bx - sa_w // 2, by - sa_h // 2
                lst.append(((pad + by) * pitch + pad + bx, (pad + oy) * pitch + pad + ox, pitch, pitch, pitch, bw, bh,
                            sa_w, sa_h, 0, 0))
    items = np.array(lst, dtype=b200.SAD_ITEM_DTYPE)
    res = b200.sad_search_batch_host(cur, refp, items)
    for it, rr in zip(items, res):
        want = sad_loop_call(lib, fn, cur, int(it["src_off"]), pitch, refp, int(it["ref_off"]), pitch, int(it["block_h"]),
                             int(it["block_w"]), pitch, 0, int(it["sa_w"]), int(it["sa_h"]))
        assert (int(rr["best_sad"]), int(rr["x"]), int(rr["y"])) == want


def test_aom_sad_mxn_and_x4d(b200, oracle):
    """svt_aom_sadMxN / svt_aom_sadMxNx4d (22 sizes): against the reference C functions when oracle/_ref is
    there, else against the plain definition; strides larger than the block, unaligned reference origins."""
    import ctypes as ct
    r = rng(77)
    for (m, n) in b200.SAD_SIZES:
        ss, rs = m + 3, m + 13
        src = r.integers(0, 256, n * ss + 8).astype(np.uint8)
        refs = [r.integers(0, 256, n * rs + 16).astype(np.uint8) for _ in range(4)]
        offs = [0, 1, 2, 5]

        def plain(ref, off):
            a = src[:n * ss].reshape(n, ss)[:, :m].astype(np.int64)
            b = ref[off:off + n * rs].reshape(n, rs)[:, :m].astype(np.int64)
            return int(np.abs(a - b).sum())
        want = [plain(refs[i], offs[i]) for i in range(4)]
        if oracle.ref is not None:
            f = getattr(oracle.ref, "svt_aom_sad%dx%d_c" % (m, n))
            f.restype = ct.c_uint32
            for i in range(4):
                assert f(ct.c_void_p(src.ctypes.data), ss, ct.c_void_p(refs[i].ctypes.data + offs[i]), rs) == want[i]
        got1 = getattr(b200.lib, "svt_b200_aom_sad%dx%d" % (m, n))(src.ctypes.data, ss, refs[0].ctypes.data + offs[0], rs)
        assert got1 == want[0], (m, n)
        arr = (ct.c_void_p * 4)(*[refs[i].ctypes.data + offs[i] for i in range(4)])
        out = np.zeros(4, np.uint32)
        getattr(b200.lib, "svt_b200_aom_sad%dx%dx4d" % (m, n))(src.ctypes.data, ss, arr, rs, out.ctypes.data)
        assert list(out) == want, (m, n)
