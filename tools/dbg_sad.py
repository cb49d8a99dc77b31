import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import svt_av1_psy_b200 as pkg
import oracle
from helpers import sad_loop_call, sad_pattern, rng
pkg.init(0)
d = pkg.dsp
r = rng(2)
for (bw, bh, sa_w, sa_h) in [(16, 16, 8, 3), (64, 64, 8, 3), (16, 16, 48, 40), (31, 7, 15, 6), (128, 128, 16, 31)]:
    rs = sa_w + bw + 9
    src, ref = sad_pattern("RANDOM", r, bw * bh + 8, rs * (sa_h + bh) + 8)
    want = sad_loop_call(oracle.ref, "svt_sad_loop_kernel_c", src, 0, bw, ref, 0, rs, bh, bw, rs, 0, sa_w, sa_h)
    got = d.svt_sad_loop_kernel(src, 0, bw, ref, 0, rs, bh, bw, rs, 0, sa_w, sa_h)
    print(bw, bh, sa_w, sa_h, got, want, "OK" if got == want else "MISMATCH", flush=True)
