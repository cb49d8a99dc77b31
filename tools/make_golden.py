#!/usr/bin/env python3
"""Generate tests/golden/frame_<W>x<H>.json: SHA-256 of every output of one synthetic frame computed by the
REFERENCE's own kernels (oracle/_ref, C tier, driven by oracle/ref_driver.c).  Run in the build container (where
/root/reference exists and `python __graft_entry__.py --oracle` has built oracle/_ref); the fixture travels with
the repository so that the GPU parity test also works where oracle/_ref is absent."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402
from svt_av1_psy_b200.workload import FrameWorkload  # noqa: E402


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).view(np.uint8).tobytes()).hexdigest()


def golden_for(width, height, seed=20260923):
    ref, _, _ = bench.load_reference()
    assert ref is not None, "oracle/_ref is not built"
    ref.ref_set_tier(0)
    wl = FrameWorkload(width, height, seed=seed)
    fr = bench.RefFrame(wl, ref)
    fr.step()
    outs = {"me_sad": fr.me_sad, "me_mv": fr.me_mv, "hme_centre": fr.me_c, "coeff": fr.coeff, "qcoeff": fr.q, "dqcoeff": fr.dq,
            "eob": fr.eobs, "recon": fr.recon, "cdef_mse": fr.mse, "cdef_dir": fr.dirs, "cdef_out": fr.cdef_out,
            "wiener_M": fr.M, "wiener_H": fr.Hm, "final": fr.final}
    return {"width": width, "height": height, "seed": seed, "reference_tier": "C (ref_set_tier(0))",
            "sha256": {k: digest(v) for k, v in outs.items()},
            "shape": {k: list(np.asarray(v).shape) for k, v in outs.items()}}


if __name__ == "__main__":
    for (w, h) in ((384, 256), (640, 360)):
        g = golden_for(w, h)
        path = os.path.join(ROOT, "tests", "golden", "frame_%dx%d.json" % (w, h))
        json.dump(g, open(path, "w"), indent=1, sort_keys=True)
        print("wrote", path)
