#!/usr/bin/env python3
"""Dev-time tool: run the reference's OWN control derivation (svt_aom_sig_deriv_me, enc_mode_config.c:681, through
oracle/ref_me_b64.c) for the workload's ME picture (workload.ME_PICTURE) at every (preset, input-resolution class) and write
the flattened MeContext controls to svt-av1-psy_b200/me_controls.json.  The generated file is committed (the GPU box has no
/root/reference); tests/test_oracle_pins.py re-derives and compares where the reference is present."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from oracle import support as sp  # noqa: E402
from oracle.frame_ref import load_workload_module  # noqa: E402

wlm = load_workload_module()
layout = wlm.dsp  # svt-av1-psy_b200/layout.py

# one picture size inside each resolution class (svt_aom_derive_input_resolution)
CLASS_SIZES = {0: (384, 256), 1: (640, 360), 2: (848, 480), 3: (1280, 720), 4: (1920, 1080), 5: (3840, 2160)}


def cfg_for(preset):
    mp = wlm.ME_PICTURE[preset]
    return sp.me_b64_cfg(preset=preset, qp=mp["qp"], n_ref=mp["n_ref"], poc_dist=wlm.ME_DIST, temporal_layer_index=wlm.ME_TEMPORAL_LAYER,
                         hierarchical_levels=wlm.ME_HIERARCHICAL_LEVELS, is_ref=wlm.ME_IS_REF, max_l=mp["max_l"], only_l_bwd=mp["only_l_bwd"],
                         safe_limit_nref=mp["safe_limit_nref"], gm_enabled=mp["gm_enabled"])


def derive(preset, w, h):
    shapes = layout.me_plane_shapes(w, h)
    blank = [np.zeros((s[0], s[1]), np.uint8) for s in shapes]
    mp = wlm.ME_PICTURE[preset]
    ctrl, _ = sp.ref_me_b64_picture(oracle.ref, blank, [blank] * (mp["n_ref"][0] + mp["n_ref"][1]), shapes, cfg_for(preset), run=False)
    return ctrl.as_dict()


def all_controls():
    out = {}
    for preset in sorted(wlm.ME_PICTURE):
        for cls, (w, h) in CLASS_SIZES.items():
            assert wlm.input_resolution_class(w, h) == cls
            out["m%d_class%d" % (preset, cls)] = derive(preset, w, h)
    return out


if __name__ == "__main__":
    path = os.path.join(ROOT, "svt-av1-psy_b200", "me_controls.json")
    with open(path, "w") as f:
        json.dump(all_controls(), f, indent=1, sort_keys=True)
    print("wrote", path)
