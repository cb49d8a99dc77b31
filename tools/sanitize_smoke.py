#!/usr/bin/env python3
"""One small frame (384x256) through the T2 pipeline, for `compute-sanitizer --tool memcheck|racecheck python
tools/sanitize_smoke.py` on a GPU box (no oracle involved: this only drives the product library)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import svt_av1_psy_b200 as pkg  # noqa: E402
from svt_av1_psy_b200.pipeline import FramePipeline  # noqa: E402
from svt_av1_psy_b200.workload import FrameWorkload  # noqa: E402

pkg.init(0)
fp = FramePipeline(FrameWorkload(384, 256), torch)
fp.load_inputs()
fp.step()
torch.cuda.synchronize()
print("done", int(fp.final.sum()), int(fp.qcoeff.abs().sum()))
