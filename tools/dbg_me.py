#!/usr/bin/env python3
"""debug helper: one ME call of a small picture (compute-sanitizer target)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import svt_av1_psy_b200 as pkg
from svt_av1_psy_b200.pipeline import FramePipeline
from svt_av1_psy_b200.workload import FrameWorkload
pkg.init(0)
fp = FramePipeline(FrameWorkload(384, 256), torch)
s = torch.cuda.current_stream().cuda_stream
fp.call_me_pyramid(s)
fp.call_me_search(s)
torch.cuda.synchronize()
print("me ok", int(fp.me["distortion"].sum()))
