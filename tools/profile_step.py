#!/usr/bin/env python3
"""One frame step of the hot path between cudaProfilerStart/Stop, every T2 call strictly in order on one stream:
   ncu --profile-from-start off ... python tools/profile_step.py [--config K]
(launch list: --metrics gpu__time_duration.sum --clock-control none; full capture: --set full --import-source on)"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=1)
ap.add_argument("--steps", type=int, default=1)
args = ap.parse_args()
import torch  # noqa: E402
import svt_av1_psy_b200 as pkg  # noqa: E402
from svt_av1_psy_b200.pipeline import FramePipeline  # noqa: E402
from svt_av1_psy_b200.workload import FrameWorkload  # noqa: E402

pkg.init(0)
fp = FramePipeline(FrameWorkload.from_config(args.config), torch)
for _ in range(3):
    fp.step()
torch.cuda.synchronize()
ev = [torch.cuda.Event() for _ in range(len(FramePipeline.CALLS) + 1)]
torch.cuda.cudart().cudaProfilerStart()
for _ in range(args.steps):
    fp.step(ev)  # with events: the calls run one after the other on the current stream
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled %d step(s) of config %d" % (args.steps, args.config))
