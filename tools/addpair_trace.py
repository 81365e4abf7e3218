#!/usr/bin/env python3
"""Developer tool: per-kernel device time of one impact step of Add Pair 10k. usage: addpair_trace.py STEP 2>&1 | python tools/trace_agg.py"""
import os, sys
os.environ["B2CU_TRACE"] = "0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "box2d-mt_b200", "python"))
import b2host, scenes
target = int(sys.argv[1])
w = b2host.HostWorld(scenes.add_pair(10000), download_bodies=False, events=False)
for _ in range(target):
    w.step()
print("settled", flush=True)
os.environ["B2CU_TRACE"] = "1"
w.step()
info = w.step_info()
print({k: int(info[k]) for k in ("contactCount", "constraintCount", "colourCount", "moveCount", "newContactCount",
                                 "kernelLaunches")}, "step_ms %.3f" % float(info["step"]), flush=True)
print("toi", int(info["toiSubSteps"]), float(info["solveTOI"]), flush=True)
os.environ["B2CU_TRACE"] = "0"
