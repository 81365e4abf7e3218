#!/usr/bin/env python3
"""Developer tool: per-kernel device time of one steady-state step (B2CU_TRACE=1). usage: trace_step.py BODIES SETTLE [STEPS] 2>&1 | python tools/trace_agg.py"""
import os, sys
os.environ["B2CU_TRACE"] = "0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "box2d-mt_b200", "python"))
import b2host, scenes
bodies, settle = int(sys.argv[1]), int(sys.argv[2])
w = b2host.HostWorld(scenes.pile(max(16, bodies // 100), 100), download_bodies=False, events=False)
for _ in range(settle):
    w.step()
print("settled", flush=True)
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
os.environ["B2CU_TRACE"] = "1"
for _ in range(steps):
    w.step()
    info = w.step_info()
    print({k: int(info[k]) for k in ("contactCount", "constraintCount", "colourCount", "moveCount", "newContactCount",
                                     "kernelLaunches")}, "step_ms %.3f" % float(info["step"]), flush=True)
os.environ["B2CU_TRACE"] = "0"
