#!/usr/bin/env python3
"""Developer tool: per-kernel device time of one steady-state step (B2CU_TRACE=1). usage: trace_step.py BODIES SETTLE"""
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
