#!/usr/bin/env python3
"""Developer tool: settle a pile through the host API, then run a few steps between cudaProfilerStart/Stop so that
`ncu --profile-from-start off` captures only steady-state steps.
usage: profile_step.py BODIES SETTLE STEPS"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "box2d-mt_b200", "python"))
import b2host, scenes

bodies, settle, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
w = b2host.HostWorld(scenes.pile(max(16, bodies // 100), 100), download_bodies=False, events=False)
for _ in range(settle):
    w.step()
rt = ctypes.CDLL("libcudart.so")
rt.cudaProfilerStart()
for _ in range(steps):
    w.step()
rt.cudaProfilerStop()
i = w.step_info()
print("bodies %d contacts %d constraints %d colours %d launches %d step %.3f ms" % (
    i["bodyCount"], i["contactCount"], i["constraintCount"], i["colourCount"], i["kernelLaunches"], i["step"]))
