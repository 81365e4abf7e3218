#!/usr/bin/env python3
"""Developer tool: step a sleeping-enabled pile and print awake count / step time as it goes to sleep.
usage: sleep_probe.py COLUMNS ROWS STEPS"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "box2d-mt_b200", "python"))
import b2host, scenes
cols, rows, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
w = b2host.HostWorld(scenes.pile(cols, rows, sleep=True), download_bodies=False, events=False)
for s in range(steps):
    w.step()
    if s % 100 == 99 or s == steps - 1:
        i = w.step_info()
        print("step %4d awake %7d contacts %8d constraints %8d launches %3d step %.3f ms" % (
            s + 1, int(i["awakeBodyCount"]), int(i["contactCount"]), int(i["constraintCount"]), int(i["kernelLaunches"]),
            float(i["step"])), flush=True)
