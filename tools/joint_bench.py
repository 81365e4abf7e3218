"""Step time of joint-heavy worlds on the device (kernel-only, world resident): hanging chains and a big Tumbler.
Usage: python tools/joint_bench.py [STEPS]"""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "box2d-mt_b200", "python"))
import numpy as np  # noqa: E402

import b2cuda_types as T  # noqa: E402
import b2host  # noqa: E402
import scenes  # noqa: E402


def run(name, scene, steps, warm=60):
    h = b2host.HostWorld(scene, download_bodies=False, events=False)
    for _ in range(warm):
        h.step()
    ms = []
    for _ in range(steps):
        h.step()
        ms.append(h.profile()[0])
    info = h.step_info()
    print("%-28s bodies %7d joints %7d contacts %8d constraints %8d colours %2d: %.3f ms/step (device), solve %.3f"
          % (name, h.counts()[0], h.joint_count(), int(info["contactCount"]), int(info["constraintCount"]),
             int(info["colourCount"]), float(np.mean(ms)), float(h.profile()[2])))


if __name__ == "__main__":
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    run("hanging_chains 1000x40", scenes.hanging_chains(1000, 40), steps)
    run("hanging_chains 5000x20", scenes.hanging_chains(5000, 20), steps)
    run("tumbler_joint 100k boxes", scenes.tumbler(100000, motor_joint=True), steps)
    run("tumbler kinematic 100k boxes", scenes.tumbler(100000), steps)
