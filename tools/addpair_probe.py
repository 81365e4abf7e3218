"""Developer probe: per-step phase times of Add Pair 10k through the host API."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "box2d-mt_b200", "python"))
import b2host, scenes
w = b2host.HostWorld(scenes.add_pair(10000), download_bodies=False, events=False)
for s in range(int(sys.argv[1]) if len(sys.argv) > 1 else 40):
    w.step()
    i = w.step_info()
    print(s, "ms %.1f" % float(i["step"]), "vel %.1f pos %.1f toi %.1f trav %.2f bp %.2f" % (i["solveVelocity"], i["solvePosition"], i["solveTOI"], i["solveTraversal"], i["broadphase"]),
          {k: int(i[k]) for k in ("contactCount", "constraintCount", "colourCount", "overflowCount", "toiCandidateCount", "toiSubSteps", "toiNewContactCount", "kernelLaunches")}, flush=True)
