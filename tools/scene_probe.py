"""Per-step device profile of a named scene: python tools/scene_probe.py NAME ARGS... STEPS  (e.g. add_pair 10000 12).
B2PROBE_CONTINUOUS=1 keeps continuous physics on (the first TOI pass runs; its sub-steps do not exist yet)."""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "box2d-mt_b200", "python"))
import b2cuda_types as T  # noqa: E402
import b2host  # noqa: E402
import scenes  # noqa: E402

name = sys.argv[1]
args = [int(a) for a in sys.argv[2:-1]]
steps = int(sys.argv[-1])
scene = getattr(scenes, name)(*args)
if not os.environ.get("B2PROBE_CONTINUOUS"):
    scene.world_flags &= ~T.WORLD_CONTINUOUS
h = b2host.HostWorld(scene, download_bodies=False, events=False)
for s in range(steps):
    t = time.time()
    h.step()
    wall = (time.time() - t) * 1e3
    i = h.step_info()
    p = h.profile()
    print("step %2d wall %8.2f ms device %8.2f (collide %.2f solve %.2f [traversal %.2f init %.2f vel %.2f pos %.2f] broad %.2f) contacts %d "
          "constraints %d colours %d overflow %d launches %d toi %d pending %d" % (s, wall, p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[8], i["contactCount"],
                                                                i["constraintCount"], i["colourCount"], i["overflowCount"], i["kernelLaunches"],
                                                                i["toiCandidateCount"], i["toiEventPending"]))
