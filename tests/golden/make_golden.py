#!/usr/bin/env python3
"""Regenerates the committed golden fixtures from the reference itself (run in the build container, where
/root/reference is present):

    python tests/golden/make_golden.py

helloworld.txt      the 60 output lines of HelloWorld/HelloWorld.cpp (stock glibc build of the reference)
oracle_hashes.json  FNV-1a trajectory hashes (SURVEY.md 8c recipe) of small scenes stepped by the PARITY oracle
                    (reference sources + interposed correctly-rounded sincos), so that a prebuilt oracle/_ref
                    can be checked on a box where the reference tree is absent
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "box2d-mt_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref  # noqa: E402
import scenes  # noqa: E402


def main():
    w = ref.RefWorld(scenes.hello_world(), threads=1, stock_libm=True)
    with open(os.path.join(HERE, "helloworld.txt"), "w") as f:
        for _ in range(60):
            w.step(1.0 / 60.0, 6, 2)
            b = w.bodies()[1]
            f.write("%4.2f %4.2f %4.2f\n" % (b["px"], b["py"], b["a"]))

    cases = {
        "pyramid20_480": ("pyramid", [20], 480),
        "pyramid8_200": ("pyramid", [8], 200),
        "pile_12x10_240": ("pile", [12, 10], 240),
        "add_pair_200_120": ("add_pair", [200], 120),
        "tumbler_100_240": ("tumbler", [100], 240),
        # joints, every type (the scenes of the joint parity tests)
        "tumbler_joint_100_240": ("tumbler", [100, 0, None, True], 240),
        "hanging_chains_4x12_300": ("hanging_chains", [4, 12], 300),
        "joint_zoo_300": ("joint_zoo", [], 300),
        "rods_and_welds_300": ("rods_and_welds", [], 300),
        "sliders_300": ("sliders", [], 300),
        "machines_300": ("machines", [], 300),
        "pulleys_and_mice_300": ("pulleys_and_mice", [], 300),
        "gears_300": ("gears", [], 300),
    }
    out = {}
    for name, (scene, args, steps) in cases.items():
        w = ref.RefWorld(getattr(scenes, scene)(*args), threads=1)
        for _ in range(steps):
            w.step()
        out[name] = {"scene": scene, "args": args, "steps": steps, "hash": "%08x" % w.hash(),
                     "contacts": w.counts()[2]}
        print(name, out[name])
    with open(os.path.join(HERE, "oracle_hashes.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
