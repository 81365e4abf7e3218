"""The numpy restatement of the broad-phase rules (oracle/b2o_broadphase.py) against the compiled reference: the
contact key set after every step is exactly survivors UNION new pairs of moved proxies, whatever the tree did."""
import numpy as np
import pytest

import b2cuda_types as T
import b2o_broadphase as bp
import parity
import ref
import scenes


@pytest.mark.parametrize("name,steps", [("add_pair", 60), ("pile", 120), ("pyramid", 80)])
def test_contact_set_follows_the_stateless_rule(name, steps):
    scene = {"add_pair": lambda: scenes.add_pair(150), "pile": lambda: scenes.pile(10, 8),
             "pyramid": lambda: scenes.pyramid(8)}[name]()
    # the rule under test is the one of the DISCRETE step (Collide's destruction + one FindNewContacts over the moved
    # proxies); the sub-steps of b2World::SolveTOI move proxies a second time and add pairs of their own
    scene.world_flags &= ~T.WORLD_CONTINUOUS
    r = ref.RefWorld(scene)
    shapes = r.shapes()
    shape_types = shapes["type"]
    r.step()  # creates the initial contacts
    for s in range(steps):
        before = r.proxies()
        contacts = r.contacts()
        bodies = r.bodies()
        old_keys = T.contact_keys(contacts)
        types = bodies["flags"] & T.BODY_TYPE_MASK
        awake = (bodies["flags"] & T.BODY_AWAKE) != 0
        # a contact is skipped by Collide while neither body is awake and non-static (b2ContactManager.cpp:212-216);
        # evaluate that on the state Collide will see (wake-ups of this step's Collide happen after the pass)
        pa, pb = contacts["proxyA"], contacts["proxyB"]
        ba, bb = before["body"][pa], before["body"][pb]
        inactive = ~((awake[ba] & (types[ba] != 0)) | (awake[bb] & (types[bb] != 0)))
        r.step()
        after = r.proxies()
        moved = (after["fat"] != before["fat"]).any(axis=1)
        want = bp.expected_contact_keys(old_keys, inactive, before, after, types, shape_types, moved)
        got = np.sort(T.contact_keys(r.contacts()))
        assert len(got) == len(want) and (got == want).all(), "step %d: %d vs %d" % (s, len(got), len(want))


def test_fat_aabb_rule_matches_reference():
    scene = scenes.pile(8, 6)
    r = ref.RefWorld(scene)
    for _ in range(40):
        before = r.proxies()
        b0 = r.bodies()
        r.step()
        after = r.proxies()
        b1 = r.bodies()
        for p in range(len(before)):
            body = before["body"][p]
            if (b1["flags"][body] & T.BODY_TYPE_MASK) == 0:
                continue
            # displacement = xf.p - xf0.p, xf0 rebuilt from the sweep start (b2ContactManager.cpp:331-357)
            s0, c0 = ref.sincos(b1["a0"][body])
            x0 = b1["c0x"][body] - (c0 * b1["lcx"][body] - s0 * b1["lcy"][body])
            y0 = b1["c0y"][body] - (s0 * b1["lcx"][body] + c0 * b1["lcy"][body])
            disp = (np.float32(b1["px"][body] - np.float32(x0)), np.float32(b1["py"][body] - np.float32(y0)))
            new = bp.move_proxy(before["fat"][p], after["aabb"][p], disp)
            want = after["fat"][p]
            if new is None:
                assert (want == before["fat"][p]).all()
            else:
                assert (new == want).all(), (p, new, want)
