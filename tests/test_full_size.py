"""Full-size checks (BASELINE.json configs[4]: the 1M-body pile) through size-independent properties.  The oracle
needs ~3 s per step at this size: parity against it is covered at the sizes it finishes in seconds
(test_gpu_parity.py, test_edge_cases.py, test_host_api.py); here the same device code runs at full size.

* determinism: two device worlds built from the same scene agree bit for bit after 40 steps (TestMT.cpp's rule,
  A/B instead of reference/GPU);
* the solver order is a proper edge colouring: no two constraints of one colour share a non-static body, every
  touching solid contact of an awake island is in it exactly once (what makes the coloured Gauss-Seidel equal to a
  sequential one);
* the contact set is sorted, duplicate-free, and complete: on a window of the pile every pair of overlapping fat
  boxes has its contact (brute force over the window), and the only extra contacts are those whose boxes separated
  in the last step (they are destroyed by the next Collide, b2ContactManager.cpp:199-207);
* physical invariants of the reference's SleepCollideTest (nothing below the floor) and bounded penetration.
"""
import numpy as np
import pytest

import b2cuda_types as T
import b2host
import scenes

pytestmark = pytest.mark.gpu

COLUMNS, ROWS = 10000, 100  # 1,000,000 bodies + the container


@pytest.fixture(scope="module")
def pile_worlds():
    scene = scenes.pile(COLUMNS, ROWS)
    a = b2host.HostWorld(scene, download_bodies=False, events=False)
    b = b2host.HostWorld(scene, download_bodies=False, events=False)
    for _ in range(40):
        a.step()
        b.step()
    return scene, a, b


def test_one_million_bodies_deterministic(pile_worlds):
    scene, a, b = pile_worlds
    ba, bb = a.bodies(), b.bodies()
    assert len(ba) == COLUMNS * ROWS + 1
    assert ba.tobytes() == bb.tobytes()
    ca, cb = a.device_world().get_contacts(), b.device_world().get_contacts()
    assert ca.tobytes() == cb.tobytes()


def test_one_million_bodies_contact_set_and_colouring(pile_worlds):
    scene, a, _ = pile_worlds
    dev = a.device_world()
    contacts = dev.get_contacts()
    keys = T.contact_keys(contacts)
    assert (np.diff(keys.astype(np.uint64)) > 0).all(), "contact keys must be strictly ascending"
    proxies = dev.get_proxies()
    bodies = a.bodies()
    body_a = proxies["body"][contacts["proxyA"]]
    body_b = proxies["body"][contacts["proxyB"]]
    assert (body_a != body_b).all()

    order_keys, colours = dev.solver_order()
    assert len(np.unique(order_keys)) == len(order_keys)
    touching = (contacts["flags"] & (T.CONTACT_TOUCHING | T.CONTACT_ENABLED)) == (T.CONTACT_TOUCHING | T.CONTACT_ENABLED)
    # no sleeping in this scene: every touching contact is a constraint
    assert set(order_keys.tolist()) == set(keys[touching].tolist())
    idx = np.searchsorted(keys, order_keys)
    ca, cb = body_a[idx], body_b[idx]
    dynamic = (bodies["flags"] & T.BODY_TYPE_MASK) == T.DYNAMIC_BODY
    for c in np.unique(colours):
        sel = colours == c
        used = np.concatenate([ca[sel][dynamic[ca[sel]]], cb[sel][dynamic[cb[sel]]]])
        assert len(np.unique(used)) == len(used), "colour %d has two constraints on one body" % c
    assert len(np.unique(colours)) <= 32


def test_one_million_bodies_invariants(pile_worlds):
    scene, a, _ = pile_worlds
    for _ in range(20):
        a.step()
    bodies = a.bodies()
    dyn = bodies[1:]
    assert np.isfinite(dyn["px"]).all() and np.isfinite(dyn["py"]).all() and np.isfinite(dyn["vx"]).all()
    # SleepCollideTest.h:104-110: nothing falls through the floor (top of the container floor is y = 0)
    assert dyn["py"].min() > -0.05
    width = COLUMNS * 0.56
    assert dyn["px"].min() > -0.5 * width - 0.05 and dyn["px"].max() < 0.5 * width + 0.05
    contacts = a.device_world().get_contacts()
    assert len(contacts) > 2 * len(dyn)


def test_window_of_the_pile_follows_the_broadphase_rule(pile_worlds):
    """all fat-box overlaps among the proxies of a window have a contact and nothing else has (the stateless form of
    the pair rule, checked by brute force on ~2000 proxies of the 1M)"""
    scene, a, _ = pile_worlds
    dev = a.device_world()
    proxies = dev.get_proxies()
    contacts = dev.get_contacts()
    fat = proxies["fat"]
    cx = 0.5 * (fat[:, 0] + fat[:, 2])
    window = np.nonzero((cx > 100.0) & (cx < 111.0))[0]
    window = window[proxies["body"][window] != 0]
    assert 1000 < len(window) < 4000
    f = fat[window]
    overlap = ((f[:, None, 0] <= f[None, :, 2]) & (f[None, :, 0] <= f[:, None, 2]) &
               (f[:, None, 1] <= f[None, :, 3]) & (f[None, :, 1] <= f[:, None, 3]))
    iu = np.triu_indices(len(window), 1)
    pairs = np.stack([window[iu[0]][overlap[iu]], window[iu[1]][overlap[iu]]], axis=1)
    want = set(((np.minimum(pairs[:, 0], pairs[:, 1]).astype(np.uint64) << np.uint64(32)) |
                np.maximum(pairs[:, 0], pairs[:, 1]).astype(np.uint64)).tolist())
    inside = np.zeros(len(proxies), bool)
    inside[window] = True
    both = inside[contacts["proxyA"]] & inside[contacts["proxyB"]]
    got = set(T.contact_keys(contacts[both]).tolist())
    # a new overlap gets its contact in the same step; a contact outlives the overlap of its boxes by at most one step
    assert want <= got, (len(want), len(got), len(want - got))
    assert len(got - want) <= max(8, len(got) // 20), (len(want), len(got))


def test_joint_chains_at_scale():
    """40,000 revolute joints (1000 hanging chains of 40 links that swing into each other and onto the floor): two
    device worlds agree bit for bit, the joint order is a proper colouring (no two joints of one class share a dynamic
    body), and the hinges hold (the reference itself leaves up to 0.045 of anchor error on a swinging 40-link chain with
    3 position iterations; here chains also hit each other)."""
    scene = scenes.hanging_chains(1000, 40)
    a = b2host.HostWorld(scene, download_bodies=False, events=False)
    b = b2host.HostWorld(scene, download_bodies=False, events=False)
    for _ in range(120):
        a.step()
        b.step()
    ba, bb = a.bodies(), b.bodies()
    assert ba.tobytes() == bb.tobytes()
    ja, jb = a.device_world().get_joints(), b.device_world().get_joints()
    assert len(ja) == 40000 and ja.tobytes() == jb.tobytes()

    # colouring: walk the order, a class ends where a dynamic body would repeat
    order = a.device_world().joint_order()
    assert sorted(order.tolist()) == list(range(len(ja)))
    dynamic = (ba["flags"] & T.BODY_TYPE_MASK) == T.DYNAMIC_BODY
    seen, classes = set(), 1
    for j in order:
        ends = [int(x) for x in (ja["bodyA"][j], ja["bodyB"][j]) if dynamic[x]]
        if any(e in seen for e in ends):
            seen, classes = set(), classes + 1
        seen.update(ends)
    assert classes <= 3   # a chain needs two classes; ids inside a class ascend, so at most one extra break

    # hinges hold: world anchors of both sides coincide
    def world(anchor, body):
        s, c = ba["qs"][body], ba["qc"][body]
        return np.stack([ba["px"][body] + c * anchor[:, 0] - s * anchor[:, 1], ba["py"][body] + s * anchor[:, 0] + c * anchor[:, 1]], 1)

    gap = np.linalg.norm(world(ja["localAnchorA"], ja["bodyA"]) - world(ja["localAnchorB"], ja["bodyB"]), axis=1)
    assert gap.max() < 0.15, gap.max()   # well inside a link's half-thickness
    assert np.abs(ja["impulse"][:, :2]).max() > 1.0   # the joints do carry load


# ---- the other BASELINE.json configurations at their stated sizes, in lockstep with the oracle ---------------------------
# (the oracle manages these sizes at 0.05 - 0.5 s per step, so a few dozen steps are affordable)

import parity  # noqa: E402
import ref  # noqa: E402


def _lockstep(scene, steps, check_every):
    import b2cuda
    r = ref.RefWorld(scene)
    g = parity.gpu_world_from_ref(b2cuda, r)
    return parity.lockstep(g, r, steps, check_every=check_every), g, r


def test_add_pair_10k_pair_set_bit_exact():
    """BASELINE configs[1]: 10,000 small circles and one fast bullet box.  The broad-phase pair set, the manifolds, the
    events and the bodies are compared with the oracle while the box ploughs through the cloud."""
    # 14 steps: the box reaches the cloud at step 5; later the compressed cloud has millions of contacts and the oracle
    # needs seconds per step
    infos, g, r = _lockstep(scenes.add_pair(10000), 14, check_every=2)
    assert max(int(i["contactCount"]) for i in infos) > 200000
    assert (np.sort(T.contact_keys(g.get_contacts())) == np.sort(T.contact_keys(r.contacts()))).all()


def test_tumbler_20k_lockstep():
    """BASELINE configs[2]: 20,000 small boxes in the rotating kinematic drum."""
    infos, g, r = _lockstep(scenes.tumbler(20000), 24, check_every=4)
    assert int(infos[-1]["contactCount"]) > 50000


def test_stacks_100k_lockstep():
    """BASELINE configs[3]: 476 pyramids of 20 rows (99,960 boxes) on thick static ground, sleeping enabled."""
    # the boxes start a quarter of a metre above their rests: first contacts after 14 steps
    infos, g, r = _lockstep(scenes.pyramids(476, 20, thick_polygon_ground=True), 30, check_every=5)
    assert int(infos[-1]["constraintCount"]) > 50000
