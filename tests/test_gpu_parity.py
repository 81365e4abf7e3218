"""GPU parity tests: the CUDA path, called through the C ABI, against the oracle (the compiled reference).

Bars (BASELINE.json north_star): broad-phase pair set, begin/end events, manifold point counts and feature ids
bit-exact; manifold geometry and single-step body state within 1e-5 relative.  Because the device keeps the
reference's fp32 operation order (no FMA, IEEE div/sqrt, one shared sincos), these tests assert the stronger
property: bit-exact everywhere (tol = 0.0), single-step and free-running.
"""
import numpy as np
import pytest

import b2cuda_types as T
import parity
import ref
import scenes

pytestmark = pytest.mark.gpu

TOL = 0.0  # bit-exact; the contract would allow 1e-5 relative


def test_sincos_bit_exact(gpu):
    rng = np.random.default_rng(7)
    xs = np.concatenate([rng.uniform(-10, 10, 20000), rng.uniform(-3000, 3000, 5000), rng.normal(0, 1e-3, 2000),
                         [0.0, -0.0, np.pi, -np.pi / 2, 1e-20]]).astype(np.float32)
    s, c = gpu.sincos(xs)
    for i in range(0, len(xs), 7):
        rs, rc = ref.sincos(xs[i])
        assert s[i] == rs and c[i] == rc, (xs[i], s[i], rs, c[i], rc)
    assert (s == np.sin(xs.astype(np.float64)).astype(np.float32)).all()
    assert (c == np.cos(xs.astype(np.float64)).astype(np.float32)).all()


def _shape_table():
    """polygons (box, regular 3-8-gons, a skewed quad), circles, edges with and without ghost vertices"""
    s = scenes.Scene()
    g = s.body(T.STATIC_BODY, (0, 0))
    ids = [s.box(0.5, 0.5), s.box(1.0, 0.25), s.box(0.3, 0.7, center=(0.2, -0.1), angle=0.4)]
    ids += [s.polygon(scenes._regular_polygon(k, 0.5)) for k in range(3, 9)]
    ids += [s.polygon([(-0.5, -0.2), (0.7, -0.4), (0.6, 0.5), (-0.3, 0.4)])]
    ids += [s.circle(0.5), s.circle(0.2, p=(0.1, 0.05))]
    ids += [s.edge((-1, 0), (1, 0)), s.edge((-1, 0), (1, 0.2), v0=(-2, 0.5), v3=(2, -0.3)),
            s.edge((-1, 0), (1, 0), v0=(-2, -0.5)), s.edge((-1, 0), (1, 0), v3=(2, 0.6))]
    for i in ids:
        s.fixture(g, i)
    w = ref.RefWorld(s)
    return w.shapes()


def test_collide_pairs_bit_exact(gpu):
    shapes = _shape_table()
    rng = np.random.default_rng(3)
    types = shapes["type"]
    n = 6000
    a = rng.integers(0, len(shapes), n)
    b = rng.integers(0, len(shapes), n)
    # primary type order (b2Contact.cpp:44-50): polygon/edge before circle, edge before polygon; no edge-edge
    rank = {T.SHAPE_EDGE: 0, T.SHAPE_POLYGON: 1, T.SHAPE_CIRCLE: 2}
    keep = []
    for i in range(n):
        ta, tb = types[a[i]], types[b[i]]
        if ta == T.SHAPE_EDGE and tb == T.SHAPE_EDGE:
            continue
        if rank[ta] > rank[tb]:
            a[i], b[i] = b[i], a[i]
        keep.append(i)
    a, b = a[keep], b[keep]
    n = len(a)
    ang_a = rng.uniform(-np.pi, np.pi, n).astype(np.float32)
    ang_b = rng.uniform(-np.pi, np.pi, n).astype(np.float32)
    xa = np.zeros((n, 4), np.float32)
    xb = np.zeros((n, 4), np.float32)
    xa[:, :2] = rng.uniform(-0.3, 0.3, (n, 2))
    xb[:, :2] = xa[:, :2] + rng.normal(0, 0.6, (n, 2))
    xa[:, 2], xa[:, 3] = np.sin(ang_a.astype(np.float64)), np.cos(ang_a.astype(np.float64))
    xb[:, 2], xb[:, 3] = np.sin(ang_b.astype(np.float64)), np.cos(ang_b.astype(np.float64))
    # axis-aligned and exactly touching cases too
    xa[:200, 2:] = (0, 1)
    xb[:200, 2:] = (0, 1)
    got = gpu.collide_pairs(shapes, a, xa, b, xb)
    touching = 0
    for i in range(n):
        want = ref.collide(shapes[a[i]], xa[i], shapes[b[i]], xb[i])
        assert got[i]["pointCount"] == want["pointCount"], i
        pc = int(want["pointCount"])
        if pc == 0:
            continue
        touching += 1
        assert got[i]["type"] == want["type"], i
        assert (got[i]["id"][:pc] == want["id"][:pc]).all(), i
        parity.assert_floats_equal("localNormal", got[i]["localNormal"], want["localNormal"], TOL)
        parity.assert_floats_equal("localPoint", got[i]["localPoint"], want["localPoint"], TOL)
        parity.assert_floats_equal("points", got[i]["points"]["localPoint"][:pc], want["points"]["localPoint"][:pc], TOL)
    assert touching > n // 10


def test_distance_pairs_bit_exact(gpu):
    """b2Distance (GJK) on the device against the reference's, on random pairs of every shape class, with and without
    radii: distance, witness points and iteration count (b2Distance.cpp:452-603)."""
    shapes = _shape_table()
    rng = np.random.default_rng(11)
    n = 4000
    a = rng.integers(0, len(shapes), n)
    b = rng.integers(0, len(shapes), n)
    ang_a = rng.uniform(-np.pi, np.pi, n)
    ang_b = rng.uniform(-np.pi, np.pi, n)
    xa = np.zeros((n, 4), np.float32)
    xb = np.zeros((n, 4), np.float32)
    xa[:, :2] = rng.uniform(-0.5, 0.5, (n, 2))
    xb[:, :2] = xa[:, :2] + rng.normal(0, 1.0, (n, 2))
    xa[:, 2], xa[:, 3] = np.sin(ang_a), np.cos(ang_a)
    xb[:, 2], xb[:, 3] = np.sin(ang_b), np.cos(ang_b)
    xb[:300] = xa[:300]          # coincident frames: deep overlap, degenerate simplices
    xb[300:600, :2] = xa[300:600, :2] + rng.normal(0, 0.05, (300, 2))
    overlapping = 0
    for use_radii in (True, False):
        got = gpu.distance_pairs(shapes, a, xa, b, xb, use_radii)
        for i in range(n):
            want = ref.distance(shapes[a[i]], xa[i], shapes[b[i]], xb[i], use_radii)
            assert got[i]["iterations"] == want["iterations"], (i, use_radii)
            parity.assert_floats_equal("distance", got[i]["distance"], want["distance"], TOL)
            parity.assert_floats_equal("pointA", got[i]["pointA"], want["pointA"], TOL)
            parity.assert_floats_equal("pointB", got[i]["pointB"], want["pointB"], TOL)
            overlapping += int(want["distance"] < 10 * np.finfo(np.float32).eps)
    assert overlapping > n // 10


def _random_sweeps(rng, n, start, spread, turn):
    """n b2Sweep records: from around `start` towards a random point `spread` away, rotating up to `turn` radians."""
    sw = np.zeros(n, T.SWEEP)
    sw["localCenter"] = rng.uniform(-0.2, 0.2, (n, 2))
    sw["c0"] = start + rng.uniform(-0.3, 0.3, (n, 2))
    sw["c"] = sw["c0"] + rng.normal(0, spread, (n, 2))
    sw["a0"] = rng.uniform(-8.0, 8.0, n)        # outside [0, 2 pi): exercises b2Sweep::Normalize
    sw["a"] = sw["a0"] + rng.uniform(-turn, turn, n)
    return sw


def test_time_of_impact_bit_exact(gpu):
    """b2TimeOfImpact on the device against the reference's (b2TimeOfImpact.cpp:256-497): state and time of random
    sweeps of every shape class -- approaching, passing, spinning in place, starting in touch or overlap."""
    shapes = _shape_table()
    rng = np.random.default_rng(23)
    n = 6000
    a = rng.integers(0, len(shapes), n)
    b = rng.integers(0, len(shapes), n)
    sa = _random_sweeps(rng, n, np.zeros(2), 0.5, 1.0)
    sb = _random_sweeps(rng, n, np.array([2.5, 0.0]), 0.5, 1.0)
    # B heads for A: most of these hit
    sb["c"][:4000] = sa["c"][:4000] + rng.normal(0, 0.4, (4000, 2))
    # fast spinners (the root finder's hard case) and sweeps that start in overlap
    sb["a"][1000:1500] = sb["a0"][1000:1500] + rng.uniform(-6.0, 6.0, 500)
    sb["c0"][1500:1800] = sa["c0"][1500:1800] + rng.normal(0, 0.1, (300, 2))
    # second halves of a step (alpha0 > 0 does not enter the routine, but travels with the record)
    sa["alpha0"][2000:2500] = rng.uniform(0, 0.9, 500)
    t_max = np.ones(n, np.float32)
    t_max[2500:3000] = rng.uniform(0.1, 1.0, 500)
    got = gpu.time_of_impact_pairs(shapes, a, sa, b, sb, t_max)
    states = np.zeros(5, int)
    for i in range(n):
        want = ref.time_of_impact(shapes[a[i]], sa[i], shapes[b[i]], sb[i], float(t_max[i]))
        assert got[i]["state"] == want["state"], (i, got[i], want)
        parity.assert_floats_equal("t", got[i]["t"], want["t"], TOL)
        states[int(want["state"])] += 1
    assert states[T.TOI_TOUCHING] > n // 5 and states[T.TOI_SEPARATED] > n // 10 and states[T.TOI_OVERLAPPED] > 50, states


SCENES = {
    "pyramid6": lambda: scenes.pyramid(6),
    "pyramid20": lambda: scenes.pyramid(20),
    "pile": lambda: scenes.pile(12, 10),
    "pile_5000": lambda: scenes.pile(100, 50),
    "pile_sleep": lambda: scenes.pile(8, 6, sleep=True),
    "tumbler": lambda: scenes.tumbler(100),
    "chains": lambda: scenes.chain_terrain(40),
    "sensors": lambda: scenes.sensors(30),
    "two_pyramids": lambda: scenes.pyramids(2, 6, thick_polygon_ground=True),
    "tumbler_joint": lambda: scenes.tumbler(100, motor_joint=True),
    "hanging_chains": lambda: scenes.hanging_chains(4, 12),
    "joint_zoo": lambda: scenes.joint_zoo(),
    "rods_and_welds": lambda: scenes.rods_and_welds(),
    "sliders": lambda: scenes.sliders(),
    "machines": lambda: scenes.machines(),
    "pulleys_and_mice": lambda: scenes.pulleys_and_mice(),
    "gears": lambda: scenes.gears(),
}


@pytest.mark.parametrize("name", sorted(SCENES))
def test_single_step_teacher_forced(gpu, name):
    """Single-step parity from identical state: before every step the oracle's bodies, fat AABBs
    and contacts (manifolds, impulses) are injected into the device; after the step everything must be equal."""
    scene = SCENES[name]()
    r = ref.RefWorld(scene)
    g = parity.gpu_world_from_ref(gpu, r)
    infos = parity.lockstep(g, r, 60, teacher=True, tol=TOL)
    assert max(i["constraintCount"] for i in infos) > 0


@pytest.mark.parametrize("name,steps", [("pyramid6", 300), ("pyramid20", 240), ("pile", 300), ("pile_5000", 200),
                                        ("pile_sleep", 400), ("chains", 300), ("sensors", 300),
                                        ("tumbler", 200), ("two_pyramids", 300), ("tumbler_joint", 200),
                                        ("hanging_chains", 400), ("joint_zoo", 400), ("rods_and_welds", 400), ("sliders", 400), ("machines", 400), ("pulleys_and_mice", 400), ("gears", 400)])
def test_free_running_lockstep(gpu, name, steps):
    """Multi-step parity: the device world runs freely; the oracle follows in the GPU's solver order.  Pair set,
    events, manifolds, fat AABBs, body state and awake flags are compared after every step."""
    scene = SCENES[name]()
    r = ref.RefWorld(scene)
    g = parity.gpu_world_from_ref(gpu, r)
    infos = parity.lockstep(g, r, steps, tol=TOL)
    assert sum(int(i["beginCount"]) for i in infos) > 0
    if name == "pile_sleep":
        # sleeping is exercised: every body asleep at the end, in both worlds
        assert infos[-1]["awakeBodyCount"] == 0
        assert not (r.bodies()["flags"][1:] & T.BODY_AWAKE).any()


@pytest.mark.parametrize("compact_min", [1, 24])
def test_lockstep_with_frequent_compaction(gpu, compact_min, monkeypatch):
    """The contact set is compacted lazily (dead flags + a small sorted tail, merged into the main region now and
    then).  Force the merge to run (almost) every step and check that nothing observable changes."""
    monkeypatch.setenv("B2CU_COMPACT_MIN", str(compact_min))
    for name in ("pile", "tumbler"):
        scene = SCENES[name]()
        r = ref.RefWorld(scene)
        g = parity.gpu_world_from_ref(gpu, r)
        parity.lockstep(g, r, 200, tol=TOL)
    scene = scenes.add_pair(300)
    r = ref.RefWorld(scene)
    g = parity.gpu_world_from_ref(gpu, r)
    parity.lockstep(g, r, 60, tol=TOL)


TOI_SCENES = {
    "hello": (scenes.hello_world, 90),              # one landing: one event
    "pyramid20": (lambda: scenes.pyramid(20), 150),  # the Testbed's Pyramid on its edge ground, reference defaults
    "bullets": (scenes.bullets, 240),                # bullets and fast bodies against thin walls and each other
    "add_pair": (lambda: scenes.add_pair(300), 90),  # the bullet box through a cloud of circles
    "chains": (lambda: scenes.chain_terrain(40), 200),
}


@pytest.mark.parametrize("name", sorted(TOI_SCENES))
def test_time_of_impact_sub_steps(gpu, name):
    """b2World::SolveTOI (b2World.cpp:1026-1093) with continuous physics ON on both sides: every FindMinToiContact pass,
    every time-of-impact event (StepSolveTOI :851-1024: the two bodies advanced, the 32-contact island taken from the
    bodies' contact lists in creation order, b2Island::SolveTOI, SynchronizeFixtures, FindNewContacts) and the final
    ClearPostSolveTOI.  The oracle runs the reference's own SolveTOI; bodies, contacts (manifolds, impulses, flags),
    fat boxes and the begin / end callbacks -- those of the sub-steps in call order -- are compared after every step."""
    make, steps = TOI_SCENES[name]
    scene = make()
    assert scene.world_flags & T.WORLD_CONTINUOUS
    r = ref.RefWorld(scene)
    g = parity.gpu_world_from_ref(gpu, r)
    infos = parity.lockstep(g, r, steps, tol=TOL)
    sub = sum(int(i["toiSubSteps"]) for i in infos)
    assert sub > 0, "the scene was meant to produce time-of-impact events"
    assert max(int(i["toiCandidateCount"]) for i in infos) > 0
    if name in ("bullets", "add_pair"):
        # events that create contacts (the per-event FindNewContacts) and raise callbacks inside the sub-steps
        assert sum(int(i["toiNewContactCount"]) for i in infos) > 0
        assert sum(int(i["toiEventCount"]) for i in infos) > 0
        # a contact can be the earliest more than once per step; the reference stops at b2_maxSubSteps (b2Contact.h:412-416)
        assert max(int(i["toiSubSteps"]) for i in infos) > 8
    # after the step the time-of-impact bookkeeping is reset (ClearPostSolveTOI): alpha0 = 0, toi = 1, toiCount = 0
    assert (g.get_bodies()["alpha0"] == 0).all()
    c = g.get_contacts()
    assert (c["toi"] == 1).all() and (c["toiCount"] == 0).all() and not (c["flags"] & (T.CONTACT_TOI | T.CONTACT_ISLAND)).any()


@pytest.mark.parametrize("name", ["bullets", "pyramid6", "add_pair"])
def test_time_of_impact_teacher_forced(gpu, name):
    """Single-step parity with continuous physics on: the oracle's state (bodies, proxies, contacts WITH their creation
    stamps, which fix the order of the bodies' contact lists) is injected before every step."""
    make = {"bullets": scenes.bullets, "pyramid6": lambda: scenes.pyramid(6), "add_pair": lambda: scenes.add_pair(200)}[name]
    r = ref.RefWorld(make())
    g = parity.gpu_world_from_ref(gpu, r)
    infos = parity.lockstep(g, r, 120, teacher=True, tol=TOL)
    assert sum(int(i["toiSubSteps"]) for i in infos) > 0


def test_sub_stepping_world(gpu):
    """b2World::SetSubStepping(true): one time-of-impact event per Step; while events remain the next Step skips the
    solve (m_stepComplete, b2World.cpp:1670, :1084-1088)."""
    scene = scenes.bullets()
    scene.world_flags |= T.WORLD_SUB_STEPPING
    r = ref.RefWorld(scene)
    g = parity.gpu_world_from_ref(gpu, r)
    infos = parity.lockstep(g, r, 300, tol=TOL)
    assert max(int(i["toiSubSteps"]) for i in infos) == 1
    assert sum(int(i["toiEventPending"]) for i in infos) > 10


def test_add_pair_pair_set(gpu):
    """BASELINE config 2 (Add Pair, scaled down): broad-phase stress with a fast bullet box; pair set bit-exact
    every step, through the bullet's impact, with continuous physics on."""
    scene = scenes.add_pair(400)
    r = ref.RefWorld(scene)
    g = parity.gpu_world_from_ref(gpu, r)
    infos = parity.lockstep(g, r, 90, tol=TOL)
    assert max(int(i["contactCount"]) for i in infos) > 1000


def test_deterministic_repeat(gpu):
    """Two device worlds built from the same scene stay bit-identical (the reference's reproducibility claim,
    README.md:161-173, restated for the GPU path)."""
    scene = scenes.pile(16, 12)
    r = ref.RefWorld(scene)
    a = parity.gpu_world_from_ref(gpu, r)
    b = parity.gpu_world_from_ref(gpu, r)
    for _ in range(120):
        a.step()
        b.step()
    ba, bb = a.get_bodies(), b.get_bodies()
    assert ba.tobytes() == bb.tobytes()
    assert a.get_contacts().tobytes() == b.get_contacts().tobytes()


def test_invariants_long_run(gpu):
    """1000-step invariants on a resting stack (Testbed/Tests/SleepCollideTest.h:104-110 rule: no body below the
    ground; bounded penetration; the pyramid stays standing)."""
    scene = scenes.pyramid(10)
    r = ref.RefWorld(scene)
    g = parity.gpu_world_from_ref(gpu, r)
    for _ in range(1000):
        g.step()
        r.step()  # the reference in its own (DFS) Gauss-Seidel order: a different but equally valid solve
    b = g.get_bodies()
    rb = r.bodies()
    assert (b["py"][1:] > 0.45).all()                       # nothing sinks into or falls below the ground
    assert np.hypot(b["vx"], b["vy"]).max() < 0.05          # the stack is at rest (momentum drift bounded)
    assert np.abs(b["py"] - rb["py"]).max() < 0.02          # and stands where the reference's stands
    assert np.abs(b["px"] - rb["px"]).max() < 0.05
    c = g.get_contacts()
    assert (c["manifold"]["pointCount"] <= 2).all()
    assert len(c) == r.counts()[2]
