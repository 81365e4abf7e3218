"""Edge cases of the step, device vs oracle in lockstep (bit-exact), through the C ABI.

Covers what the reference's own small tests and Testbed scenes exercise around the hot path besides stacks and
piles: empty and contact-free worlds, dt = 0, every contact class against edge ground (ghost vertices included),
restitution (the velocity-bias branch, b2ContactSolver.cpp:213-217), zero / high friction, damping, gravity scale,
fixed rotation, kinematic bodies, collision filtering (category / mask / group, b2WorldCallbacks.cpp:24-38),
several fixtures per body, warm starting off, sleeping off, a changing time step (dtRatio, b2ContactSolver.cpp:113).
"""
import numpy as np
import pytest

import b2cuda_types as T
import b2scene
import parity
import ref
import scenes

pytestmark = pytest.mark.gpu


def _scene(**kw):
    s = scenes.Scene(**kw)
    return s


def _run(gpu, scene, steps, **kw):
    r = ref.RefWorld(scene)
    g = parity.gpu_world_from_ref(gpu, r)
    return parity.lockstep(g, r, steps, tol=0.0, **kw), g, r


def test_empty_world(gpu):
    infos, g, r = _run(gpu, _scene(), 3)
    assert int(infos[-1]["bodyCount"]) == 0 and int(infos[-1]["contactCount"]) == 0


def test_bodies_without_fixtures_and_no_contacts(gpu):
    s = _scene()
    s.body(T.STATIC_BODY, (0, 0))
    s.body(T.DYNAMIC_BODY, (0, 5))                     # no fixture: unit mass, falls
    b = s.body(T.DYNAMIC_BODY, (3, 5), vel=(1.0, 2.0), w=0.5, linear_damping=0.3, angular_damping=0.2, gravity_scale=0.5)
    s.fixture(b, s.circle(0.5), density=1.0)
    infos, g, r = _run(gpu, s, 60)
    assert int(infos[-1]["contactCount"]) == 0
    assert g.get_bodies()["py"][1] < 0.0  # it really fell


def test_zero_time_step(gpu):
    s = _scene()
    ground = s.body(T.STATIC_BODY, (0, 0))
    s.fixture(ground, s.box(10, 0.5, center=(0, -0.5)), thick=True)
    b = s.body(T.DYNAMIC_BODY, (0, 0.4))
    s.fixture(b, s.box(0.5, 0.5), density=1.0)
    r = ref.RefWorld(s)
    g = parity.gpu_world_from_ref(gpu, r)
    parity.lockstep(g, r, 5, tol=0.0)
    before = g.get_bodies().copy()
    parity.lockstep(g, r, 2, dt=0.0, tol=0.0)          # b2World::Step with dt = 0: collide only
    after = g.get_bodies()
    assert (before["px"] == after["px"]).all() and (before["py"] == after["py"]).all()
    parity.lockstep(g, r, 10, tol=0.0)                 # and the warm-start ratio after a zero step (inv_dt0 = 0)


def _edge_ground(s):
    """a chain of three edges with ghost vertices (the adjacency b2EPCollider uses), slightly V-shaped"""
    g = s.body(T.STATIC_BODY, (0, 0))
    pts = [(-12.0, 1.0), (-4.0, 0.0), (4.0, 0.0), (12.0, 1.5)]
    s.fixture(g, s.edge(pts[0], pts[1], v3=pts[2]))
    s.fixture(g, s.edge(pts[1], pts[2], v0=pts[0], v3=pts[3]))
    s.fixture(g, s.edge(pts[2], pts[3], v0=pts[1]))
    return g


def test_every_shape_on_edge_ground(gpu):
    s = _scene()
    _edge_ground(s)
    rng = np.random.default_rng(5)
    shapes = [s.circle(0.3), s.box(0.4, 0.25)] + [s.polygon(scenes._regular_polygon(k, 0.35)) for k in range(3, 9)]
    for i in range(40):
        b = s.body(T.DYNAMIC_BODY, (-9.0 + 0.45 * i, 2.0 + 0.7 * (i % 3)), angle=float(rng.uniform(0, 6.28)))
        s.fixture(b, shapes[i % len(shapes)], density=1.0, friction=0.3)
    infos, g, r = _run(gpu, s, 300)
    assert sum(int(i["beginCount"]) for i in infos) > 40


def test_restitution_friction_and_fixed_rotation(gpu):
    s = _scene()
    ground = s.body(T.STATIC_BODY, (0, 0))
    s.fixture(ground, s.box(20, 0.5, center=(0, -0.5), angle=0.05), thick=True, friction=0.6)
    for i, (rest, fric) in enumerate([(0.0, 0.0), (0.3, 1.0), (0.8, 0.2), (1.0, 0.5)]):
        b = s.body(T.DYNAMIC_BODY, (-6.0 + 4.0 * i, 4.0), vel=(1.0, -3.0))
        s.fixture(b, s.circle(0.4), density=1.0, restitution=rest, friction=fric)
        c = s.body(T.DYNAMIC_BODY, (-4.5 + 4.0 * i, 3.0), angle=0.3,
                   flags=b2scene.BODYDEF_DEFAULT | (b2scene.BODYDEF_FIXED_ROTATION if i % 2 else 0))
        s.fixture(c, s.box(0.4, 0.3), density=2.0, restitution=rest, friction=fric)
    infos, g, r = _run(gpu, s, 240)
    assert sum(int(i["endCount"]) for i in infos) > 0  # the bouncing ones leave the ground again


def test_kinematic_platform_and_compound_bodies(gpu):
    s = _scene()
    ground = s.body(T.STATIC_BODY, (0, 0))
    s.fixture(ground, s.box(30, 0.5, center=(0, -0.5)), thick=True)
    lift = s.body(T.KINEMATIC_BODY, (0, 1.0), vel=(0.5, 0.2), w=0.1)
    s.fixture(lift, s.box(3.0, 0.2))
    for i in range(12):
        b = s.body(T.DYNAMIC_BODY, (-2.5 + 0.45 * i, 2.0 + 0.5 * (i % 2)))
        # two fixtures per body: an L of two boxes, and a box with a circle on top
        s.fixture(b, s.box(0.2, 0.1), density=1.0)
        if i % 2:
            s.fixture(b, s.box(0.05, 0.2, center=(0.15, 0.25)), density=1.0)
        else:
            s.fixture(b, s.circle(0.12, p=(0.0, 0.2)), density=3.0)
    infos, g, r = _run(gpu, s, 200)
    assert max(int(i["constraintCount"]) for i in infos) > 10


def test_collision_filtering(gpu):
    s = _scene()
    ground = s.body(T.STATIC_BODY, (0, 0))
    s.fixture(ground, s.box(20, 0.5, center=(0, -0.5)), thick=True)
    # same negative group never collide; same positive group always; otherwise category & mask both ways
    specs = [dict(group=-1), dict(group=-1), dict(group=2, category=0x2, mask=0x1), dict(group=2, category=0x4, mask=0x1),
             dict(category=0x2, mask=0xFFFD), dict(category=0x2, mask=0xFFFF), dict(category=0x8, mask=0x1)]
    for i, f in enumerate(specs * 3):
        b = s.body(T.DYNAMIC_BODY, (0.05 * (i % 7), 0.6 + 0.55 * i))
        s.fixture(b, s.box(0.4, 0.25), density=1.0, **f)
    infos, g, r = _run(gpu, s, 240)
    assert int(infos[-1]["contactCount"]) > 0


@pytest.mark.parametrize("flags", [T.WORLD_DEFAULT & ~T.WORLD_WARM_STARTING,
                                   T.WORLD_DEFAULT & ~T.WORLD_ALLOW_SLEEP,
                                   T.WORLD_DEFAULT & ~T.WORLD_CLEAR_FORCES])
def test_world_switches(gpu, flags):
    s = scenes.pile(6, 5, sleep=True)
    s.world_flags = flags
    infos, g, r = _run(gpu, s, 200)
    assert max(int(i["constraintCount"]) for i in infos) > 0


def test_varying_time_step_and_iterations(gpu):
    s = scenes.pyramid(5)
    r = ref.RefWorld(s)
    g = parity.gpu_world_from_ref(gpu, r)
    for dt, vi, pi in [(1 / 60.0, 8, 3), (1 / 30.0, 4, 1), (1 / 120.0, 10, 4), (1 / 60.0, 1, 0), (1 / 45.0, 6, 2)]:
        parity.lockstep(g, r, 25, dt=dt, vel_iters=vi, pos_iters=pi, tol=0.0)


def test_custom_pair_filter(gpu):
    """A user b2ContactFilter replaces the default rule for new pairs (b2ContactManager.cpp:280-285): the oracle gets
    a b2ContactFilter subclass, the device the same rule through b2cuSetPairFilter."""
    s = scenes.pile(10, 8)
    r = ref.RefWorld(s)
    r.set_modulo_filter(5)
    g = parity.gpu_world_from_ref(gpu, r)
    calls = []

    def rule(keys):
        calls.append(len(keys))
        a = (keys >> np.uint64(32)).astype(np.int64)
        b = (keys & np.uint64(0xFFFFFFFF)).astype(np.int64)
        return (a + b) % 5 != 0

    g.set_pair_filter(rule)
    infos = parity.lockstep(g, r, 200, tol=0.0)
    assert sum(calls) > 0
    keys = T.contact_keys(g.get_contacts())
    assert (((keys >> np.uint64(32)).astype(np.int64) + (keys & np.uint64(0xFFFFFFFF)).astype(np.int64)) % 5 != 0).all()
    # bodies really pass through each other where the rule says so: more overlap than the default world would allow
    assert int(infos[-1]["contactCount"]) > 0


def test_more_than_32_contacts_on_one_body(gpu):
    """A heavy disc buried in small ones has more touching contacts than there are colours (32): the surplus goes to
    the serial overflow list of the solver, which must still reproduce the sequential Gauss-Seidel of the oracle."""
    s = _scene()
    ground = s.body(T.STATIC_BODY, (0, 0))
    s.fixture(ground, s.box(12, 0.5, center=(0, -0.5)), thick=True)
    s.fixture(ground, s.box(0.5, 6, center=(-5.5, 6)), thick=True)
    s.fixture(ground, s.box(0.5, 6, center=(5.5, 6)), thick=True)
    big = s.body(T.DYNAMIC_BODY, (0.0, 3.2))
    s.fixture(big, s.circle(3.0), density=0.2, friction=0.3)
    small = s.circle(0.12)
    k = 0
    for row in range(14):
        for col in range(40):
            x, y = -4.9 + 0.25 * col + 0.01 * (row % 2), 6.6 + 0.25 * row
            b = s.body(T.DYNAMIC_BODY, (x, y))
            s.fixture(b, small, density=1.0, friction=0.2)
            k += 1
    r = ref.RefWorld(s)
    g = parity.gpu_world_from_ref(gpu, r)
    infos = parity.lockstep(g, r, 220, tol=0.0, check_every=4)
    assert max(int(i["overflowCount"]) for i in infos) > 0, "the scene must exhaust the colours"
    assert max(int(i["colourCount"]) for i in infos) >= 30


def test_proxy_larger_than_the_coarsest_grid_level(gpu):
    """Tiny bodies make the finest grid cell tiny; a ground box a million times larger than them is beyond the coarsest
    of the 24 grid levels and takes the `huge` list path of the broad-phase.  Pair set and state stay exact."""
    s = _scene()
    ground = s.body(T.STATIC_BODY, (0, 0))
    s.fixture(ground, s.box(40000.0, 0.5, center=(0, -0.5)), thick=True)
    tiny = s.circle(0.004)
    box = s.box(0.004, 0.003)
    for i in range(60):
        b = s.body(T.DYNAMIC_BODY, (-0.3 + 0.01 * i, 0.02 + 0.011 * (i % 5)))
        s.fixture(b, tiny if i % 2 else box, density=1.0)
    mover = s.body(T.KINEMATIC_BODY, (0.0, 0.3), vel=(0.0, -0.05))
    s.fixture(mover, s.box(30000.0, 0.05))          # a huge MOVING proxy as well
    infos, g, r = _run(gpu, s, 200)
    assert int(infos[-1]["contactCount"]) > 60


def test_capacity_growth_from_nothing(gpu):
    """Device buffers grow geometrically on their own: a world created with minimal capacities receives a scene, and the
    contact set outgrows its buffer during the run (b2cuSetCounts, the contact capacity inside b2cuStep)."""
    s = scenes.pile(20, 12)
    r = ref.RefWorld(s)
    bodies, shapes, proxies, contacts = parity.ref_state(r)
    g = gpu.World(gravity=r.gravity, flags=r.world_flags, body_capacity=1, proxy_capacity=1, shape_capacity=1,
                  contact_capacity=1)
    g.load_state(bodies, shapes, proxies, contacts, inv_dt0=r.inv_dt0())
    infos = parity.lockstep(g, r, 150, tol=0.0, check_every=3)
    assert int(infos[-1]["contactCount"]) > 1000



def test_joint_table_edge_cases(gpu):
    """b2cuSetJoints: bad rows are refused with a message and leave the table as it was; an empty table removes every
    joint; a joint between two static bodies is accepted and never solved; replacing the table keeps stepping exact."""
    s = _scene()
    g0 = s.body(T.STATIC_BODY, (0, 0))
    s.fixture(g0, s.box(20, 0.5, center=(0, -0.5)), thick=True)
    g1 = s.body(T.STATIC_BODY, (0, 10))
    a = s.body(T.DYNAMIC_BODY, (0.0, 5.0))
    s.fixture(a, s.box(0.5, 0.5), density=1.0)
    b = s.body(T.DYNAMIC_BODY, (2.0, 5.0))
    s.fixture(b, s.circle(0.5), density=1.0)
    s.revolute_joint(g1, a, (0.0, 0.0), (0.0, 5.0))
    s.distance_joint(a, b, (0.0, 0.0), (0.0, 0.0), 2.0)
    s.weld_joint(g0, g1, (0.0, 10.0), (0.0, 0.0))           # static - static: in no island, never solved
    r = ref.RefWorld(s)
    g = parity.gpu_world_from_ref(gpu, r)
    parity.lockstep(g, r, 40, tol=0.0)
    assert (g.get_joints()["impulse"][2] == 0.0).all()

    good = g.get_joints()
    for field, value, code in (("bodyA", 99, -3), ("bodyB", -1, -3), ("type", 12, -4), ("type", 0, -4)):
        bad = good.copy()
        bad[field][1] = value
        with pytest.raises(gpu.B2cuError) as err:
            g.set_joints(bad)
        assert err.value.code == code, (field, value)
    same = good.copy()
    same["bodyB"][1] = same["bodyA"][1]                       # a joint from a body to itself
    with pytest.raises(gpu.B2cuError):
        g.set_joints(same)
    gear = good[:1].copy()
    gear["type"] = T.JOINT_GEAR
    gear["limitState"] = 77                                   # body C out of range
    with pytest.raises(gpu.B2cuError):
        g.set_joints(gear)
    # the refused tables changed nothing
    assert g.get_joints().tobytes() == good.tobytes()
    parity.lockstep(g, r, 20, tol=0.0)

    # cut the rod on both sides, then remove everything
    keep = g.get_joints()[[0, 2]]
    g.set_joints(keep)
    r.destroy_joint(1)
    parity.lockstep(g, r, 40, tol=0.0)
    g.set_joints(keep[:0])
    r.destroy_joint(1)
    r.destroy_joint(0)
    assert len(g.get_joints()) == 0 and len(g.joint_order()) == 0
    parity.lockstep(g, r, 40, tol=0.0)


@pytest.mark.parametrize("name", ["joint_zoo", "machines", "gears", "rods_and_welds", "sliders", "pulleys_and_mice"])
def test_joints_with_varying_dt_and_without_warm_starting(gpu, name):
    """The joints' warm start scales the stored impulses by dt / dt0 (the gear joint, alone, does not), soft constraints
    and motors use dt and 1 / dt directly: step with a changing dt, then with warm starting switched off."""
    scene = getattr(scenes, name)()
    r = ref.RefWorld(scene)
    g = parity.gpu_world_from_ref(gpu, r)
    for dt in (1.0 / 60.0, 1.0 / 30.0, 1.0 / 120.0, 1.0 / 45.0, 1.0 / 60.0):
        parity.lockstep(g, r, 12, dt=dt, tol=0.0)
    parity.lockstep(g, r, 3, dt=0.0, tol=0.0)           # dt = 0: collide only, joints untouched
    scene = getattr(scenes, name)()
    scene.world_flags &= ~T.WORLD_WARM_STARTING
    r = ref.RefWorld(scene)
    g = parity.gpu_world_from_ref(gpu, r)
    parity.lockstep(g, r, 60, tol=0.0, vel_iters=5, pos_iters=2)
    assert (g.get_joints()["impulse"] != 0.0).any()
