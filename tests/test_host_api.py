"""The C++ host API (b2World, b2Body, b2Fixture, b2CudaStepExecutor) against the oracle.

CPU part: worlds built with CreateBody / CreateFixture hold exactly the state the reference's constructors
compute (mass data, sweeps, tight and fat AABBs, flags, shape table).  GPU part: the same worlds stepped through
b2World::Step(dt, vIters, pIters, b2CudaStepExecutor&) stay bit-identical to the oracle, callbacks included, and
the reference's HelloWorld program compiles and runs unchanged apart from the executor type."""
import os
import subprocess

import numpy as np
import pytest

import b2cuda_types as T
import b2host
import parity
import ref
import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "box2d-mt_b200", "host")

BUILD_SCENES = {
    "hello": scenes.hello_world,
    "pyramid": lambda: scenes.pyramid(6),
    "pile": lambda: scenes.pile(10, 8),
    "tumbler": lambda: scenes.tumbler(60),
    "add_pair": lambda: scenes.add_pair(80),
    "stacks": lambda: scenes.pyramids(2, 5, thick_polygon_ground=True),
    "chains": lambda: scenes.chain_terrain(12),
    "sensors": lambda: scenes.sensors(12),
}


@pytest.mark.parametrize("name", sorted(BUILD_SCENES))
def test_host_built_world_equals_reference(name):
    scene = BUILD_SCENES[name]()
    r = ref.RefWorld(scene)
    h = b2host.HostWorld(scene)
    rb, hb = r.bodies(), h.bodies()
    parity.compare_bodies(hb, rb)
    for f in ("invMass", "invI", "lcx", "lcy", "linearDamping", "angularDamping", "gravityScale", "alpha0"):
        parity.assert_floats_equal("body." + f, hb[f], rb[f])
    _, rp = parity.dedupe_shapes(r.shapes(), r.proxies())
    hp = h.proxies()
    parity.compare_proxies(hp, rp)
    for f in ("body", "shape", "flags", "categoryBits", "maskBits", "groupIndex", "fixture"):
        assert (hp[f] == rp[f]).all(), f
    for f in ("friction", "restitution"):
        parity.assert_floats_equal("proxy." + f, hp[f], rp[f])
    assert h.hash() == r.hash()  # GetBodyList() order (newest first) and transforms


def test_host_mass_data_matches_reference():
    # multi-fixture body (summation order matters) and polygons built by hull computation
    scene = scenes.tumbler(10)
    r = ref.RefWorld(scene)
    h = b2host.HostWorld(scene)
    rb = r.bodies()
    m = h.mass()
    inv = np.where(m[:, 0] > 0, 1.0 / m[:, 0].astype(np.float64), 0.0)
    dyn = (rb["flags"] & T.BODY_TYPE_MASK) == T.DYNAMIC_BODY
    assert np.allclose(inv[dyn], rb["invMass"][dyn], rtol=1e-6)


def test_host_mutators_match_reference():
    scene = scenes.pile(6, 5)
    r = ref.RefWorld(scene)
    h = b2host.HostWorld(scene)
    for w in (r, h):
        w.set_transform(5, 0.3, 4.2, 0.7)
        w.set_velocity(6, 1.5, -2.0, 0.25)
        w.apply_force(7, 3.0, 4.0, 0.5)
        w.set_awake(8, False)
    parity.compare_bodies(h.bodies(), r.bodies())
    parity.compare_proxies(h.proxies(), r.proxies())


def test_step_without_gpu_fails_loudly():
    import b2cuda
    if b2cuda.device_count() > 0:
        pytest.skip("device present")
    code = ("import sys; sys.path.insert(0, %r); import b2host, scenes; "
            "w = b2host.HostWorld(scenes.hello_world()); w.step()" % os.path.join(ROOT, "box2d-mt_b200", "python"))
    p = subprocess.run(["python", "-c", code], capture_output=True, text=True)
    assert p.returncode != 0
    assert "no CPU step path" in p.stderr or "b2World::Step failed" in p.stderr


def test_hello_world_program_compiles_against_host_api(tmp_path):
    exe = tmp_path / "hello"
    subprocess.run(["g++", "-std=c++11", "-O2", "-I", HOST, "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "hello_world.cpp"), "-L", os.path.join(ROOT, "box2d-mt_b200"),
                    "-lbox2d_b200", "-lb2cuda", "-Wl,-rpath," + os.path.join(ROOT, "box2d-mt_b200"), "-o", str(exe)],
                   check=True)
    assert exe.exists()


# ---------------------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------------------

def _host_lockstep(scene, steps, check_events=True):
    r = ref.RefWorld(scene)
    h = b2host.HostWorld(scene)
    begins = 0
    for s in range(steps):
        h.step()
        if r.joint_count:
            r.set_joint_order(h.joint_order())
        assert r.step_ordered(h.solver_order()) == 0, s
        try:
            parity.compare_bodies(h.bodies(), r.bodies())
            if r.joint_count:
                # b2RevoluteJoint::GetReactionForce / GetReactionTorque / GetMotorTorque / GetJointAngle / GetJointSpeed
                parity.assert_floats_equal("joint readings", h.joint_readings(), r.joint_readings())
            xya, awake = h.transforms()
            rb = r.bodies()
            assert (xya[:, 0] == rb["px"]).all() and (xya[:, 1] == rb["py"]).all() and (xya[:, 2] == rb["a"]).all()
            assert (awake == ((rb["flags"] & T.BODY_AWAKE) != 0)).all()
            if check_events:
                for kind in (T.EVENT_BEGIN, T.EVENT_END):
                    g, w = h.events(kind), r.events(kind)
                    assert len(g) == len(w) and (g == w).all(), ("events", kind)
                begins += len(h.events(T.EVENT_BEGIN))
            assert h.counts()[2] == r.counts()[2]
        except AssertionError as e:
            raise AssertionError("step %d: %s" % (s, e))
    return h, r, begins


@pytest.mark.gpu
@pytest.mark.parametrize("name,steps", [("chains", 300), ("sensors", 250), ("pyramid", 200), ("pile", 200), ("tumbler", 150), ("stacks", 200)])
def test_host_api_lockstep(gpu, name, steps):
    """TestMT.cpp:91-110 rule (position, angle, awake of every body after every step) between the reference and a
    world stepped through b2World::Step with a b2CudaStepExecutor; deferred Begin/End callbacks in the same order."""
    h, r, begins = _host_lockstep(BUILD_SCENES[name](), steps)
    assert begins > 0
    assert h.hash() == r.hash()
    keys, touching, points = h.contacts()  # b2World::GetContactList snapshot
    rc = r.contacts()
    assert (keys == T.contact_keys(rc)).all()
    assert (touching == ((rc["flags"] & T.CONTACT_TOUCHING) != 0)).all()
    assert (points == rc["manifold"]["pointCount"]).all()
    prof = h.profile()
    assert prof[0] > 0 and prof[2] > 0  # step and solve times come from device events


@pytest.mark.gpu
@pytest.mark.parametrize("name,steps", [("tumbler_joint", 150), ("hanging_chains", 300), ("joint_zoo", 300),
                                        ("rods_and_welds", 300), ("sliders", 300), ("machines", 300), ("pulleys_and_mice", 300), ("gears", 300)])
def test_host_api_joints_lockstep(gpu, name, steps):
    """Worlds with revolute, distance and weld joints built through b2World::CreateJoint (the Testbed's Tumbler with its
    motor joint, hanging chains, every branch of the revolute joint, the Web and the Cantilever) and stepped through
    b2World::Step stay bit-identical to the reference: bodies, events and what the joints' accessors report."""
    make = {"tumbler_joint": lambda: scenes.tumbler(60, motor_joint=True), "hanging_chains": lambda: scenes.hanging_chains(3, 10),
            "joint_zoo": scenes.joint_zoo, "rods_and_welds": scenes.rods_and_welds, "sliders": scenes.sliders, "machines": scenes.machines, "pulleys_and_mice": scenes.pulleys_and_mice, "gears": scenes.gears}[name]
    h, r, begins = _host_lockstep(make(), steps)
    assert h.joint_count() == r.joint_count > 0
    assert h.hash() == r.hash()


@pytest.mark.gpu
def test_host_api_joint_edits_between_steps(gpu):
    """SetMotorSpeed / EnableMotor / SetLimits / EnableLimit between steps (they wake the bodies and reset the limit
    impulse), DestroyJoint, and DestroyBody taking its joints along: the worlds stay in lockstep."""
    scene = scenes.joint_zoo()
    r = ref.RefWorld(scene)
    h = b2host.HostWorld(scene)

    def run(n):
        for s in range(n):
            h.step()
            r.set_joint_order(h.joint_order())
            assert r.step_ordered(h.solver_order()) == 0
            parity.compare_bodies(h.bodies(), r.bodies())
            parity.assert_floats_equal("joint readings", h.joint_readings(), r.joint_readings())

    run(40)
    for w in (h, r):
        w.joint_set_motor(0, True, 2.0, 50.0)       # a free pendulum gets a motor
        w.joint_set_motor(4, False, 1.0, 5.0)       # a motor is switched off
        w.joint_set_limits(6, True, -0.3, 0.4)      # limits appear
        w.joint_set_limits(1, False, -0.5, 0.5)     # limits go away
    run(60)
    for w in (h, r):
        w.joint_set_limits(3, True, -1.0, 1.0)      # limits move while the joint is at one of them
        w.destroy_joint(8)                          # the hub's motor joint: the hub falls with its spokes
    run(80)
    for w in (h, r):
        w.destroy_joint(w.joint_count() - 1 if w is h else w.joint_count - 1)
    run(40)
    assert h.joint_count() == r.joint_count


@pytest.mark.gpu
def test_host_api_tumbler_with_bodies_created_between_steps(gpu):
    """The Testbed's Tumbler as it runs there (Tumbler.h:70-93): the drum on its motor joint, one box created at the
    same spot before every step.  In lockstep with the reference, which creates the same boxes."""
    scene = scenes.tumbler(0, motor_joint=True)
    r = ref.RefWorld(scene)
    h = b2host.HostWorld(scene)
    for s in range(260):
        if s < 200:
            one = scenes.Scene()
            b = one.body(T.DYNAMIC_BODY, (0.0, 10.0))
            one.fixture(b, one.box(0.125, 0.125), density=1.0)
            for w in (h, r):
                w.add(one)
        h.step()
        r.set_joint_order(h.joint_order())
        assert r.step_ordered(h.solver_order()) == 0, s
        try:
            parity.compare_bodies(h.bodies(), r.bodies())
        except AssertionError as e:
            raise AssertionError("step %d: %s" % (s, e))
    assert h.counts()[2] == r.counts()[2]
    parity.assert_floats_equal("joint readings", h.joint_readings(), r.joint_readings())


@pytest.mark.gpu
def test_host_api_slider_edits_between_steps(gpu):
    """b2PrismaticJoint::EnableMotor / SetMotorSpeed / SetMaxMotorForce / EnableLimit / SetLimits between steps."""
    scene = scenes.sliders()
    r = ref.RefWorld(scene)
    h = b2host.HostWorld(scene)

    def run(n):
        for s in range(n):
            h.step()
            r.set_joint_order(h.joint_order())
            assert r.step_ordered(h.solver_order()) == 0
            parity.compare_bodies(h.bodies(), r.bodies())
            parity.assert_floats_equal("joint readings", h.joint_readings(), r.joint_readings())

    run(60)
    for w in (h, r):
        w.joint_set_motor(0, True, -5.0, 2000.0)   # the strong motor reverses
        w.joint_set_motor(1, True, 1.0, 100.0)     # a passive slider gets a motor
        w.joint_set_limits(3, True, -1.0, 1.0)     # an unlimited one gets limits
        w.joint_set_limits(4, False, -0.5, 3.0)    # limits removed
        w.joint_set_limits(2, True, 0.5, 2.5)      # equal limits open up
    run(120)


@pytest.mark.gpu
def test_host_api_mouse_drag(gpu):
    """b2MouseJoint::SetTarget every few steps (a drag), b2MotorJoint::SetLinearOffset, and the mouse joint released
    (DestroyJoint), as an interactive program does."""
    scene = scenes.pulleys_and_mice()
    r = ref.RefWorld(scene)
    h = b2host.HostWorld(scene)
    for s in range(200):
        if s % 5 == 0:
            for w in (h, r):
                w.joint_set_target(3, -29.75 + 0.1 * s, 5.25 + 2.0 * np.sin(0.1 * s))
                w.joint_set_target(5, -20.0 - 0.05 * s, 0.5 + 0.02 * s)   # wakes the sleeping body
        if s == 120:
            for w in (h, r):
                w.destroy_joint(4)
        h.step()
        r.set_joint_order(h.joint_order())
        assert r.step_ordered(h.solver_order()) == 0, s
        try:
            parity.compare_bodies(h.bodies(), r.bodies())
            parity.assert_floats_equal("joint readings", h.joint_readings(), r.joint_readings())
        except AssertionError as e:
            raise AssertionError("step %d: %s" % (s, e))


@pytest.mark.gpu
def test_host_api_jointed_islands_sleep_and_wake(gpu):
    """Islands held together by joints fall asleep as a whole (the joints' position error has to be within tolerance
    for that, b2Island.cpp:363-395) and wake as a whole when a joint is edited."""
    scene = scenes.resting_linkage()
    r = ref.RefWorld(scene)
    h = b2host.HostWorld(scene)

    def run(n):
        for s in range(n):
            h.step()
            r.set_joint_order(h.joint_order())
            assert r.step_ordered(h.solver_order()) == 0
            parity.compare_bodies(h.bodies(), r.bodies())
            parity.assert_floats_equal("joint readings", h.joint_readings(), r.joint_readings())

    run(200)
    awake = (r.bodies()["flags"][1:] & T.BODY_AWAKE) != 0
    assert not awake.any(), "everything should have come to rest and gone to sleep"
    assert int(h.step_info()["awakeBodyCount"]) == 0
    for w in (h, r):
        w.joint_set_motor(0, True, 1.0, 20.0)     # wakes the two bars of the linkage, and only them
    run(3)
    awake = (r.bodies()["flags"][1:] & T.BODY_AWAKE) != 0
    assert awake[:2].all() and not awake[2:].any()
    run(100)


@pytest.mark.gpu
def test_host_api_joints_of_inactive_bodies_rest(gpu):
    """A joint with an inactive body is not simulated (b2World.cpp:1300-1304) and does not link islands; it comes back
    with the body (b2Body::SetActive)."""
    scene = scenes.joint_zoo()
    r = ref.RefWorld(scene)
    h = b2host.HostWorld(scene)

    def run(n):
        for s in range(n):
            h.step()
            r.set_joint_order(h.joint_order())
            assert r.step_ordered(h.solver_order()) == 0
            parity.compare_bodies(h.bodies(), r.bodies())
            parity.assert_floats_equal("joint readings", h.joint_readings(), r.joint_readings())

    run(30)
    for w in (h, r):
        w.set_active(2, False)      # a pendulum with limits: its joint to the ground goes quiet
        w.set_active(9, False)      # the hub: thirteen joints go quiet, the spokes fall free
    run(60)
    before = h.joint_readings()[1].copy()
    run(5)
    assert (h.joint_readings()[1][:4] == before[:4]).all()   # impulses of the resting joint do not change
    for w in (h, r):
        w.set_active(9, True)
        w.set_active(2, True)
    run(90)


@pytest.mark.gpu
def test_host_api_jointed_body_changes_type_and_teleports(gpu):
    """b2Body::SetType on a link in the middle of a chain (the joint colouring depends on which bodies are dynamic: a
    static link may be shared by both of its joints' classes; islands stop at it) and b2Body::SetTransform of a link far
    away from its hinges (the position solver pulls it back under its correction clamps)."""
    scene = scenes.hanging_chains(2, 10)
    r = ref.RefWorld(scene)
    h = b2host.HostWorld(scene)

    def run(n):
        for s in range(n):
            h.step()
            r.set_joint_order(h.joint_order())
            assert r.step_ordered(h.solver_order()) == 0
            parity.compare_bodies(h.bodies(), r.bodies())
            parity.assert_floats_equal("joint readings", h.joint_readings(), r.joint_readings())

    run(40)
    for w in (h, r):
        w.set_type(5, T.STATIC_BODY)          # link 5 of the first chain freezes where it is
    run(60)
    for w in (h, r):
        w.set_type(5, T.DYNAMIC_BODY)
        w.set_transform(14, 0.0, 30.0, 1.0)   # a link of the second chain is moved two metres off
    run(80)


@pytest.mark.gpu
def test_host_api_destroy_body_takes_its_joints(gpu):
    """b2World::DestroyBody destroys the joints attached to the body first (b2World.cpp:594-610): the hub of the joint
    zoo goes, its motor joint and twelve spoke joints go with it, the spokes fall."""
    scene = scenes.joint_zoo()
    h = b2host.HostWorld(scene)
    for _ in range(30):
        h.step()
    n = h.joint_count()
    h.destroy_body(9)      # the hub: bodies 1..8 are the pendulums
    assert h.joint_count() == n - 13
    for _ in range(120):
        h.step()
    b = h.bodies()
    assert np.isfinite(b["px"]).all() and np.isfinite(b["py"]).all()
    spokes = b[9:21]       # rows shift down by one after the hub's row is removed
    assert (spokes["py"] < 8.0).all(), "the spokes (12.5 .. 15.5 up) are no longer held: they lie on the pendulums or the ground"
    assert len(h.device_world().get_joints()) == n - 13


@pytest.mark.gpu
def test_host_api_spring_edits_between_steps(gpu):
    """b2DistanceJoint::SetLength / SetFrequency / SetDampingRatio and b2WeldJoint::SetFrequency / SetDampingRatio
    between steps (they do not wake anything), and rods cut with DestroyJoint."""
    scene = scenes.rods_and_welds()
    r = ref.RefWorld(scene)
    h = b2host.HostWorld(scene)

    def run(n):
        for s in range(n):
            h.step()
            r.set_joint_order(h.joint_order())
            assert r.step_ordered(h.solver_order()) == 0
            parity.compare_bodies(h.bodies(), r.bodies())
            parity.assert_floats_equal("joint readings", h.joint_readings(), r.joint_readings())

    run(50)
    for w in (h, r):
        w.joint_set_spring(0, 9.0, 4.0, 0.5)     # a web rod gets shorter and stiffer
        w.joint_set_spring(5, 12.0, 0.0, 0.0)    # another one turns rigid
        w.joint_set_spring(14, 0.0, 3.0, 0.2)    # a rigid weld gets a soft angle
        w.joint_set_spring(22, 0.0, 0.0, 0.0)    # a soft weld turns rigid
    run(80)
    for w in (h, r):
        w.destroy_joint(1)                       # cut a web rod
        w.destroy_joint(7)                       # and the pendulum chain's top rod (index shifted by one)
    run(80)


@pytest.mark.gpu
def test_host_api_mutation_between_steps(gpu):
    scene = scenes.pile(8, 6)
    r = ref.RefWorld(scene)
    h = b2host.HostWorld(scene)
    for s in range(120):
        if s % 20 == 10:
            for w in (r, h):
                w.set_velocity(3 + s // 20, 2.0, 5.0, 1.0)
                w.apply_force(10, 0.0, 30.0, 0.0)
        if s == 60:
            for w in (r, h):
                w.set_transform(12, 0.1, 9.0, 0.3)
        h.step()
        assert r.step_ordered(h.solver_order()) == 0
        parity.compare_bodies(h.bodies(), r.bodies())


def _contact_sets_equal(h, r):
    hk, ht, hp = h.contacts()
    rc = r.contacts()
    rk = T.contact_keys(rc)
    assert (hk == rk).all() if len(hk) == len(rk) else False, (len(hk), len(rk))
    rt = ((rc["flags"] & T.CONTACT_TOUCHING) != 0).astype(np.int32)
    assert (ht == rt).all()
    assert (hp == rc["manifold"]["pointCount"]).all()


@pytest.mark.gpu
def test_host_api_teleport_and_refilter_timing(gpu):
    """SetTransform and Refilter buffer a proxy move: the reference finds the new pairs at the END of the next step (no
    e_newFixture), and contacts of a refiltered fixture are re-checked by the next Collide (b2Fixture.cpp:187-220,
    b2ContactManager.cpp:186-197).  Contact sets, touching flags and bodies must agree after every step."""
    scene = scenes.pile(8, 6)
    r = ref.RefWorld(scene)
    h = b2host.HostWorld(scene)
    for s in range(150):
        if s == 40:
            # teleport a body from the top of the pile right on top of another one
            target = h.bodies()[20]
            for w in (r, h):
                w.set_transform(45, float(target["px"]) + 0.05, float(target["py"]) + 0.1, 0.2)
        if s == 80:
            # three neighbouring bodies of the bottom row (fixture i belongs to body i - 2: the container has three
            # fixtures) join a negative group: the contacts among them are destroyed, all their others stay
            for w in (r, h):
                for fixture in (10, 11, 12):
                    w.set_filter(fixture, 0x0001, 0xFFFF, -3)
        if s == 110:
            for w in (r, h):
                for fixture in (10, 11, 12):
                    w.set_filter(fixture, 0x0001, 0xFFFF, 0)
        h.step()
        assert r.step_ordered(h.solver_order()) == 0
        try:
            _contact_sets_equal(h, r)
            parity.compare_bodies(h.bodies(), r.bodies())
        except AssertionError as e:
            raise AssertionError("step %d: %s" % (s, e))


@pytest.mark.gpu
def test_host_api_set_active_between_steps(gpu):
    """b2Body::SetActive(false / true) after the world has been stepped (b2Body.cpp:496-544): the body leaves the
    simulation with its contacts, keeps its state, and comes back with new contacts one step later."""
    scene = scenes.pile(8, 6)
    r = ref.RefWorld(scene)
    h = b2host.HostWorld(scene)
    changes = {40: (12, False), 41: (13, False), 75: (12, True), 110: (13, True), 111: (30, False)}
    for s in range(160):
        if s in changes:
            body, on = changes[s]
            for w in (r, h):
                w.set_active(body, on)
            assert h.counts()[2] == len(r.contacts())
        h.step()
        assert r.step_ordered(h.solver_order()) == 0
        try:
            _contact_sets_equal(h, r)
            parity.compare_bodies(h.bodies(), r.bodies())
        except AssertionError as e:
            raise AssertionError("step %d: %s" % (s, e))


@pytest.mark.gpu
def test_host_api_set_type_between_steps(gpu):
    """b2Body::SetType after the world has been stepped: the body's contacts are destroyed at once, its proxies
    touched, mass data reset (b2Body.cpp:118-188).  dynamic -> static -> dynamic -> kinematic."""
    scene = scenes.pile(8, 6)
    r = ref.RefWorld(scene)
    h = b2host.HostWorld(scene)
    changes = {50: (14, T.STATIC_BODY), 90: (14, T.DYNAMIC_BODY), 120: (30, T.KINEMATIC_BODY), 121: (31, T.STATIC_BODY)}
    for s in range(170):
        if s in changes:
            body, kind = changes[s]
            for w in (r, h):
                w.set_type(body, kind)
            assert h.counts()[2] == len(r.contacts()), "contacts of the body must be gone right away"
        h.step()
        assert r.step_ordered(h.solver_order()) == 0
        try:
            _contact_sets_equal(h, r)
            parity.compare_bodies(h.bodies(), r.bodies())
        except AssertionError as e:
            raise AssertionError("step %d: %s" % (s, e))


@pytest.mark.gpu
def test_host_api_custom_contact_filter(gpu):
    """b2World::SetContactFilter with a user subclass: evaluated on the host for the new pairs of every step."""
    scene = scenes.pile(8, 6)
    r = ref.RefWorld(scene)
    h = b2host.HostWorld(scene)
    for w in (r, h):
        w.set_modulo_filter(4)
    for s in range(150):
        h.step()
        assert r.step_ordered(h.solver_order()) == 0
        try:
            _contact_sets_equal(h, r)
            parity.compare_bodies(h.bodies(), r.bodies())
        except AssertionError as e:
            raise AssertionError("step %d: %s" % (s, e))
    hk, _, _ = h.contacts()
    assert len(hk) > 0 and (((hk >> np.uint64(32)) + (hk & np.uint64(0xFFFFFFFF))) % np.uint64(4) != 0).all()


@pytest.mark.gpu
def test_host_api_post_solve_reports(gpu):
    """b2CudaStepOptions::reportPostSolve: one PostSolve per solved contact and step, with the impulses the reference
    reports (b2Island::Report): equal call counts and an equal digest over (key, count, impulses) of all calls."""
    scene = scenes.pile(8, 6)
    r = ref.RefWorld(scene)
    h = b2host.HostWorld(scene)
    r.record_post_solve()
    h.record_post_solve()
    for s in range(120):
        h.step()
        assert r.step_ordered(h.solver_order()) == 0
        assert h.post_solve_digest() == r.post_solve_digest(), "step %d" % s
    assert h.post_solve_digest()[1] > 1000


@pytest.mark.gpu
def test_host_api_pre_solve_can_disable_contacts(gpu):
    """b2CudaStepOptions::reportPreSolve: PreSolve runs between the narrow phase and the solver of the device step
    with the previous manifold, and SetEnabled(false) there keeps the contact out of the step's solve.  Both sides
    disable every contact whose key is a multiple of 3: same calls (digest), same bodies, after every step."""
    scene = scenes.pile(8, 6)
    r = ref.RefWorld(scene)
    h = b2host.HostWorld(scene)
    r.set_pre_solve_rule(3)
    h.set_pre_solve_rule(3)
    for s in range(150):
        h.step()
        assert r.step_ordered(h.solver_order()) == 0
        assert h.pre_solve_digest() == r.pre_solve_digest(), "step %d" % s
        try:
            parity.compare_bodies(h.bodies(), r.bodies())
        except AssertionError as e:
            raise AssertionError("step %d: %s" % (s, e))
    assert h.pre_solve_digest()[1] > 1000
    # bodies sink into each other where their contact is switched off: the rule really took effect
    keys, touching, _ = h.contacts()
    assert ((keys % np.uint64(3) == 0) & (touching != 0)).any()


def _check_queries(h, r, rng, count=40):
    for _ in range(count):
        c = rng.uniform(-6.0, 6.0, 2)
        e = rng.uniform(0.1, 2.5, 2)
        box = [c[0] - e[0], c[1] - e[1] + 2.0, c[0] + e[0], c[1] + e[1] + 2.0]
        assert (h.query_aabb(box) == r.query_aabb(box)).all()
        p1 = rng.uniform(-8.0, 8.0, 2) + [0.0, 6.0]
        p2 = rng.uniform(-8.0, 8.0, 2) + [0.0, -1.0]
        hp, ho = h.ray_cast_closest(p1, p2)
        rp, ro = r.ray_cast_closest(p1, p2)
        # the closest hit does not depend on the order in which the candidates are visited
        assert hp == rp, (p1, p2)
        if hp >= 0:
            assert (ho.view(np.uint32) == ro.view(np.uint32)).all()


def test_world_queries_before_the_first_step():
    scene = scenes.chain_terrain(12)
    r = ref.RefWorld(scene)
    h = b2host.HostWorld(scene)
    _check_queries(h, r, np.random.default_rng(2))


@pytest.mark.gpu
def test_world_queries_between_steps(gpu):
    """b2World::QueryAABB / RayCast answered from the device's fat boxes, after steps and after host edits that have
    not been stepped yet (SetTransform), and ShiftOrigin."""
    scene = scenes.chain_terrain(24)
    r = ref.RefWorld(scene)
    h = b2host.HostWorld(scene)
    rng = np.random.default_rng(4)
    for s in range(90):
        h.step()
        assert r.step_ordered(h.solver_order()) == 0
        if s % 30 == 29:
            _check_queries(h, r, rng, 25)
    for w in (r, h):
        w.set_transform(5, 1.0, 6.0, 0.4)
    _check_queries(h, r, rng, 25)
    h.shift_origin(3.0, -1.0)
    moved = h.bodies()
    assert np.allclose(moved["px"] + 3.0, r.bodies()["px"], atol=1e-5) and np.allclose(moved["py"] - 1.0, r.bodies()["py"], atol=1e-5)
    for _ in range(10):
        h.step()
    assert np.isfinite(h.bodies()["py"]).all()


@pytest.mark.gpu
def test_host_api_lazy_download(gpu):
    """downloadBodies=false: the mirror is refreshed on first access only; results are the same."""
    a = b2host.HostWorld(scenes.pile(8, 6), download_bodies=True)
    b = b2host.HostWorld(scenes.pile(8, 6), download_bodies=False, events=False)
    for _ in range(60):
        a.step()
        b.step()
    assert a.bodies().tobytes() == b.bodies().tobytes()


@pytest.mark.gpu
def test_host_api_setters_without_body_download(gpu):
    """downloadBodies=false leaves the host rows stale after a step; every setter must edit the CURRENT row (fetched on
    demand), never write a stale one back: damping, gravity scale, bullet flag, sleeping allowed, forces.  In lockstep
    with the reference, which gets the same calls."""
    scene = scenes.pile(8, 6, sleep=True)
    r = ref.RefWorld(scene)
    h = b2host.HostWorld(scene, download_bodies=False, events=False)
    H = b2host.HostWorld
    for s in range(240):
        for w in (h, r):
            if s == 20:
                w.set_body_param(5, H.LINEAR_DAMPING, 0.4)
                w.set_body_param(6, H.ANGULAR_DAMPING, 0.8)
            if s == 35:
                w.set_body_param(7, H.GRAVITY_SCALE, -0.5)
                w.set_body_param(8, H.BULLET, 1)
            if s == 50:
                w.set_body_param(9, H.SLEEPING_ALLOWED, 0)
                w.set_body_param(10, H.SLEEPING_ALLOWED, 1)
            if s == 70:
                w.set_body_param(7, H.GRAVITY_SCALE, 1.0)
                w.apply_force(11, 3.0, 40.0, 0.5)
        h.step()
        assert r.step_ordered(h.solver_order()) == 0, s
        if s % 10 == 0 or s > 230:
            try:
                parity.compare_bodies(h.bodies(), r.bodies())
            except AssertionError as e:
                raise AssertionError("step %d: %s" % (s, e))
    a = h.bodies()
    # body 9 may not sleep, the rest of the pile does; body 8 is a bullet (its contacts became time-of-impact candidates)
    assert a["flags"][9] & T.BODY_AWAKE and not (a["flags"][9] & T.BODY_AUTOSLEEP)
    assert a["flags"][8] & T.BODY_BULLET
    rc = r.contacts()
    assert ((rc["flags"] & T.CONTACT_TOI_CANDIDATE) != 0).sum() > 0


@pytest.mark.gpu
def test_host_api_full_upload_keeps_joint_impulses(gpu):
    """DestroyFixture on a body that has nothing to do with the joints makes the host re-upload everything, the joint
    table included.  The table must carry the joints' CURRENT accumulated impulses and limit states (warm starting),
    not the ones the objects held before the last steps: the world stays in lockstep with the reference, where the
    same fixture is destroyed."""
    scene = scenes.hanging_chains(3, 10)
    ground2 = scene.body(T.STATIC_BODY, (60.0, -1.0))
    scene.fixture(ground2, scene.box(5.0, 1.0), density=0.0)
    lone = scene.body(T.DYNAMIC_BODY, (60.0, 1.0))
    scene.fixture(lone, scene.box(0.5, 0.5), density=1.0)
    scene.fixture(lone, scene.circle(0.3, (0.0, 0.6)), density=1.0)   # the last fixture of the world
    r = ref.RefWorld(scene)
    h = b2host.HostWorld(scene)
    n_fixtures = h.counts()[1]
    for s in range(200):
        if s == 60:
            h.destroy_fixture(n_fixtures - 1)
            r.destroy_last_fixture()
        h.step()
        r.set_joint_order(h.joint_order())
        assert r.step_ordered(h.solver_order()) == 0, s
        try:
            parity.compare_bodies(h.bodies(), r.bodies())
            parity.assert_floats_equal("joint readings", h.joint_readings(), r.joint_readings())
        except AssertionError as e:
            raise AssertionError("step %d: %s" % (s, e))
    assert h.counts()[1] == n_fixtures - 1


@pytest.mark.gpu
def test_host_api_destroy_body(gpu):
    scene = scenes.pile(8, 6)
    h = b2host.HostWorld(scene)
    for _ in range(90):
        h.step()
    n0 = h.counts()
    ends_before = len(h.events(T.EVENT_END))
    h.destroy_body(10)
    h.destroy_body(20)
    assert h.counts()[0] == n0[0] - 2
    for _ in range(60):
        h.step()
    b = h.bodies()
    assert np.isfinite(b["py"]).all() and (b["py"][1:] > -0.5).all()
    assert ends_before >= 0


@pytest.mark.gpu
def test_hello_world_program_runs(gpu, tmp_path):
    """The reference's HelloWorld, executor type swapped, prints the reference's 60 lines."""
    exe = tmp_path / "hello"
    subprocess.run(["g++", "-std=c++11", "-O2", "-I", HOST, "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "hello_world.cpp"), "-L", os.path.join(ROOT, "box2d-mt_b200"),
                    "-lbox2d_b200", "-lb2cuda", "-Wl,-rpath," + os.path.join(ROOT, "box2d-mt_b200"), "-o", str(exe)],
                   check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.strip().splitlines()
    with open(os.path.join(ROOT, "tests", "golden", "helloworld.txt")) as f:
        want = [ln.strip() for ln in f if ln.strip()]
    got = [ln.strip() for ln in out]
    # all 60 lines, the landing through b2World::SolveTOI included
    assert got == want


def _compile_cpp(tmp_path, name):
    exe = tmp_path / name
    subprocess.run(["g++", "-std=c++11", "-O2", "-I", HOST, "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", name + ".cpp"), "-L", os.path.join(ROOT, "box2d-mt_b200"),
                    "-lbox2d_b200", "-lb2cuda", "-Wl,-rpath," + os.path.join(ROOT, "box2d-mt_b200"), "-o", str(exe)],
                   check=True)
    return exe


def test_world_dump_rebuilds_the_same_world(tmp_path):
    """b2World::Dump writes C++ that rebuilds the world (b2World.cpp:2107-2164): a world with every shape class and
    eight joints (two gears among them) is dumped, rebuilt from the dump, and dumped again: same text, same counts."""
    args = ["g++", "-std=c++11", "-O1", "-I", HOST, "-I", os.path.join(ROOT, "include"), "-I", str(tmp_path),
            os.path.join(ROOT, "tests", "cpp", "dump_world.cpp"), "-L", os.path.join(ROOT, "box2d-mt_b200"),
            "-lbox2d_b200", "-lb2cuda", "-Wl,-rpath," + os.path.join(ROOT, "box2d-mt_b200")]
    first = tmp_path / "dump_first"
    subprocess.run(args + ["-o", str(first)], check=True)
    a = subprocess.run([str(first)], capture_output=True, text=True, check=True)
    assert a.stderr.strip() == "4 bodies 8 joints 8 proxies"
    assert a.stdout.count("CreateJoint") == 8 and a.stdout.count("CreateFixture") == 6
    assert "b2GearJointDef" in a.stdout and "b2ChainShape" in a.stdout and "fd.thickShape = true" in a.stdout
    (tmp_path / "dump.inc").write_text(a.stdout)
    second = tmp_path / "dump_second"
    subprocess.run(args + ["-DREBUILD", "-o", str(second)], check=True)
    b = subprocess.run([str(second)], capture_output=True, text=True, check=True)
    assert b.stderr == a.stderr
    assert b.stdout == a.stdout


def test_tumbler_program_compiles_against_host_api(tmp_path):
    assert _compile_cpp(tmp_path, "tumbler").exists()


@pytest.mark.gpu
def test_tumbler_program_runs(gpu, tmp_path):
    """The Testbed's Tumbler as user code: the motor joint turns the drum at its set speed, every box stays inside."""
    exe = _compile_cpp(tmp_path, "tumbler")
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    angle, speed, torque, x0, y0, x1, y1 = (float(v) for v in out)
    assert abs(speed - 0.05 * np.pi) < 1e-4
    assert abs(angle - 0.05 * np.pi * 400 / 60) < 1e-2
    assert torque != 0.0
    assert -9.5 < x0 and x1 < 9.5 and -9.5 < y0 and y1 < 9.5     # drum frame: inside the four walls
