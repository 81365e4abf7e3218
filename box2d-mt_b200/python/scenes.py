"""Headless restatements of the scenes named by BASELINE.json / SURVEY.md 8d, as flat Scene descriptions.

The Testbed scene classes of the reference derive from a GL `Test` base and cannot be built headless, so the
recipes are restated here (citing the scene file they follow).  All position arithmetic that the reference
does in float32 is done in np.float32 here so that body placement is bit-identical.
"""
import numpy as np

import b2cuda_types as T
from b2scene import (Scene, BODY_DEF, FIXTURE_DEF, BODYDEF_DEFAULT, BODYDEF_BULLET, BODYDEF_ALLOW_SLEEP,
                     BODYDEF_FIXED_ROTATION, BODYDEF_AWAKE)

F = np.float32


def hello_world():
    """HelloWorld/HelloWorld.cpp:29-76: ground box 50x10 at (0,-10), dynamic 1x1 box at (0,4), friction 0.3."""
    s = Scene()
    g = s.body(T.STATIC_BODY, (0.0, -10.0))
    s.fixture(g, s.box(50.0, 10.0), density=0.0)
    b = s.body(T.DYNAMIC_BODY, (0.0, 4.0))
    s.fixture(b, s.box(1.0, 1.0), density=1.0, friction=0.3)
    return s


def _pyramid_rows(s, shape, x0, y0, count, density=5.0):
    """Testbed/Tests/Pyramid.h:49-68: row i has count-i boxes, offsets (0.5625, 1.25) / (1.125, 0)."""
    x = np.array([x0, y0], dtype=F)
    dx = np.array([0.5625, 1.25], dtype=F)
    dy = np.array([1.125, 0.0], dtype=F)
    for i in range(count):
        y = x.copy()
        for _ in range(i, count):
            b = s.body(T.DYNAMIC_BODY, (y[0], y[1]))
            s.fixture(b, shape, density=density)
            y = (y + dy).astype(F)
        x = (x + dx).astype(F)


def pyramid(count=20, continuous=True):
    """Testbed/Tests/Pyramid.h:30-69 exactly: edge ground (-40,0)-(40,0), boxes 0.5 half-extent, density 5."""
    flags = T.WORLD_DEFAULT if continuous else (T.WORLD_DEFAULT & ~T.WORLD_CONTINUOUS)
    s = Scene(world_flags=flags)
    g = s.body(T.STATIC_BODY, (0.0, 0.0))
    s.fixture(g, s.edge((-40.0, 0.0), (40.0, 0.0)), density=0.0)
    _pyramid_rows(s, s.box(0.5, 0.5), -7.0, 0.75, count)
    return s


def pyramids(pyramid_count=40, size=20, thick_polygon_ground=False, continuous=True):
    """Testbed/Tests/SleepCollidePerf.h:35-83 (pyramids only; SURVEY.md Appendix B layout).

    thick_polygon_ground replaces the edge ground by a static thick-shape box (config C4 of SURVEY.md 8d),
    which removes the ground contacts from the TOI candidate set."""
    flags = T.WORLD_DEFAULT if continuous else (T.WORLD_DEFAULT & ~T.WORLD_CONTINUOUS)
    s = Scene(world_flags=flags)
    g = s.body(T.STATIC_BODY, (0.0, 0.0))
    half = float(F(20.0) * F(pyramid_count))
    if thick_polygon_ground:
        s.fixture(g, s.box(half + 20.0, 1.0, center=(0.0, -1.0), angle=0.0), density=0.0, thick=True)
    else:
        s.fixture(g, s.edge((-half, 0.0), (half, 0.0)), density=0.0)
    spacing = F(1.125) * F(size)
    x_init = F(F(F(-spacing * F(pyramid_count)) * F(0.5)) - F(7.0))
    box = s.box(0.5, 0.5)
    for _ in range(pyramid_count):
        _pyramid_rows(s, box, x_init, 0.75, size)
        x_init = F(x_init + spacing)
    return s


class _Rand:
    """rand() & 32767 based RandomFloat of Testbed/Framework/Test.h:45-60, with a portable LCG behind it."""

    def __init__(self, seed=0):
        self.rng = np.random.RandomState(seed)

    def uniform(self, lo, hi, n=None):
        r = self.rng.randint(0, 32768, size=n).astype(F) / F(32767.0)
        return (F(hi - lo) * r + F(lo)).astype(F)


def add_pair(count=400, seed=0):
    """Testbed/Tests/AddPair.h:26-60 scaled to `count` circles at constant density:
    circles r=0.1 density 0.01 in [-6s,0]x[5-s,5+s], s = sqrt(count/400); zero gravity; bullet box 1.5
    half-extent at (-40,5) with v=(150,0)."""
    s = Scene(gravity=(0.0, 0.0))
    scale = float(np.sqrt(count / 400.0))
    rnd = _Rand(seed)
    circle = s.circle(0.1)
    xs = rnd.uniform(-6.0 * scale, 0.0, count)
    ys = rnd.uniform(5.0 - scale, 5.0 + scale, count)
    for i in range(count):
        b = s.body(T.DYNAMIC_BODY, (xs[i], ys[i]))
        s.fixture(b, circle, density=0.01)
    b = s.body(T.DYNAMIC_BODY, (-40.0, 5.0), vel=(150.0, 0.0), flags=BODYDEF_DEFAULT | BODYDEF_BULLET)
    s.fixture(b, s.box(1.5, 1.5), density=1.0)
    return s


def tumbler(count=800, seed=0, scale=None, motor_joint=False):
    """Testbed/Tests/Tumbler.h:47-54 walls on a KINEMATIC body at (0,10) spinning at 0.05*pi rad/s
    (the revolute-joint motor is replaced by a kinematic container), scaled so that `count`
    boxes of half-extent 0.125 pre-placed on a jittered grid fit inside."""
    s = Scene(world_flags=T.WORLD_DEFAULT & ~T.WORLD_ALLOW_SLEEP)
    if scale is None:
        scale = max(1.0, float(np.sqrt(count / 800.0)))
    k = scale
    if motor_joint:
        # the Testbed's own arrangement (Tumbler.h:33-68): a dynamic container on a revolute joint with a motor
        g = s.body(T.STATIC_BODY, (0.0, 0.0))
        c = s.body(T.DYNAMIC_BODY, (0.0, 10.0 * k), flags=BODYDEF_DEFAULT & ~BODYDEF_ALLOW_SLEEP)
        s.revolute_joint(g, c, (0.0, 10.0 * k), (0.0, 0.0), motor=(0.05 * np.pi, 1e8))
    else:
        c = s.body(T.KINEMATIC_BODY, (0.0, 10.0 * k), w=0.05 * np.pi, flags=BODYDEF_DEFAULT & ~BODYDEF_ALLOW_SLEEP)
    s.fixture(c, s.box(0.5 * k, 10.0 * k, center=(10.0 * k, 0.0), angle=0.0), density=5.0, thick=True)
    s.fixture(c, s.box(0.5 * k, 10.0 * k, center=(-10.0 * k, 0.0), angle=0.0), density=5.0, thick=True)
    s.fixture(c, s.box(10.0 * k, 0.5 * k, center=(0.0, 10.0 * k), angle=0.0), density=5.0, thick=True)
    s.fixture(c, s.box(10.0 * k, 0.5 * k, center=(0.0, -10.0 * k), angle=0.0), density=5.0, thick=True)
    box = s.box(0.125, 0.125)
    rnd = _Rand(seed)
    n = int(np.ceil(np.sqrt(count)))
    pitch = 0.3
    jx = rnd.uniform(-0.02, 0.02, count)
    jy = rnd.uniform(-0.02, 0.02, count)
    for i in range(count):
        gx, gy = i % n, i // n
        x = (gx - 0.5 * (n - 1)) * pitch + jx[i]
        y = 10.0 * k + (gy - 0.5 * (n - 1)) * pitch + jy[i]
        b = s.body(T.DYNAMIC_BODY, (x, y))
        s.fixture(b, box, density=1.0)
    return s


def _regular_polygon(sides, radius):
    """Testbed/Tests/ManyBodies.h:277-291: regular polygon, vertex k at angle 2*pi*k/sides."""
    ang = (2.0 * np.pi * np.arange(sides) / sides)
    return [(float(F(radius * np.cos(a))), float(F(radius * np.sin(a)))) for a in ang]


def pile(columns, rows, seed=0, pitch=0.56, radius=0.25, sleep=False, wall_height=None):
    """mixed circle / regular 3-8-gon pile (circumradius 0.25, density 1) dropped from a jittered
    grid into a static thick-shape container.  columns*rows bodies; 50% circles."""
    flags = T.WORLD_DEFAULT if sleep else (T.WORLD_DEFAULT & ~T.WORLD_ALLOW_SLEEP)
    s = Scene(world_flags=flags)
    width = columns * pitch
    height = rows * pitch if wall_height is None else wall_height
    g = s.body(T.STATIC_BODY, (0.0, 0.0))
    s.fixture(g, s.box(0.5 * width + 2.0, 1.0, center=(0.0, -1.0), angle=0.0), thick=True)
    s.fixture(g, s.box(1.0, 0.5 * height + 2.0, center=(-0.5 * width - 1.0, 0.5 * height), angle=0.0), thick=True)
    s.fixture(g, s.box(1.0, 0.5 * height + 2.0, center=(0.5 * width + 1.0, 0.5 * height), angle=0.0), thick=True)
    shapes = np.array([s.circle(radius)] + [s.polygon(_regular_polygon(k, radius)) for k in range(3, 9)])
    rnd = _Rand(seed)
    n = columns * rows
    kind = rnd.rng.randint(0, 12, size=n)
    jx = rnd.uniform(-0.02, 0.02, n)
    jy = rnd.uniform(-0.02, 0.02, n)
    ang = rnd.uniform(0.0, 2.0 * np.pi, n)
    idx = np.arange(n)
    cx, cy = idx % columns, idx // columns
    bodies = np.zeros(n, BODY_DEF)
    bodies["type"] = T.DYNAMIC_BODY
    bodies["px"] = ((cx + 0.5) * pitch - 0.5 * width + jx).astype(F)
    bodies["py"] = ((cy + 0.5) * pitch + 0.05 + jy).astype(F)
    bodies["angle"] = ang
    bodies["gravityScale"] = 1.0
    bodies["flags"] = BODYDEF_DEFAULT
    fixtures = np.zeros(n, FIXTURE_DEF)
    fixtures["body"] = idx
    fixtures["shape"] = np.where(kind >= 6, shapes[0], shapes[1 + np.minimum(kind, 5)])
    fixtures["density"] = 1.0
    fixtures["friction"] = 0.2
    fixtures["categoryBits"] = 0x0001
    fixtures["maskBits"] = 0xFFFF
    s.add_bulk(bodies, fixtures)
    return s


def chain_terrain(count=40, seed=3):
    """Mixed bodies dropped on chain shapes: an open 8-vertex chain as a bumpy floor (continued on both sides by
    plain edges with ghost vertices) and a closed loop as an obstacle -- every chain segment is its own proxy."""
    s = Scene()
    g = s.body(T.STATIC_BODY, (0.0, 0.0))
    floor = [(-14.0, 3.0), (-10.0, 0.5), (-6.0, 1.0), (-2.0, 0.0), (2.0, 0.4), (6.0, 0.0), (10.0, 1.2), (14.0, 3.5)]
    s.fixture(g, s.chain(floor), friction=0.5)
    s.fixture(g, s.chain([(-1.5, 2.0), (0.0, 1.4), (1.5, 2.0), (0.0, 2.8)], loop=True), friction=0.3)
    s.fixture(g, s.edge((-18.0, 6.0), (-14.0, 3.0), v3=(-10.0, 0.5)))
    rnd = _Rand(seed)
    shapes = [s.circle(0.3), s.box(0.35, 0.25)] + [s.polygon(_regular_polygon(k, 0.33)) for k in (3, 5, 8)]
    for i in range(count):
        b = s.body(T.DYNAMIC_BODY, (-11.0 + 0.55 * i, 4.5 + 0.9 * (i % 4)), angle=float(rnd.uniform(0.0, 6.28, 1)[0]))
        s.fixture(b, shapes[i % len(shapes)], density=1.0, friction=0.4, restitution=0.1 * (i % 3))
    return s


def sensors(count=30, seed=5):
    """Sensor fixtures (b2Fixture::IsSensor): a static detection zone (box and circle sensors), dynamic bodies that
    carry a sensor halo besides their solid fixture, all shape classes falling through the zones onto a floor."""
    s = Scene()
    g = s.body(T.STATIC_BODY, (0.0, 0.0))
    s.fixture(g, s.box(12.0, 0.5, center=(0.0, -0.5)), thick=True)
    s.fixture(g, s.box(3.0, 0.6, center=(-3.0, 3.0), angle=0.2), sensor=True)
    s.fixture(g, s.circle(1.2, p=(4.0, 2.5)), sensor=True)
    s.fixture(g, s.edge((-8.0, 4.5), (-5.0, 4.0)), sensor=True)
    rnd = _Rand(seed)
    shapes = [s.circle(0.25), s.box(0.3, 0.2)] + [s.polygon(_regular_polygon(k, 0.28)) for k in (3, 6)]
    halo = s.circle(0.6)
    for i in range(count):
        b = s.body(T.DYNAMIC_BODY, (-8.0 + 0.55 * i, 2.2 + 0.8 * (i % 3)), angle=float(rnd.uniform(0.0, 6.28, 1)[0]))
        s.fixture(b, shapes[i % len(shapes)], density=1.0, friction=0.3)
        if i % 3 == 0:
            s.fixture(b, halo, sensor=True)
    return s



def hanging_chains(chains=6, links=20, seed=0):
    """Testbed/Tests/Chain.h:26-60, several times over: chains of box links (0.6 x 0.125 half-extents, density 20) hinged
    end to end by revolute joints that do not let neighbours collide, the first link hinged to the static ground; the
    chains hang side by side over a thick floor and swing into each other.  A few hinges carry angle limits."""
    s = Scene()
    g = s.body(T.STATIC_BODY, (0.0, 0.0))
    s.fixture(g, s.box(40.0, 1.0, center=(0.0, -1.0), angle=0.0), thick=True)
    link = s.box(0.6, 0.125)
    y = 25.0
    for c in range(chains):
        x0 = -12.0 + 5.0 * c
        prev = g
        for i in range(links):
            b = s.body(T.DYNAMIC_BODY, (x0 + 0.5 + i, y))
            s.fixture(b, link, density=20.0, friction=0.2)
            anchor = (float(F(x0 + i)), y)
            la = anchor if prev == g else (0.5, 0.0)
            limits = (-0.25 * np.pi, 0.5 * np.pi) if (i % 5 == 2 and c % 2 == 0) else None
            s.revolute_joint(prev, b, la, (-0.5, 0.0), limits=limits)
            prev = b
    return s


def joint_zoo(seed=0):
    """Every branch of the revolute joint: free hinges, angle limits (lower / upper / equal), motors weak and strong, a
    hub with more joints than there are parallel colour classes, a joint between two dynamic bodies that may collide,
    a body with fixed rotation, a joint to a kinematic body, sleeping allowed."""
    s = Scene()
    g = s.body(T.STATIC_BODY, (0.0, 0.0))
    s.fixture(g, s.edge((-40.0, 0.0), (40.0, 0.0)))
    bar = s.box(1.0, 0.1)
    small = s.box(0.25, 0.25)
    # pendulums on the ground body
    for i in range(8):
        x = -20.0 + 4.0 * i
        b = s.body(T.DYNAMIC_BODY, (x + 1.0, 6.0), w=float(i) - 3.0)
        s.fixture(b, bar, density=1.0)
        limits = [None, (-0.5, 0.5), (0.0, 0.0), (-2.0, 0.2), None, (-0.1, 1.5), None, (0.3, 0.3)][i]
        motor = [None, None, None, None, (1.0, 5.0), (-2.0, 1000.0), (0.5, 0.0), None][i]
        s.revolute_joint(g, b, (x, 6.0), (-1.0, 0.0), limits=limits, motor=motor)
    # a hub with 12 spokes: more joints on one dynamic body than parallel classes
    hub = s.body(T.DYNAMIC_BODY, (0.0, 14.0))
    s.fixture(hub, s.circle(0.5), density=2.0)
    s.revolute_joint(g, hub, (0.0, 14.0), (0.0, 0.0), motor=(1.0, 200.0))
    for k in range(12):
        a = 2.0 * np.pi * k / 12
        ca, sa = float(F(np.cos(a))), float(F(np.sin(a)))
        b = s.body(T.DYNAMIC_BODY, (float(F(1.5 * ca)), float(F(14.0 + 1.5 * sa))), angle=float(F(a)))
        s.fixture(b, bar, density=0.5)
        s.revolute_joint(hub, b, (float(F(0.5 * ca)), float(F(0.5 * sa))), (-1.0, 0.0), collide_connected=(k % 3 == 0))
    # two free bodies hinged together, falling on the ground; one cannot rotate
    a = s.body(T.DYNAMIC_BODY, (10.0, 3.0))
    s.fixture(a, small, density=1.0)
    b = s.body(T.DYNAMIC_BODY, (10.6, 3.0), flags=BODYDEF_DEFAULT | BODYDEF_FIXED_ROTATION)
    s.fixture(b, small, density=1.0)
    s.revolute_joint(a, b, (0.3, 0.0), (-0.3, 0.0), collide_connected=True)
    # a crank on a kinematic body
    k = s.body(T.KINEMATIC_BODY, (-10.0, 12.0), w=1.0)
    s.fixture(k, small)
    b = s.body(T.DYNAMIC_BODY, (-8.5, 12.0))
    s.fixture(b, bar, density=1.0)
    s.revolute_joint(k, b, (0.5, 0.0), (-1.0, 0.0))
    return s


def rods_and_welds(seed=0):
    """Distance and weld joints in every variant: the Testbed's Web (Web.h:27-110: four boxes held by eight soft rods), a
    rigid-rod pendulum chain, a cantilever of boxes welded rigidly (Cantilever.h:41-70) and one welded with a soft angle,
    a weld between bodies that cannot rotate, rods whose anchors coincide.  Everything falls on / hangs over a ground."""
    s = Scene()
    g = s.body(T.STATIC_BODY, (0.0, 0.0))
    s.fixture(g, s.edge((-40.0, 0.0), (40.0, 0.0)))
    small = s.box(0.5, 0.5)
    # web
    corners = [(-5.0, 5.0), (5.0, 5.0), (5.0, 15.0), (-5.0, 15.0)]
    web = []
    for c in corners:
        b = s.body(T.DYNAMIC_BODY, c)
        s.fixture(b, small, density=5.0)
        web.append(b)
    outer = [(-10.0, 0.0), (10.0, 0.0), (10.0, 20.0), (-10.0, 20.0)]
    inner = [(-0.5, -0.5), (0.5, -0.5), (0.5, 0.5), (-0.5, 0.5)]
    for k in range(4):
        p1 = np.array(outer[k], np.float32)
        p2 = np.array(corners[k], np.float32) + np.array(inner[k], np.float32)
        s.distance_joint(g, web[k], outer[k], inner[k], float(F(np.hypot(*(p2 - p1)))), frequency_hz=2.0, damping_ratio=0.0)
    sides = [((0.5, 0.0), (-0.5, 0.0)), ((0.0, 0.5), (0.0, -0.5)), ((-0.5, 0.0), (0.5, 0.0)), ((0.0, -0.5), (0.0, 0.5))]
    for k in range(4):
        a, b = web[k], web[(k + 1) % 4]
        la, lb = sides[k]
        p1 = np.array(corners[k], np.float32) + np.array(la, np.float32)
        p2 = np.array(corners[(k + 1) % 4], np.float32) + np.array(lb, np.float32)
        s.distance_joint(a, b, la, lb, float(F(np.hypot(*(p2 - p1)))), frequency_hz=2.0, damping_ratio=0.3)
    # rigid-rod pendulum chain hanging from the ground body
    prev, anchor = g, (20.0, 18.0)
    for i in range(6):
        b = s.body(T.DYNAMIC_BODY, (20.0 + 1.5 * (i + 1), 18.0))
        s.fixture(b, s.circle(0.3), density=2.0)
        s.distance_joint(prev, b, anchor, (0.0, 0.0), 1.5, collide_connected=(i % 2 == 0))
        prev, anchor = b, (0.0, 0.0)
    # cantilevers: rigid welds, and welds with a soft angle
    plank = s.box(0.5, 0.125)
    for row, (hz, ratio) in enumerate([(0.0, 0.0), (5.0, 0.7)]):
        prev = g
        y = 8.0 + 4.0 * row
        for i in range(6):
            b = s.body(T.DYNAMIC_BODY, (-30.0 + 0.5 + i, y))
            s.fixture(b, plank, density=20.0)
            la = (-30.0 + i, y) if prev == g else (0.5, 0.0)
            s.weld_joint(prev, b, la, (-0.5, 0.0), frequency_hz=hz, damping_ratio=ratio)
            prev = b
    # two bodies without rotation welded together; two bodies whose rod has zero length
    a = s.body(T.DYNAMIC_BODY, (-18.0, 4.0), flags=BODYDEF_DEFAULT | BODYDEF_FIXED_ROTATION)
    s.fixture(a, small, density=1.0)
    b = s.body(T.DYNAMIC_BODY, (-16.8, 4.0), flags=BODYDEF_DEFAULT | BODYDEF_FIXED_ROTATION)
    s.fixture(b, small, density=1.0)
    s.weld_joint(a, b, (0.6, 0.0), (-0.6, 0.0))
    a = s.body(T.DYNAMIC_BODY, (-14.0, 6.0))
    s.fixture(a, small, density=1.0)
    b = s.body(T.DYNAMIC_BODY, (-14.0, 6.0), vel=(1.0, 0.0))
    s.fixture(b, s.circle(0.4), density=1.0)
    s.distance_joint(a, b, (0.0, 0.0), (0.0, 0.0), 0.0)
    return s


def sliders(seed=0):
    """Prismatic joints in every branch (Testbed/Tests/Prismatic.h:26-70 and more): sliders on the ground body along
    slanted, non-unit axes with limits (lower / upper / equal / none) and motors (strong, weak, none), a piston between
    two dynamic bodies (SliderCrank.h:27-110: crank and rod on revolute joints, piston on a prismatic one), bodies that
    cannot rotate, a slider on a kinematic carrier; over a ground edge that the free bodies fall on."""
    s = Scene()
    g = s.body(T.STATIC_BODY, (0.0, 0.0))
    s.fixture(g, s.edge((-40.0, 0.0), (40.0, 0.0)))
    box = s.box(1.0, 0.5)
    axes = [(2.0, 1.0), (1.0, 0.0), (0.0, 1.0), (-1.0, 3.0), (1.0, 1.0), (0.6, 0.8), (5.0, 0.0), (1.0, -0.2)]
    limits = [(0.0, 4.0), (-2.0, 2.0), (1.0, 1.0), None, (-0.5, 3.0), (-3.0, 0.0), None, (0.0, 0.0)]
    motors = [(10.0, 10000.0), None, None, (1.0, 50.0), (-2.0, 20.0), None, (0.5, 0.0), (3.0, 1000.0)]
    for i in range(8):
        x = -28.0 + 8.0 * i
        b = s.body(T.DYNAMIC_BODY, (x, 10.0), angle=0.1 * i, w=0.5 * (i - 3),
                   flags=BODYDEF_DEFAULT | (BODYDEF_FIXED_ROTATION if i == 5 else 0))
        s.fixture(b, box, density=5.0)
        s.prismatic_joint(g, b, (x, 10.0), (0.0, 0.0), axes[i], reference_angle=float(F(0.1 * i)), limits=limits[i],
                          motor=motors[i], collide_connected=(i % 2 == 1))
    # slider crank
    crank = s.body(T.DYNAMIC_BODY, (0.0, 22.0))
    s.fixture(crank, s.box(0.5, 2.0), density=2.0)
    s.revolute_joint(g, crank, (0.0, 20.0), (0.0, -2.0), motor=(1.0 * np.pi, 10000.0))
    rod = s.body(T.DYNAMIC_BODY, (0.0, 28.0))
    s.fixture(rod, s.box(0.5, 4.0), density=2.0)
    s.revolute_joint(crank, rod, (0.0, 2.0), (0.0, -4.0))
    piston = s.body(T.DYNAMIC_BODY, (0.0, 32.0), flags=BODYDEF_DEFAULT | BODYDEF_FIXED_ROTATION)
    s.fixture(piston, s.box(1.5, 1.5), density=2.0)
    s.revolute_joint(rod, piston, (0.0, 4.0), (0.0, 0.0))
    s.prismatic_joint(g, piston, (0.0, 32.0), (0.0, 0.0), (0.0, 1.0), motor=(0.0, 1000.0))
    payload = s.body(T.DYNAMIC_BODY, (0.0, 38.0))
    s.fixture(payload, s.box(1.5, 1.5), density=2.0)
    # a telescope of two dynamic bodies, and a slider carried by a kinematic body
    a = s.body(T.DYNAMIC_BODY, (20.0, 3.0))
    s.fixture(a, box, density=1.0)
    b = s.body(T.DYNAMIC_BODY, (22.5, 3.0))
    s.fixture(b, box, density=1.0)
    s.prismatic_joint(a, b, (1.0, 0.0), (-1.5, 0.0), (1.0, 0.0), limits=(-0.5, 1.0), collide_connected=True)
    k = s.body(T.KINEMATIC_BODY, (-20.0, 25.0), vel=(0.5, 0.0), w=0.2)
    s.fixture(k, s.box(0.5, 0.5))
    b = s.body(T.DYNAMIC_BODY, (-20.0, 23.0))
    s.fixture(b, box, density=1.0)
    s.prismatic_joint(k, b, (0.0, -0.5), (0.0, 1.5), (0.0, -2.0), limits=(-1.0, 3.0))
    return s


def machines(seed=0):
    """Wheel, rope, friction and motor joints: the Testbed's Car (Car.h:36-230: chassis on two sprung wheels, one driven)
    on bumpy ground, a second car whose suspension is rigid (frequency 0) and undriven, a weight on a rope hung from a
    swinging arm (RopeJoint.h), boxes dragged over the ground by friction joints to it (ApplyForce.h:80-110) and a body
    steered by a motor joint (MotorJoint.h:30-80), strong and saturated."""
    s = Scene()
    g = s.body(T.STATIC_BODY, (0.0, 0.0))
    s.fixture(g, s.edge((-60.0, 0.0), (60.0, 0.0)), friction=0.6)
    for i in range(8):
        s.fixture(g, s.box(0.6, 0.15, center=(6.0 + 4.0 * i, 0.1), angle=0.2 * (i % 3 - 1)), friction=0.6)
    wheel = s.circle(0.4)
    for n, (x, hz, motor) in enumerate([(-20.0, 4.0, (-12.0, 20.0)), (-40.0, 0.0, None)]):
        chassis = s.body(T.DYNAMIC_BODY, (x, 1.0))
        s.fixture(chassis, s.polygon([(-1.5, -0.5), (1.5, -0.5), (1.5, 0.0), (0.0, 0.9), (-1.15, 0.9), (-1.5, 0.2)]), density=1.0)
        w1 = s.body(T.DYNAMIC_BODY, (x - 1.0, 0.35))
        s.fixture(w1, wheel, density=1.0, friction=0.9)
        w2 = s.body(T.DYNAMIC_BODY, (x + 1.0, 0.4))
        s.fixture(w2, wheel, density=1.0, friction=0.9)
        s.wheel_joint(chassis, w1, (-1.0, -0.65), (0.0, 0.0), (0.0, 1.0), frequency_hz=hz, damping_ratio=0.7, motor=motor)
        s.wheel_joint(chassis, w2, (1.0, -0.6), (0.0, 0.0), (0.0, 1.0), frequency_hz=hz, damping_ratio=0.7,
                      motor=(0.0, 10.0) if n == 0 else None)
    # rope: an arm on a hinge, a weight on a rope from its tip (slack at first), a second rope with coinciding anchors
    arm = s.body(T.DYNAMIC_BODY, (2.0, 14.0), w=2.0)
    s.fixture(arm, s.box(2.0, 0.125), density=20.0)
    s.revolute_joint(g, arm, (0.0, 14.0), (-2.0, 0.0))
    weight = s.body(T.DYNAMIC_BODY, (4.0, 12.5))
    s.fixture(weight, s.box(0.75, 0.75), density=10.0)
    s.rope_joint(arm, weight, (2.0, 0.0), (0.0, 0.75), 3.0)
    bead = s.body(T.DYNAMIC_BODY, (4.0, 14.0))
    s.fixture(bead, s.circle(0.2), density=1.0)
    s.rope_joint(arm, bead, (2.0, 0.0), (0.0, 0.0), 1.0, collide_connected=True)
    # friction joints to the ground: top-down style drag (gravity scale 0 would be the Testbed's; here they also rest)
    for i in range(4):
        b = s.body(T.DYNAMIC_BODY, (-8.0 + 2.5 * i, 0.5), vel=(6.0, 0.0), w=3.0 * (i - 1.5), gravity_scale=0.0)
        s.fixture(b, s.box(0.5, 0.5), density=1.0, friction=0.3)
        s.friction_joint(g, b, (-8.0 + 2.5 * i, 0.5), (0.0, 0.0), [0.0, 2.0, 10.0, 500.0][i], [0.5, 0.0, 1.0, 100.0][i],
                         collide_connected=(i % 2 == 0))
    # motor joints: one strong enough to hold its body in the air at an offset, one saturated, one between two dynamic bodies
    b = s.body(T.DYNAMIC_BODY, (30.0, 8.0))
    s.fixture(b, s.box(2.0, 0.5), density=2.0, friction=0.6)
    s.motor_joint(g, b, (33.0, 10.0), 0.5, max_force=1000.0, max_torque=1000.0, collide_connected=True)
    b = s.body(T.DYNAMIC_BODY, (40.0, 8.0))
    s.fixture(b, s.box(1.0, 0.5), density=2.0, friction=0.6)
    s.motor_joint(g, b, (40.0, 12.0), -1.0, max_force=10.0, max_torque=2.0, correction_factor=0.8, collide_connected=True)
    a = s.body(T.DYNAMIC_BODY, (50.0, 3.0))
    s.fixture(a, s.box(1.0, 1.0), density=1.0)
    b = s.body(T.DYNAMIC_BODY, (50.0, 6.0))
    s.fixture(b, s.box(0.5, 0.5), density=1.0)
    s.motor_joint(a, b, (0.0, 2.5), 0.3, max_force=200.0, max_torque=50.0, collide_connected=True)
    return s


def pulleys_and_mice(seed=0):
    """Pulley joints (Testbed/Tests/Pulleys.h:27-72: two boxes over two ground anchors; here also a block and tackle with
    ratio 2 and an unbalanced pair that runs one rope out) and mouse joints (the Testbed's drag: a box pulled towards a
    target that the test moves, one pull saturating its force limit, one on a body that is asleep at first)."""
    s = Scene()
    g = s.body(T.STATIC_BODY, (0.0, 0.0))
    s.fixture(g, s.edge((-40.0, 0.0), (40.0, 0.0)))
    s.fixture(g, s.circle(2.0, p=(-10.0, 18.0)))
    s.fixture(g, s.circle(2.0, p=(10.0, 18.0)))
    box = s.box(1.0, 2.0)
    for n, (x0, ratio, d1, d2) in enumerate([(-10.0, 1.5, 5.0, 5.0), (12.0, 1.0, 5.0, 20.0), (26.0, 2.0, 5.0, 5.0)]):
        a = s.body(T.DYNAMIC_BODY, (x0 - 4.0, 12.0))
        s.fixture(a, box, density=d1)
        b = s.body(T.DYNAMIC_BODY, (x0 + 4.0, 12.0))
        s.fixture(b, box, density=d2)
        s.pulley_joint(a, b, (x0 - 4.0, 22.0), (x0 + 4.0, 22.0), (0.0, 2.0), (0.0, 2.0), 8.0, 8.0, ratio=ratio,
                       collide_connected=(n != 1))
    small = s.box(0.5, 0.5)
    b = s.body(T.DYNAMIC_BODY, (-30.0, 5.0))
    s.fixture(b, small, density=1.0)
    s.mouse_joint(g, b, (0.25, 0.25), (-29.75, 5.25), 1000.0)
    b = s.body(T.DYNAMIC_BODY, (-25.0, 0.5))
    s.fixture(b, small, density=5.0)
    s.mouse_joint(g, b, (0.0, 0.5), (-25.0, 1.0), 10.0, frequency_hz=2.0, damping_ratio=0.1)
    b = s.body(T.DYNAMIC_BODY, (-20.0, 0.5), flags=BODYDEF_DEFAULT & ~BODYDEF_AWAKE)
    s.fixture(b, small, density=1.0)
    s.mouse_joint(g, b, (0.0, 0.0), (-20.0, 0.5), 500.0)
    return s


def gears(seed=0):
    """Gear joints (Testbed/Tests/Gears.h:27-185): two wheels on revolute joints to the ground geared r2 : r1, the larger
    one geared to a rack on a prismatic joint; a second train whose wheels ride on a dynamic frame (so that the gear
    moves four dynamic bodies), driven by a motor; a prismatic-prismatic gear (two sliders moving in opposition)."""
    s = Scene()
    g = s.body(T.STATIC_BODY, (0.0, 0.0))
    s.fixture(g, s.edge((-50.0, 0.0), (50.0, 0.0)))
    r1, r2 = 1.0, 2.0
    b1 = s.body(T.DYNAMIC_BODY, (-3.0, 12.0), w=2.0)
    s.fixture(b1, s.circle(r1), density=5.0)
    j1 = s.revolute_joint(g, b1, (-3.0, 12.0), (0.0, 0.0))
    b2 = s.body(T.DYNAMIC_BODY, (0.0, 12.0))
    s.fixture(b2, s.circle(r2), density=5.0)
    j2 = s.revolute_joint(g, b2, (0.0, 12.0), (0.0, 0.0))
    b3 = s.body(T.DYNAMIC_BODY, (2.5, 12.0))
    s.fixture(b3, s.box(0.5, 5.0), density=5.0)
    j3 = s.prismatic_joint(g, b3, (2.5, 12.0), (0.0, 0.0), (0.0, 1.0), limits=(-5.0, 5.0))
    s.gear_joint(j1, j2, r2 / r1)
    s.gear_joint(j2, j3, -1.0 / r2)
    # a gear train on a dynamic frame that stands on the ground
    frame = s.body(T.DYNAMIC_BODY, (20.0, 1.0))
    s.fixture(frame, s.box(4.0, 1.0), density=1.0, friction=0.8)
    wa = s.body(T.DYNAMIC_BODY, (18.0, 4.0))
    s.fixture(wa, s.circle(1.0), density=2.0)
    ja = s.revolute_joint(frame, wa, (-2.0, 3.0), (0.0, 0.0), motor=(3.0, 50.0))
    wb = s.body(T.DYNAMIC_BODY, (21.0, 4.0))
    s.fixture(wb, s.circle(2.0), density=2.0)
    jb = s.revolute_joint(frame, wb, (1.0, 3.0), (0.0, 0.0))
    s.gear_joint(ja, jb, 2.0, collide_connected=True)
    # two sliders in opposition
    sa = s.body(T.DYNAMIC_BODY, (-20.0, 6.0), vel=(2.0, 0.0))
    s.fixture(sa, s.box(1.0, 0.5), density=1.0)
    pa = s.prismatic_joint(g, sa, (-20.0, 6.0), (0.0, 0.0), (1.0, 0.0), limits=(-4.0, 4.0))
    sb = s.body(T.DYNAMIC_BODY, (-20.0, 9.0))
    s.fixture(sb, s.box(1.0, 0.5), density=3.0)
    pb = s.prismatic_joint(g, sb, (-20.0, 9.0), (0.0, 0.0), (3.0, 0.0))
    s.gear_joint(pa, pb, 1.0)
    return s


def resting_linkage(seed=0):
    """Jointed bodies that come to rest and fall asleep: a hinged two-bar linkage, a welded pair and a pair on a distance
    rod lying on the ground, each in its own island.  Sleeping is on."""
    s = Scene()
    g = s.body(T.STATIC_BODY, (0.0, 0.0))
    s.fixture(g, s.box(40.0, 1.0, center=(0.0, -1.0), angle=0.0), thick=True)
    bar = s.box(1.0, 0.25)
    a = s.body(T.DYNAMIC_BODY, (-10.0, 0.5))
    s.fixture(a, bar, density=1.0, friction=0.6)
    b = s.body(T.DYNAMIC_BODY, (-8.0, 0.5))
    s.fixture(b, bar, density=1.0, friction=0.6)
    s.revolute_joint(a, b, (1.0, 0.0), (-1.0, 0.0), limits=(-0.5, 0.5))
    a = s.body(T.DYNAMIC_BODY, (0.0, 0.6))
    s.fixture(a, bar, density=1.0, friction=0.6)
    b = s.body(T.DYNAMIC_BODY, (2.0, 0.6))
    s.fixture(b, bar, density=1.0, friction=0.6)
    s.weld_joint(a, b, (1.0, 0.0), (-1.0, 0.0))
    a = s.body(T.DYNAMIC_BODY, (10.0, 0.5))
    s.fixture(a, s.circle(0.5), density=1.0, friction=0.6)
    b = s.body(T.DYNAMIC_BODY, (13.0, 0.5))
    s.fixture(b, bar, density=1.0, friction=0.6)
    s.distance_joint(a, b, (0.0, 0.0), (-1.0, 0.0), 2.0)
    return s


def bullets(count=12, seed=7):
    """Continuous-collision scene: bullets and fast ordinary bodies against thin static geometry (an edge floor, a thin
    box wall, an edge ceiling) and against each other -- the situations of Testbed/Tests/BulletTest.h, ContinuousTest.h
    and TunnelingTest.h in one world.  Every contact with the static geometry is a time-of-impact candidate (no
    thick-shape fixtures), bullets are candidates against everything; at 60 Hz the fast bodies cross the wall's
    thickness in one step, so the reference only keeps them out through b2World::SolveTOI."""
    s = Scene()
    g = s.body(T.STATIC_BODY, (0.0, 0.0))
    s.fixture(g, s.edge((-30.0, 0.0), (30.0, 0.0)), density=0.0)
    s.fixture(g, s.box(0.05, 6.0, center=(10.0, 6.0), angle=0.0), density=0.0)
    s.fixture(g, s.edge((-30.0, 14.0), (30.0, 14.0)), density=0.0)
    s.fixture(g, s.box(0.05, 6.0, center=(-12.0, 6.0), angle=0.0), density=0.0)
    rnd = _Rand(seed)
    ys = rnd.uniform(1.0, 11.0, count)
    vs = rnd.uniform(60.0, 240.0, count)
    box = s.box(0.25, 0.25)
    ball = s.circle(0.2)
    tri = s.polygon(_regular_polygon(3, 0.3))
    for i in range(count):
        bullet = i % 3 != 2
        flags = BODYDEF_DEFAULT | (BODYDEF_BULLET if bullet else 0)
        direction = 1.0 if i % 2 == 0 else -1.0
        b = s.body(T.DYNAMIC_BODY, (-2.0 + 0.7 * i * direction * 0.1, float(ys[i])), angle=0.3 * i,
                   vel=(direction * float(vs[i]), -20.0 + 5.0 * (i % 5)), w=3.0 * (i % 4), flags=flags)
        s.fixture(b, (box, ball, tri)[i % 3], density=1.0 + (i % 2), restitution=0.2 * (i % 3))
    # a small heap of ordinary boxes in front of the wall that the bullets plough through
    for i in range(12):
        b = s.body(T.DYNAMIC_BODY, (6.0 + 0.55 * (i % 4), 0.3 + 0.55 * (i // 4)))
        s.fixture(b, box, density=0.5)
    return s
