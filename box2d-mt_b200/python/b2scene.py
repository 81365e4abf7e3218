"""Flat scene descriptions: the arguments user code passes to b2World::CreateBody / b2Body::CreateFixture,
as numpy records (reference: Box2D/Dynamics/b2Body.h:52-127 b2BodyDef, Box2D/Dynamics/b2Fixture.h:56-98
b2FixtureDef).  Consumed by the host-API binding (bulk create) and, in tests, by the oracle."""
import numpy as np

import b2cuda_types as T

BODY_DEF = np.dtype([
    ("type", "i4"), ("px", "f4"), ("py", "f4"), ("angle", "f4"),
    ("vx", "f4"), ("vy", "f4"), ("w", "f4"),
    ("linearDamping", "f4"), ("angularDamping", "f4"), ("gravityScale", "f4"),
    ("flags", "u4"),
])
SHAPE_DEF = np.dtype([
    ("kind", "i4"), ("count", "i4"), ("radius", "f4"), ("flags", "u4"),
    ("v", "f4", (8, 2)), ("n", "f4", (8, 2)), ("centroid", "f4", (2,)),
])
FIXTURE_DEF = np.dtype([
    ("body", "i4"), ("shape", "i4"), ("density", "f4"), ("friction", "f4"), ("restitution", "f4"),
    ("flags", "u4"), ("categoryBits", "u2"), ("maskBits", "u2"), ("groupIndex", "i2"), ("pad", "u2"),
])

BODYDEF_ALLOW_SLEEP, BODYDEF_AWAKE, BODYDEF_FIXED_ROTATION, BODYDEF_BULLET, BODYDEF_ACTIVE = 1, 2, 4, 8, 16
BODYDEF_DEFAULT = BODYDEF_ALLOW_SLEEP | BODYDEF_AWAKE | BODYDEF_ACTIVE
KIND_CIRCLE, KIND_EDGE, KIND_POLYGON, KIND_BOX, KIND_RAW = 0, 1, 2, 3, 4


class Scene:
    """Flat scene description: what user code passes to CreateBody / CreateFixture."""

    def __init__(self, gravity=(0.0, -10.0), world_flags=T.WORLD_DEFAULT):
        self.gravity = gravity
        self.world_flags = world_flags
        self.bodies = []
        self.shapes = []
        self.fixtures = []
        self._shape_cache = {}

    # -- shapes ------------------------------------------------------------------
    def _add_shape(self, rec):
        key = rec.tobytes()
        idx = self._shape_cache.get(key)
        if idx is None:
            idx = len(self.shapes)
            self.shapes.append(rec)
            self._shape_cache[key] = idx
        return idx

    def circle(self, radius, p=(0.0, 0.0)):
        s = np.zeros((), SHAPE_DEF)
        s["kind"] = KIND_CIRCLE
        s["radius"] = radius
        s["v"][0] = p
        return self._add_shape(s)

    def box(self, hx, hy, center=None, angle=0.0):
        s = np.zeros((), SHAPE_DEF)
        s["kind"] = KIND_BOX
        s["v"][0] = (hx, hy)
        if center is not None:
            s["flags"] = 1
            s["v"][1] = center
            s["v"][2] = (angle, 0.0)
        return self._add_shape(s)

    def polygon(self, verts):
        s = np.zeros((), SHAPE_DEF)
        s["kind"] = KIND_POLYGON
        s["count"] = len(verts)
        s["v"][:len(verts)] = np.asarray(verts, dtype=np.float32)
        return self._add_shape(s)

    def edge(self, v1, v2, v0=None, v3=None):
        s = np.zeros((), SHAPE_DEF)
        s["kind"] = KIND_EDGE
        s["v"][0] = v1
        s["v"][1] = v2
        fl = 0
        if v0 is not None:
            s["v"][2] = v0
            fl |= T.EDGE_HAS_VERTEX0
        if v3 is not None:
            s["v"][3] = v3
            fl |= T.EDGE_HAS_VERTEX3
        s["flags"] = fl
        return self._add_shape(s)

    # -- bodies / fixtures -------------------------------------------------------
    def body(self, btype, pos, angle=0.0, vel=(0.0, 0.0), w=0.0, linear_damping=0.0, angular_damping=0.0,
             gravity_scale=1.0, flags=BODYDEF_DEFAULT):
        b = np.zeros((), BODY_DEF)
        b["type"] = btype
        b["px"], b["py"] = pos
        b["angle"] = angle
        b["vx"], b["vy"] = vel
        b["w"] = w
        b["linearDamping"] = linear_damping
        b["angularDamping"] = angular_damping
        b["gravityScale"] = gravity_scale
        b["flags"] = flags
        self.bodies.append(b)
        return len(self.bodies) - 1

    def fixture(self, body, shape, density=0.0, friction=0.2, restitution=0.0, sensor=False, thick=False,
                category=0x0001, mask=0xFFFF, group=0):
        assert not self.fixtures or self.fixtures[-1]["body"] <= body, "fixtures must be added in body order"
        f = np.zeros((), FIXTURE_DEF)
        f["body"] = body
        f["shape"] = shape
        f["density"] = density
        f["friction"] = friction
        f["restitution"] = restitution
        f["flags"] = (T.PROXY_SENSOR if sensor else 0) | (T.PROXY_THICK if thick else 0)
        f["categoryBits"] = category
        f["maskBits"] = mask
        f["groupIndex"] = group
        self.fixtures.append(f)
        return len(self.fixtures) - 1

    def arrays(self):
        b = np.array(self.bodies, dtype=BODY_DEF) if self.bodies else np.zeros(0, BODY_DEF)
        s = np.array(self.shapes, dtype=SHAPE_DEF) if self.shapes else np.zeros(0, SHAPE_DEF)
        f = np.array(self.fixtures, dtype=FIXTURE_DEF) if self.fixtures else np.zeros(0, FIXTURE_DEF)
        return b, s, f
