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
KIND_CIRCLE, KIND_EDGE, KIND_POLYGON, KIND_BOX, KIND_RAW, KIND_CHAIN = 0, 1, 2, 3, 4, 5


class Scene:
    """Flat scene description: what user code passes to CreateBody / CreateFixture."""

    def __init__(self, gravity=(0.0, -10.0), world_flags=T.WORLD_DEFAULT):
        self.gravity = gravity
        self.world_flags = world_flags
        self.shapes = []
        self._shape_cache = {}
        self._body_chunks = []
        self._fixture_chunks = []
        self._pending_bodies = []
        self._pending_fixtures = []
        self.body_count = 0
        self.fixture_count = 0
        self._last_fixture_body = -1
        self.joints = []

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

    def chain(self, verts, loop=False):
        """b2ChainShape of up to 8 vertices: CreateLoop(verts) or CreateChain(verts); one proxy per segment"""
        assert 2 <= len(verts) <= 8
        s = np.zeros((), SHAPE_DEF)
        s["kind"] = KIND_CHAIN
        s["count"] = len(verts)
        s["flags"] = 1 if loop else 0
        s["v"][:len(verts)] = np.asarray(verts, dtype=np.float32)
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
        self._pending_bodies.append(b)
        self.body_count += 1
        return self.body_count - 1

    def fixture(self, body, shape, density=0.0, friction=0.2, restitution=0.0, sensor=False, thick=False,
                category=0x0001, mask=0xFFFF, group=0):
        assert self._last_fixture_body <= body, "fixtures must be added in body order"
        self._last_fixture_body = body
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
        self._pending_fixtures.append(f)
        self.fixture_count += 1
        return self.fixture_count - 1

    def revolute_joint(self, body_a, body_b, local_anchor_a, local_anchor_b, reference_angle=0.0, collide_connected=False,
                       limits=None, motor=None):
        """b2RevoluteJointDef with explicit local anchors (b2RevoluteJoint.h:36-86); limits = (lower, upper) enables the
        limit, motor = (speed, max torque) enables the motor.  Returns the joint id."""
        j = np.zeros((), T.JOINT)
        j["type"] = T.JOINT_REVOLUTE
        j["bodyA"], j["bodyB"] = body_a, body_b
        j["localAnchorA"] = local_anchor_a
        j["localAnchorB"] = local_anchor_b
        j["referenceAngle"] = reference_angle
        flags = T.JOINT_COLLIDE_CONNECTED if collide_connected else 0
        if limits is not None:
            flags |= T.JOINT_ENABLE_LIMIT
            j["lowerAngle"], j["upperAngle"] = limits
        if motor is not None:
            flags |= T.JOINT_ENABLE_MOTOR
            j["motorSpeed"], j["maxMotorTorque"] = motor
        j["flags"] = flags
        self.joints.append(j)
        return len(self.joints) - 1

    def _joint(self, jtype, body_a, body_b, local_anchor_a, local_anchor_b, collide_connected):
        j = np.zeros((), T.JOINT)
        j["type"] = jtype
        j["bodyA"], j["bodyB"] = body_a, body_b
        j["localAnchorA"] = local_anchor_a
        j["localAnchorB"] = local_anchor_b
        j["flags"] = T.JOINT_COLLIDE_CONNECTED if collide_connected else 0
        return j

    def distance_joint(self, body_a, body_b, local_anchor_a, local_anchor_b, length, frequency_hz=0.0, damping_ratio=0.0,
                       collide_connected=False):
        """b2DistanceJointDef (b2DistanceJoint.h:32-64): a rod of the given length, a spring-damper when frequency_hz > 0"""
        j = self._joint(T.JOINT_DISTANCE, body_a, body_b, local_anchor_a, local_anchor_b, collide_connected)
        j["length"] = length
        j["frequencyHz"] = frequency_hz
        j["dampingRatio"] = damping_ratio
        self.joints.append(j)
        return len(self.joints) - 1

    def weld_joint(self, body_a, body_b, local_anchor_a, local_anchor_b, reference_angle=0.0, frequency_hz=0.0,
                   damping_ratio=0.0, collide_connected=False):
        """b2WeldJointDef (b2WeldJoint.h:28-58): the two bodies glued at the anchor, the angle soft when frequency_hz > 0"""
        j = self._joint(T.JOINT_WELD, body_a, body_b, local_anchor_a, local_anchor_b, collide_connected)
        j["referenceAngle"] = reference_angle
        j["frequencyHz"] = frequency_hz
        j["dampingRatio"] = damping_ratio
        self.joints.append(j)
        return len(self.joints) - 1

    def prismatic_joint(self, body_a, body_b, local_anchor_a, local_anchor_b, local_axis_a, reference_angle=0.0,
                        collide_connected=False, limits=None, motor=None):
        """b2PrismaticJointDef (b2PrismaticJoint.h:30-83): body B slides along an axis fixed in body A; limits = (lower,
        upper) translation, motor = (speed, max force).  The axis need not be a unit vector."""
        j = self._joint(T.JOINT_PRISMATIC, body_a, body_b, local_anchor_a, local_anchor_b, collide_connected)
        j["axis"] = local_axis_a
        j["referenceAngle"] = reference_angle
        flags = int(j["flags"])
        if limits is not None:
            flags |= T.JOINT_ENABLE_LIMIT
            j["lowerAngle"], j["upperAngle"] = limits
        if motor is not None:
            flags |= T.JOINT_ENABLE_MOTOR
            j["motorSpeed"], j["maxMotorTorque"] = motor
        j["flags"] = flags
        self.joints.append(j)
        return len(self.joints) - 1

    def wheel_joint(self, body_a, body_b, local_anchor_a, local_anchor_b, local_axis_a, frequency_hz=2.0, damping_ratio=0.7,
                    motor=None, collide_connected=False):
        """b2WheelJointDef (b2WheelJoint.h:30-72): B's anchor stays on a line of A, held by a spring; motor = (speed,
        max torque) drives B's rotation"""
        j = self._joint(T.JOINT_WHEEL, body_a, body_b, local_anchor_a, local_anchor_b, collide_connected)
        j["axis"] = local_axis_a
        j["frequencyHz"] = frequency_hz
        j["dampingRatio"] = damping_ratio
        if motor is not None:
            j["flags"] = int(j["flags"]) | T.JOINT_ENABLE_MOTOR
            j["motorSpeed"], j["maxMotorTorque"] = motor
        self.joints.append(j)
        return len(self.joints) - 1

    def rope_joint(self, body_a, body_b, local_anchor_a, local_anchor_b, max_length, collide_connected=False):
        """b2RopeJointDef (b2RopeJoint.h:28-48)"""
        j = self._joint(T.JOINT_ROPE, body_a, body_b, local_anchor_a, local_anchor_b, collide_connected)
        j["length"] = max_length
        self.joints.append(j)
        return len(self.joints) - 1

    def friction_joint(self, body_a, body_b, local_anchor_a, local_anchor_b, max_force, max_torque, collide_connected=False):
        """b2FrictionJointDef (b2FrictionJoint.h:25-51)"""
        j = self._joint(T.JOINT_FRICTION, body_a, body_b, local_anchor_a, local_anchor_b, collide_connected)
        j["length"] = max_force
        j["maxMotorTorque"] = max_torque
        self.joints.append(j)
        return len(self.joints) - 1

    def motor_joint(self, body_a, body_b, linear_offset, angular_offset, max_force=1.0, max_torque=1.0, correction_factor=0.3,
                    collide_connected=False):
        """b2MotorJointDef (b2MotorJoint.h:25-54)"""
        j = self._joint(T.JOINT_MOTOR, body_a, body_b, (0.0, 0.0), (0.0, 0.0), collide_connected)
        j["axis"] = linear_offset
        j["referenceAngle"] = angular_offset
        j["length"] = max_force
        j["maxMotorTorque"] = max_torque
        j["dampingRatio"] = correction_factor
        self.joints.append(j)
        return len(self.joints) - 1

    def pulley_joint(self, body_a, body_b, ground_anchor_a, ground_anchor_b, local_anchor_a, local_anchor_b, length_a,
                     length_b, ratio=1.0, collide_connected=True):
        """b2PulleyJointDef (b2PulleyJoint.h:29-74)"""
        j = self._joint(T.JOINT_PULLEY, body_a, body_b, local_anchor_a, local_anchor_b, collide_connected)
        j["axis"] = ground_anchor_a
        j["lowerAngle"], j["upperAngle"] = ground_anchor_b
        j["length"] = length_a
        j["referenceAngle"] = length_b
        j["motorSpeed"] = ratio
        self.joints.append(j)
        return len(self.joints) - 1

    def mouse_joint(self, body_a, body_b, local_anchor_b, target, max_force, frequency_hz=5.0, damping_ratio=0.7):
        """b2MouseJointDef (b2MouseJoint.h:28-60).  local_anchor_b must be the target in body B's frame at creation
        (which is what the reference's constructor computes from the target)."""
        j = self._joint(T.JOINT_MOUSE, body_a, body_b, (0.0, 0.0), local_anchor_b, False)
        j["axis"] = target
        j["length"] = max_force
        j["frequencyHz"] = frequency_hz
        j["dampingRatio"] = damping_ratio
        self.joints.append(j)
        return len(self.joints) - 1

    def gear_joint(self, joint1, joint2, ratio, collide_connected=False):
        """b2GearJointDef (b2GearJoint.h:28-53) over two earlier revolute / prismatic joints of this scene; everything
        else the reference's constructor derives from them."""
        j1, j2 = self.joints[joint1], self.joints[joint2]
        j = self._joint(T.JOINT_GEAR, int(j1["bodyB"]), int(j2["bodyB"]), (0.0, 0.0), (0.0, 0.0), collide_connected)
        j["frequencyHz"], j["dampingRatio"] = joint1, joint2
        j["motorSpeed"] = ratio
        j["limitState"], j["reserved"] = int(j1["bodyA"]), int(j2["bodyA"])
        self.joints.append(j)
        return len(self.joints) - 1

    def joint_array(self):
        return np.array(self.joints, dtype=T.JOINT) if self.joints else np.zeros(0, T.JOINT)

    def _flush(self):
        if self._pending_bodies:
            self._body_chunks.append(np.array(self._pending_bodies, dtype=BODY_DEF))
            self._pending_bodies = []
        if self._pending_fixtures:
            self._fixture_chunks.append(np.array(self._pending_fixtures, dtype=FIXTURE_DEF))
            self._pending_fixtures = []

    def add_bulk(self, bodies, fixtures):
        """Append many bodies at once (vectorised scene builders).  `bodies` is a BODY_DEF array; `fixtures` a
        FIXTURE_DEF array sorted by body, whose `body` field is relative to the first body of this batch."""
        self._flush()
        base = self.body_count
        fx = np.array(fixtures, dtype=FIXTURE_DEF, copy=True)
        fx["body"] += base
        self._body_chunks.append(np.ascontiguousarray(bodies, dtype=BODY_DEF))
        self._fixture_chunks.append(fx)
        self.body_count += len(bodies)
        self.fixture_count += len(fx)
        if len(fx):
            self._last_fixture_body = int(fx["body"][-1])
        return base

    def arrays(self):
        self._flush()
        b = np.concatenate(self._body_chunks) if self._body_chunks else np.zeros(0, BODY_DEF)
        f = np.concatenate(self._fixture_chunks) if self._fixture_chunks else np.zeros(0, FIXTURE_DEF)
        s = np.array(self.shapes, dtype=SHAPE_DEF) if self.shapes else np.zeros(0, SHAPE_DEF)
        return b, s, f
