"""numpy dtypes mirroring the plain-C records of include/b2cuda.h (same field order, no padding)."""
import numpy as np

BODY = np.dtype([
    ("px", "f4"), ("py", "f4"), ("qs", "f4"), ("qc", "f4"),
    ("cx", "f4"), ("cy", "f4"), ("a", "f4"),
    ("c0x", "f4"), ("c0y", "f4"), ("a0", "f4"), ("alpha0", "f4"),
    ("lcx", "f4"), ("lcy", "f4"),
    ("vx", "f4"), ("vy", "f4"), ("w", "f4"),
    ("fx", "f4"), ("fy", "f4"), ("torque", "f4"),
    ("invMass", "f4"), ("invI", "f4"),
    ("linearDamping", "f4"), ("angularDamping", "f4"), ("gravityScale", "f4"),
    ("sleepTime", "f4"),
    ("flags", "u4"),
])
assert BODY.itemsize == 104

BODY_STATE = np.dtype([
    ("px", "f4"), ("py", "f4"), ("qs", "f4"), ("qc", "f4"),
    ("cx", "f4"), ("cy", "f4"), ("a", "f4"),
    ("vx", "f4"), ("vy", "f4"), ("w", "f4"),
    ("sleepTime", "f4"), ("flags", "u4"),
])
assert BODY_STATE.itemsize == 48

SHAPE = np.dtype([
    ("type", "i4"), ("count", "i4"), ("radius", "f4"), ("flags", "u4"),
    ("v", "f4", (8, 2)), ("n", "f4", (8, 2)), ("centroid", "f4", (2,)), ("pad", "f4", (2,)),
])
assert SHAPE.itemsize == 160

PROXY = np.dtype([
    ("aabb", "f4", (4,)), ("fat", "f4", (4,)),
    ("body", "i4"), ("shape", "i4"),
    ("friction", "f4"), ("restitution", "f4"),
    ("categoryBits", "u2"), ("maskBits", "u2"), ("groupIndex", "i2"), ("flags", "u2"),
    ("fixture", "i4"), ("child", "i4"),
])
assert PROXY.itemsize == 64

MANIFOLD = np.dtype([
    ("localNormal", "f4", (2,)), ("localPoint", "f4", (2,)),
    ("points", [("localPoint", "f4", (2,)), ("normalImpulse", "f4"), ("tangentImpulse", "f4")], (2,)),
    ("id", "u4", (2,)), ("type", "i4"), ("pointCount", "i4"),
])
assert MANIFOLD.itemsize == 64

CONTACT = np.dtype([
    ("proxyA", "i4"), ("proxyB", "i4"), ("flags", "u4"),
    ("friction", "f4"), ("restitution", "f4"), ("tangentSpeed", "f4"),
    ("toiCount", "i4"), ("toi", "f4"),
    ("manifold", MANIFOLD),
    ("stamp", "u4"), ("reserved", "u4"),
])
assert CONTACT.itemsize == 104

DISTANCE_RESULT = np.dtype([("distance", "f4"), ("pointA", "f4", (2,)), ("pointB", "f4", (2,)), ("iterations", "i4")])
assert DISTANCE_RESULT.itemsize == 24
SWEEP = np.dtype([("localCenter", "f4", (2,)), ("c0", "f4", (2,)), ("c", "f4", (2,)), ("a0", "f4"), ("a", "f4"), ("alpha0", "f4")])
assert SWEEP.itemsize == 36
TOI_RESULT = np.dtype([("state", "i4"), ("t", "f4")])
assert TOI_RESULT.itemsize == 8
TOI_UNKNOWN, TOI_FAILED, TOI_OVERLAPPED, TOI_TOUCHING, TOI_SEPARATED = range(5)

WORLD_DEF = np.dtype([
    ("device", "i4"), ("gravity", "f4", (2,)), ("flags", "u4"),
    ("bodyCapacity", "i4"), ("proxyCapacity", "i4"), ("shapeCapacity", "i4"), ("contactCapacity", "i4"),
])

SHARD_LINK = np.dtype([
    ("ipcHandle", "u1", (64,)), ("localPointer", "u8"), ("processId", "i4"), ("device", "i4"),
    ("ghostCount", "i4"), ("exportCount", "i4"), ("rank", "i4"), ("rankCount", "i4"),
])
assert SHARD_LINK.itemsize == 96

STEP_INFO = np.dtype([
    ("step", "f4"), ("collide", "f4"), ("solve", "f4"), ("solveTraversal", "f4"), ("solveInit", "f4"),
    ("solveVelocity", "f4"), ("solvePosition", "f4"), ("solveTOI", "f4"), ("broadphase", "f4"),
    ("broadphaseSyncFixtures", "f4"), ("broadphaseFindContacts", "f4"), ("locking", "f4"), ("reserved", "f4"),
    ("bodyCount", "i4"), ("proxyCount", "i4"), ("contactCount", "i4"),
    ("touchingCount", "i4"), ("constraintCount", "i4"), ("colourCount", "i4"), ("overflowCount", "i4"),
    ("islandBodyCount", "i4"), ("awakeBodyCount", "i4"), ("moveCount", "i4"),
    ("newContactCount", "i4"), ("destroyedContactCount", "i4"),
    ("beginCount", "i4"), ("endCount", "i4"), ("toiCandidateCount", "i4"), ("kernelLaunches", "i4"),
    ("toiEventPending", "i4"), ("toiMinKey", "u8"), ("toiMinAlpha", "f4"), ("toiSubSteps", "i4"),
    ("toiEventCount", "i4"), ("toiNewContactCount", "i4"),
])
assert STEP_INFO.itemsize == 144 and STEP_INFO.fields["toiMinKey"][1] == 120

JOINT = np.dtype([
    ("type", "i4"), ("bodyA", "i4"), ("bodyB", "i4"), ("flags", "u4"),
    ("localAnchorA", "f4", (2,)), ("localAnchorB", "f4", (2,)),
    ("referenceAngle", "f4"), ("lowerAngle", "f4"), ("upperAngle", "f4"), ("maxMotorTorque", "f4"), ("motorSpeed", "f4"),
    ("length", "f4"), ("frequencyHz", "f4"), ("dampingRatio", "f4"), ("axis", "f4", (2,)),
    ("lastSolve", "f4", (4,)), ("work", "f4", (4,)), ("impulse", "f4", (3,)), ("motorImpulse", "f4"), ("limitState", "i4"),
    ("reserved", "i4"),
])
assert JOINT.itemsize == 128
JOINT_REVOLUTE, JOINT_PRISMATIC, JOINT_DISTANCE, JOINT_WELD = 1, 2, 3, 8
JOINT_WHEEL, JOINT_FRICTION, JOINT_ROPE, JOINT_MOTOR = 7, 9, 10, 11
JOINT_PULLEY, JOINT_MOUSE, JOINT_GEAR = 4, 5, 6
JOINT_GEAR_PRISMATIC_1, JOINT_GEAR_PRISMATIC_2 = 0x100, 0x200
JOINT_COLLIDE_CONNECTED, JOINT_ENABLE_LIMIT, JOINT_ENABLE_MOTOR = 1, 2, 4

# enums
STATIC_BODY, KINEMATIC_BODY, DYNAMIC_BODY = 0, 1, 2
BODY_TYPE_MASK = 0x3
BODY_ISLAND, BODY_AWAKE, BODY_AUTOSLEEP, BODY_BULLET, BODY_FIXED_ROTATION, BODY_ACTIVE = 0x4, 0x8, 0x10, 0x20, 0x40, 0x80
BODY_GHOST = 0x100
SHAPE_CIRCLE, SHAPE_EDGE, SHAPE_POLYGON = 0, 1, 2
EDGE_HAS_VERTEX0, EDGE_HAS_VERTEX3, EDGE_CHAIN_CHILD = 1, 2, 4
PROXY_SENSOR, PROXY_THICK, PROXY_MOVED = 1, 2, 4
PROXY_NEW, PROXY_REFILTER, PROXY_INACTIVE = 0x10, 0x20, 0x40
CONTACT_ISLAND, CONTACT_TOUCHING, CONTACT_ENABLED, CONTACT_FILTER = 0x1, 0x2, 0x4, 0x8
CONTACT_BULLET_HIT, CONTACT_TOI, CONTACT_TOI_CANDIDATE, CONTACT_INACTIVE = 0x10, 0x20, 0x40, 0x80
MANIFOLD_CIRCLES, MANIFOLD_FACE_A, MANIFOLD_FACE_B = 0, 1, 2
WORLD_ALLOW_SLEEP, WORLD_WARM_STARTING, WORLD_CONTINUOUS, WORLD_SUB_STEPPING, WORLD_CLEAR_FORCES = 1, 2, 4, 8, 16
WORLD_DEFAULT = WORLD_ALLOW_SLEEP | WORLD_WARM_STARTING | WORLD_CONTINUOUS | WORLD_CLEAR_FORCES
EVENT_BEGIN, EVENT_END = 0, 1

OK, ERR_CUDA, ERR_CAPACITY, ERR_ARGUMENT, ERR_UNSUPPORTED, ERR_NO_DEVICE = 0, -1, -2, -3, -4, -5


def contact_keys(contacts):
    """(min<<32 | max) proxy-id key of each b2cuContact record."""
    a = contacts["proxyA"].astype(np.uint64)
    b = contacts["proxyB"].astype(np.uint64)
    lo = np.minimum(a, b)
    hi = np.maximum(a, b)
    return (lo << np.uint64(32)) | hi
