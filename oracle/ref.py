"""oracle/ref.py -- TEST INFRASTRUCTURE: ctypes wrapper over oracle/_ref/libb2ref.so (the compiled reference).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
sys.path.insert(0, os.path.join(_ROOT, "box2d-mt_b200", "python"))
import b2cuda_types as T  # noqa: E402

from b2scene import (BODY_DEF, SHAPE_DEF, FIXTURE_DEF, Scene, BODYDEF_ALLOW_SLEEP, BODYDEF_AWAKE,  # noqa: E402,F401
                     BODYDEF_FIXED_ROTATION, BODYDEF_BULLET, BODYDEF_ACTIVE, BODYDEF_DEFAULT,
                     KIND_CIRCLE, KIND_EDGE, KIND_POLYGON, KIND_BOX, KIND_RAW)

_libs = {}


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def load(stock=False):
    """Load (building first if the reference tree is present) the oracle library."""
    if stock in _libs:
        return _libs[stock]
    sys.path.insert(0, _HERE)
    import build_ref
    build_ref.build()
    if stock == "mt":
        # the reference with b2_maxThreads raised to 32 (build_ref.build_mt): timing rows of bench.py only
        path = build_ref.build_mt()
        if path is None:
            raise RuntimeError("oracle/_ref/libb2ref_mt32.so was not prebuilt")
        lib = ctypes.CDLL(path)
    else:
        lib = ctypes.CDLL(build_ref.lib_path(stock))
    vp, i32, f32, u32 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_float, ctypes.c_uint32
    lib.b2ref_create.restype = vp
    lib.b2ref_create.argtypes = [f32, f32, u32, i32]
    lib.b2ref_destroy.argtypes = [vp]
    lib.b2ref_build.argtypes = [vp, i32, vp, i32, vp, i32, vp]
    lib.b2ref_step.argtypes = [vp, f32, i32, i32]
    lib.b2ref_step_ordered.argtypes = [vp, f32, i32, i32, i32, vp]
    lib.b2ref_counts.argtypes = [vp, vp, vp, vp]
    lib.b2ref_inv_dt0.restype = f32
    lib.b2ref_inv_dt0.argtypes = [vp]
    lib.b2ref_export_bodies.argtypes = [vp, vp]
    lib.b2ref_export_shapes.argtypes = [vp, vp]
    lib.b2ref_export_proxies.argtypes = [vp, vp, vp]
    lib.b2ref_export_contacts.argtypes = [vp, i32, vp]
    lib.b2ref_events.argtypes = [vp, i32, i32, vp]
    lib.b2ref_toi_candidates.argtypes = [vp, i32, vp]
    lib.b2ref_first_toi.argtypes = [vp, vp, vp]
    lib.b2ref_first_toi.restype = i32
    lib.b2ref_set_joints.argtypes = [vp, i32, vp]
    lib.b2ref_set_joints.restype = i32
    lib.b2ref_set_joint_order.argtypes = [vp, i32, vp]
    lib.b2ref_export_joints.argtypes = [vp, vp]
    lib.b2ref_joint_set_motor.argtypes = [vp, i32, i32, f32, f32]
    lib.b2ref_joint_set_limits.argtypes = [vp, i32, i32, f32, f32]
    lib.b2ref_destroy_joint.argtypes = [vp, i32]
    lib.b2ref_joint_set_target.argtypes = [vp, i32, f32, f32]
    lib.b2ref_joint_set_spring.argtypes = [vp, i32, f32, f32, f32]
    lib.b2ref_joint_readings.argtypes = [vp, f32, vp]
    lib.b2ref_profile.argtypes = [vp, vp]
    lib.b2ref_set_transform.argtypes = [vp, i32, f32, f32, f32]
    lib.b2ref_set_type.argtypes = [vp, i32, i32]
    lib.b2ref_set_active.argtypes = [vp, i32, i32]
    lib.b2ref_query_aabb.argtypes = [vp, vp, i32, vp]
    lib.b2ref_query_aabb.restype = i32
    lib.b2ref_ray_cast_closest.argtypes = [vp, vp, vp, vp]
    lib.b2ref_ray_cast_closest.restype = i32
    lib.b2ref_record_post_solve.argtypes = [vp, i32]
    lib.b2ref_set_pre_solve_rule.argtypes = [vp, i32]
    lib.b2ref_pre_solve_digest.argtypes = [vp, vp, vp]
    lib.b2ref_post_solve_digest.argtypes = [vp, vp, vp]
    lib.b2ref_set_modulo_filter.argtypes = [vp, i32]
    lib.b2ref_set_filter.argtypes = [vp, i32, ctypes.c_uint16, ctypes.c_uint16, ctypes.c_int16]
    lib.b2ref_set_velocity.argtypes = [vp, i32, f32, f32, f32]
    lib.b2ref_apply_force.argtypes = [vp, i32, f32, f32, f32]
    lib.b2ref_set_awake.argtypes = [vp, i32, i32]
    lib.b2ref_set_body_param.argtypes = [vp, i32, i32, f32]
    lib.b2ref_destroy_last_fixture.argtypes = [vp]
    lib.b2ref_hash.restype = u32
    lib.b2ref_hash.argtypes = [vp]
    lib.b2ref_collide.argtypes = [vp, vp, vp, vp, vp]
    lib.b2ref_sincos.argtypes = [f32, vp, vp]
    _libs[stock] = lib
    return lib


class RefWorld:
    """A reference b2World built from a Scene (or raw def arrays)."""

    def __init__(self, scene=None, threads=1, stock_libm=False, gravity=None, world_flags=None, arrays=None):
        self.lib = load(stock_libm)
        if scene is not None:
            gravity = scene.gravity
            world_flags = scene.world_flags
            arrays = scene.arrays()
        self.h = self.lib.b2ref_create(gravity[0], gravity[1], world_flags, threads)
        b, s, f = arrays
        b = np.ascontiguousarray(b, BODY_DEF)
        s = np.ascontiguousarray(s, SHAPE_DEF)
        f = np.ascontiguousarray(f, FIXTURE_DEF)
        rc = self.lib.b2ref_build(self.h, len(b), _ptr(b), len(s), _ptr(s), len(f), _ptr(f))
        if rc != 0:
            raise RuntimeError("b2ref_build failed: %d" % rc)
        self.gravity = gravity
        self.world_flags = world_flags
        self.joint_count = 0
        if scene is not None and getattr(scene, "joints", None):
            self.set_joints(scene.joint_array())

    def add(self, scene):
        """CreateBody / CreateFixture for every body of another Scene (appended; between steps)"""
        b, s, f = scene.arrays()
        rc = self.lib.b2ref_build(self.h, len(b), _ptr(np.ascontiguousarray(b, BODY_DEF)), len(s),
                                  _ptr(np.ascontiguousarray(s, SHAPE_DEF)), len(f), _ptr(np.ascontiguousarray(f, FIXTURE_DEF)))
        if rc != 0:
            raise RuntimeError("b2ref_build failed: %d" % rc)

    def set_joints(self, joints):
        j = np.ascontiguousarray(joints, T.JOINT)
        if self.lib.b2ref_set_joints(self.h, len(j), _ptr(j)) != 0:
            raise RuntimeError("b2ref_set_joints failed")
        self.joint_count = len(j)

    def set_joint_order(self, ids):
        ids = np.ascontiguousarray(ids, np.int32)
        assert len(ids) == self.joint_count
        self.lib.b2ref_set_joint_order(self.h, len(ids), _ptr(ids))

    def joint_set_motor(self, joint, enable, speed, max_torque):
        self.lib.b2ref_joint_set_motor(self.h, joint, int(enable), ctypes.c_float(speed), ctypes.c_float(max_torque))

    def joint_set_limits(self, joint, enable, lower, upper):
        self.lib.b2ref_joint_set_limits(self.h, joint, int(enable), ctypes.c_float(lower), ctypes.c_float(upper))

    def joint_set_spring(self, joint, length, frequency_hz, damping_ratio):
        self.lib.b2ref_joint_set_spring(self.h, joint, ctypes.c_float(length), ctypes.c_float(frequency_hz),
                                        ctypes.c_float(damping_ratio))

    def joint_set_target(self, joint, x, y):
        self.lib.b2ref_joint_set_target(self.h, joint, ctypes.c_float(x), ctypes.c_float(y))

    def destroy_joint(self, joint):
        self.lib.b2ref_destroy_joint(self.h, joint)
        self.joint_count -= 1

    def joint_readings(self, inv_dt=60.0):
        out = np.zeros((self.joint_count, 6), np.float32)
        if self.joint_count:
            self.lib.b2ref_joint_readings(self.h, ctypes.c_float(inv_dt), _ptr(out))
        return out

    def joints(self):
        out = np.zeros(self.joint_count, T.JOINT)
        if self.joint_count:
            self.lib.b2ref_export_joints(self.h, _ptr(out))
        return out

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.b2ref_destroy(self.h)
            self.h = None

    def counts(self):
        a, b, c = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
        self.lib.b2ref_counts(self.h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
        return a.value, b.value, c.value

    def step(self, dt=1.0 / 60.0, vel_iters=8, pos_iters=3):
        self.lib.b2ref_step(self.h, dt, vel_iters, pos_iters)

    def step_ordered(self, keys, dt=1.0 / 60.0, vel_iters=8, pos_iters=3):
        keys = np.ascontiguousarray(keys, np.uint64)
        return self.lib.b2ref_step_ordered(self.h, dt, vel_iters, pos_iters, len(keys), _ptr(keys))

    def inv_dt0(self):
        return self.lib.b2ref_inv_dt0(self.h)

    def bodies(self):
        out = np.zeros(self.counts()[0], T.BODY)
        self.lib.b2ref_export_bodies(self.h, _ptr(out))
        return out

    def shapes(self):
        out = np.zeros(self.counts()[1], T.SHAPE)
        self.lib.b2ref_export_shapes(self.h, _ptr(out))
        return out

    def proxies(self, with_tree_ids=False):
        n = self.counts()[1]
        out = np.zeros(n, T.PROXY)
        ids = np.zeros(n, np.int32)
        self.lib.b2ref_export_proxies(self.h, _ptr(out), _ptr(ids))
        return (out, ids) if with_tree_ids else out

    def contacts(self):
        n = self.counts()[2]
        out = np.zeros(n, T.CONTACT)
        m = self.lib.b2ref_export_contacts(self.h, n, _ptr(out))
        assert m == n
        return out

    def events(self, kind):
        cap = 1 << 16
        while True:
            out = np.zeros(cap, np.uint64)
            n = self.lib.b2ref_events(self.h, kind, cap, _ptr(out))
            if n <= cap:
                return out[:n]
            cap = n

    def first_toi(self):
        """(key, alpha) of the contact b2World::SolveTOI would pick first on the current state, or (None, 1.0)."""
        key = ctypes.c_uint64()
        alpha = ctypes.c_float()
        found = self.lib.b2ref_first_toi(self.h, ctypes.byref(key), ctypes.byref(alpha))
        return (int(key.value) if found else None), float(alpha.value)

    def toi_candidates(self):
        cap = max(16, self.counts()[2])
        out = np.zeros(cap, np.uint64)
        n = self.lib.b2ref_toi_candidates(self.h, cap, _ptr(out))
        return out[:n]

    def profile(self):
        out = np.zeros(13, np.float32)
        self.lib.b2ref_profile(self.h, _ptr(out))
        return out

    def hash(self):
        return self.lib.b2ref_hash(self.h)

    def set_transform(self, body, x, y, angle):
        self.lib.b2ref_set_transform(self.h, body, x, y, angle)

    def set_modulo_filter(self, modulus):
        self.lib.b2ref_set_modulo_filter(self.h, modulus)

    def set_pre_solve_rule(self, modulus):
        self.lib.b2ref_set_pre_solve_rule(self.h, modulus)

    def pre_solve_digest(self):
        d = np.zeros(1, np.uint64)
        c = np.zeros(1, np.int64)
        self.lib.b2ref_pre_solve_digest(self.h, _ptr(d), _ptr(c))
        return int(d[0]), int(c[0])

    def record_post_solve(self, on=True):
        self.lib.b2ref_record_post_solve(self.h, int(on))

    def post_solve_digest(self):
        d = np.zeros(1, np.uint64)
        c = np.zeros(1, np.int64)
        self.lib.b2ref_post_solve_digest(self.h, _ptr(d), _ptr(c))
        return int(d[0]), int(c[0])

    def query_aabb(self, box):
        a = np.asarray(box, np.float32)
        out = np.zeros(1 << 16, np.int32)
        n = self.lib.b2ref_query_aabb(self.h, _ptr(a), len(out), _ptr(out))
        return out[:n]

    def ray_cast_closest(self, p1, p2):
        a, b = np.asarray(p1, np.float32), np.asarray(p2, np.float32)
        out = np.zeros(5, np.float32)
        proxy = self.lib.b2ref_ray_cast_closest(self.h, _ptr(a), _ptr(b), _ptr(out))
        return proxy, out

    def set_active(self, body, on):
        self.lib.b2ref_set_active(self.h, body, int(on))

    def set_type(self, body, body_type):
        self.lib.b2ref_set_type(self.h, body, body_type)

    def set_filter(self, fixture, category, mask, group):
        self.lib.b2ref_set_filter(self.h, fixture, category, mask, group)

    def set_velocity(self, body, vx, vy, w):
        self.lib.b2ref_set_velocity(self.h, body, vx, vy, w)

    def apply_force(self, body, fx, fy, torque):
        self.lib.b2ref_apply_force(self.h, body, fx, fy, torque)

    def set_awake(self, body, awake):
        self.lib.b2ref_set_awake(self.h, body, int(awake))

    def set_body_param(self, body, which, value):
        self.lib.b2ref_set_body_param(self.h, body, which, float(value))

    def destroy_last_fixture(self):
        self.lib.b2ref_destroy_last_fixture(self.h)


def collide(shape_a, xf_a, shape_b, xf_b, stock_libm=False):
    """Reference manifold for two b2cuShape records and transforms (p.x, p.y, sin, cos)."""
    lib = load(stock_libm)
    sa = np.ascontiguousarray(shape_a, T.SHAPE)
    sb = np.ascontiguousarray(shape_b, T.SHAPE)
    xa = np.ascontiguousarray(xf_a, np.float32)
    xb = np.ascontiguousarray(xf_b, np.float32)
    out = np.zeros((), T.MANIFOLD)
    lib.b2ref_collide(_ptr(sa), _ptr(xa), _ptr(sb), _ptr(xb), _ptr(out))
    return out


def distance(shape_a, xf_a, shape_b, xf_b, use_radii=True):
    """Reference b2Distance (cold cache) for two b2cuShape records and transforms (p.x, p.y, sin, cos)."""
    lib = load(False)
    sa = np.ascontiguousarray(shape_a, T.SHAPE)
    sb = np.ascontiguousarray(shape_b, T.SHAPE)
    xa = np.ascontiguousarray(xf_a, np.float32)
    xb = np.ascontiguousarray(xf_b, np.float32)
    out = np.zeros((), T.DISTANCE_RESULT)
    lib.b2ref_distance(_ptr(sa), _ptr(xa), _ptr(sb), _ptr(xb), 1 if use_radii else 0, _ptr(out))
    return out


def time_of_impact(shape_a, sweep_a, shape_b, sweep_b, t_max=1.0):
    """Reference b2TimeOfImpact for two b2cuShape records and SWEEP records; returns a TOI_RESULT scalar."""
    lib = load(False)
    sa = np.ascontiguousarray(shape_a, T.SHAPE)
    sb = np.ascontiguousarray(shape_b, T.SHAPE)
    wa = np.ascontiguousarray(sweep_a, T.SWEEP)
    wb = np.ascontiguousarray(sweep_b, T.SWEEP)
    out = np.zeros((), T.TOI_RESULT)
    lib.b2ref_time_of_impact(_ptr(sa), _ptr(wa), _ptr(sb), _ptr(wb), ctypes.c_float(t_max), _ptr(out))
    return out


def sincos(x, stock_libm=False):
    lib = load(stock_libm)
    s, c = ctypes.c_float(), ctypes.c_float()
    lib.b2ref_sincos(np.float32(x), ctypes.byref(s), ctypes.byref(c))
    return np.float32(s.value), np.float32(c.value)
