"""ctypes binding of libbox2d_b200.so's flat C interface (host/capi/b2host_capi.cpp): the C++ host API --
b2World, b2Body::CreateFixture, b2World::Step(dt, vIters, pIters, b2CudaStepExecutor&) -- driven from Python."""
import ctypes
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.dirname(_HERE)
sys.path.insert(0, _HERE)
import b2cuda_types as T  # noqa: E402
from b2scene import BODY_DEF, SHAPE_DEF, FIXTURE_DEF  # noqa: E402

_lib = None


def lib_path():
    return os.path.join(_PKG, "libbox2d_b200.so")


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(lib_path()):
        import importlib.util
        spec = importlib.util.spec_from_file_location("b2cuda_build", os.path.join(_PKG, "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build_all()
    lib = ctypes.CDLL(lib_path())
    vp, i32, f32, u32 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_float, ctypes.c_uint32
    lib.b2h_create.restype = vp
    lib.b2h_create.argtypes = [f32, f32, u32, i32, i32, i32]
    lib.b2h_destroy.argtypes = [vp]
    lib.b2h_shard_configure.argtypes = [vp, i32, i32, vp, i32, vp, i32, f32]
    lib.b2h_shard_link.argtypes = [vp, vp]
    lib.b2h_shard_connect.argtypes = [vp, vp, vp]
    lib.b2h_set_options.argtypes = [vp, i32, i32]
    lib.b2h_apply_force_range.argtypes = [vp, i32, i32, f32, f32]
    lib.b2h_build.argtypes = [vp, i32, vp, i32, vp, i32, vp]
    lib.b2h_step.argtypes = [vp, f32, i32, i32]
    lib.b2h_last_error.argtypes = [vp]
    lib.b2h_last_error.restype = ctypes.c_char_p
    lib.b2h_counts.argtypes = [vp, vp, vp, vp]
    lib.b2h_get_bodies.argtypes = [vp, vp]
    lib.b2h_get_proxies.argtypes = [vp, vp]
    lib.b2h_get_transforms.argtypes = [vp, vp, vp]
    lib.b2h_get_mass.argtypes = [vp, vp]
    lib.b2h_get_contacts.argtypes = [vp, i32, vp, vp, vp]
    lib.b2h_events.argtypes = [vp, i32, i32, vp]
    lib.b2h_solver_order.argtypes = [vp, i32, vp]
    lib.b2h_create_joints.argtypes = [vp, i32, vp]
    lib.b2h_destroy_joint.argtypes = [vp, i32]
    lib.b2h_joint_readings.argtypes = [vp, f32, vp]
    lib.b2h_joint_set_motor.argtypes = [vp, i32, i32, f32, f32]
    lib.b2h_joint_set_limits.argtypes = [vp, i32, i32, f32, f32]
    lib.b2h_joint_count.argtypes = [vp]
    lib.b2h_joint_set_target.argtypes = [vp, i32, f32, f32]
    lib.b2h_joint_set_spring.argtypes = [vp, i32, f32, f32, f32]
    lib.b2h_joint_order.argtypes = [vp, i32, vp]
    lib.b2h_profile.argtypes = [vp, vp]
    lib.b2h_step_info.argtypes = [vp, vp]
    lib.b2h_host_timings.argtypes = [vp, vp]
    lib.b2h_device_handle.argtypes = [vp]
    lib.b2h_device_handle.restype = ctypes.c_void_p
    lib.b2h_sum_y.argtypes = [vp, ctypes.c_int32, ctypes.c_int32]
    lib.b2h_sum_y.restype = ctypes.c_double
    lib.b2h_hash.argtypes = [vp]
    lib.b2h_hash.restype = u32
    lib.b2h_set_transform.argtypes = [vp, i32, f32, f32, f32]
    lib.b2h_set_velocity.argtypes = [vp, i32, f32, f32, f32]
    lib.b2h_set_type.argtypes = [vp, i32, i32]
    lib.b2h_set_active.argtypes = [vp, i32, i32]
    lib.b2h_query_aabb.argtypes = [vp, vp, i32, vp]
    lib.b2h_query_aabb.restype = i32
    lib.b2h_ray_cast_closest.argtypes = [vp, vp, vp, vp]
    lib.b2h_ray_cast_closest.restype = i32
    lib.b2h_shift_origin.argtypes = [vp, f32, f32]
    lib.b2h_record_post_solve.argtypes = [vp, i32]
    lib.b2h_set_pre_solve_rule.argtypes = [vp, i32]
    lib.b2h_pre_solve_digest.argtypes = [vp, vp, vp]
    lib.b2h_post_solve_digest.argtypes = [vp, vp, vp]
    lib.b2h_set_modulo_filter.argtypes = [vp, i32]
    lib.b2h_set_filter.argtypes = [vp, i32, ctypes.c_uint16, ctypes.c_uint16, ctypes.c_int16]
    lib.b2h_apply_force.argtypes = [vp, i32, f32, f32, f32]
    lib.b2h_set_awake.argtypes = [vp, i32, i32]
    lib.b2h_destroy_body.argtypes = [vp, i32]
    lib.b2h_set_body_param.argtypes = [vp, i32, i32, f32]
    lib.b2h_destroy_fixture.argtypes = [vp, i32]
    lib.b2h_plan_strip.argtypes = [vp, i32, i32, f32, i32, vp, vp, vp, vp, vp]
    lib.b2h_sharded_create.argtypes = [vp, i32, f32, vp, f32]
    lib.b2h_sharded_create.restype = vp
    lib.b2h_sharded_status.argtypes = [vp]
    lib.b2h_sharded_error.argtypes = [vp]
    lib.b2h_sharded_error.restype = ctypes.c_char_p
    lib.b2h_sharded_step.argtypes = [vp, f32, i32, i32]
    lib.b2h_sharded_gather.argtypes = [vp, vp]
    lib.b2h_sharded_strip_transforms.argtypes = [vp, i32, i32, vp]
    lib.b2h_sharded_destroy.argtypes = [vp]
    lib.b2h_sharded_rebalance.argtypes = [vp, vp]
    lib.b2h_sharded_lost_contacts.argtypes = [vp]
    lib.b2h_sharded_set_transport.argtypes = [vp, i32, i32]
    lib.b2h_sharded_set_rebalance_interval.argtypes = [vp, i32]
    lib.b2h_sharded_bounds.argtypes = [vp, vp]
    lib.b2h_sharded_strip_plan.argtypes = [vp, i32, vp, vp, vp, vp]
    lib.b2h_sharded_strip_bodies.argtypes = [vp, i32, vp]
    lib.b2h_sharded_solver_order.argtypes = [vp, i32, i32, vp, vp]
    lib.b2h_sharded_contact_keys.argtypes = [vp, i32, i32, vp]
    _lib = lib
    return lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class HostWorld:
    """A b2World built from a Scene with CreateBody / CreateFixture and stepped by a b2CudaStepExecutor."""

    def __init__(self, scene=None, device=0, download_bodies=True, events=True, gravity=None, world_flags=None,
                 arrays=None):
        self.lib = load()
        if scene is not None:
            gravity, world_flags, arrays = scene.gravity, scene.world_flags, scene.arrays()
        self.h = self.lib.b2h_create(gravity[0], gravity[1], world_flags, device, int(download_bodies), int(events))
        b, s, f = arrays
        b = np.ascontiguousarray(b, BODY_DEF)
        s = np.ascontiguousarray(s, SHAPE_DEF)
        f = np.ascontiguousarray(f, FIXTURE_DEF)
        rc = self.lib.b2h_build(self.h, len(b), _ptr(b), len(s), _ptr(s), len(f), _ptr(f))
        if rc != 0:
            raise RuntimeError("b2h_build failed: %d" % rc)
        if scene is not None and getattr(scene, "joints", None):
            self.create_joints(scene.joint_array())

    def add(self, scene):
        """CreateBody / CreateFixture for every body of another Scene (appended; between steps)"""
        b, s, f = scene.arrays()
        rc = self.lib.b2h_build(self.h, len(b), _ptr(np.ascontiguousarray(b, BODY_DEF)), len(s),
                                _ptr(np.ascontiguousarray(s, SHAPE_DEF)), len(f), _ptr(np.ascontiguousarray(f, FIXTURE_DEF)))
        if rc != 0:
            raise RuntimeError("b2h_build failed: %d" % rc)

    # ---- joints (b2World::CreateJoint / b2RevoluteJoint) ----
    def create_joints(self, joints):
        j = np.ascontiguousarray(joints, T.JOINT)
        rc = self.lib.b2h_create_joints(self.h, len(j), _ptr(j))
        if rc != 0:
            raise RuntimeError("b2World::CreateJoint failed: %d" % rc)

    def destroy_joint(self, joint):
        self.lib.b2h_destroy_joint(self.h, joint)

    def joint_count(self):
        return self.lib.b2h_joint_count(self.h)

    def joint_order(self):
        out = np.zeros(self.joint_count(), np.int32)
        n = self.lib.b2h_joint_order(self.h, len(out), _ptr(out))
        return out[:n]

    def joint_readings(self, inv_dt=60.0):
        """per joint: reaction force x, y, reaction torque, motor torque, joint angle, joint speed"""
        out = np.zeros((self.joint_count(), 6), np.float32)
        if len(out):
            self.lib.b2h_joint_readings(self.h, ctypes.c_float(inv_dt), _ptr(out))
        return out

    def joint_set_motor(self, joint, enable, speed, max_torque):
        self.lib.b2h_joint_set_motor(self.h, joint, int(enable), ctypes.c_float(speed), ctypes.c_float(max_torque))

    def joint_set_spring(self, joint, length, frequency_hz, damping_ratio):
        self.lib.b2h_joint_set_spring(self.h, joint, ctypes.c_float(length), ctypes.c_float(frequency_hz),
                                      ctypes.c_float(damping_ratio))

    def joint_set_target(self, joint, x, y):
        self.lib.b2h_joint_set_target(self.h, joint, ctypes.c_float(x), ctypes.c_float(y))

    def joint_set_limits(self, joint, enable, lower, upper):
        self.lib.b2h_joint_set_limits(self.h, joint, int(enable), ctypes.c_float(lower), ctypes.c_float(upper))

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.b2h_destroy(self.h)
            self.h = None

    # ---- b2CudaShardedWorld (Box2D/MT/b2CudaShardedWorld.h) ----
    def plan_strip(self, shard_count, rank, margin=2.0):
        """The C++ planner alone (no device): (scene body ids, ghost local ids, export local ids, fixtures, bounds) of a strip."""
        n = self.counts()[0]
        ids = np.zeros(n, np.int32)
        ghosts = np.zeros(n, np.int32)
        exports = np.zeros(n, np.int32)
        counts = np.zeros(4, np.int32)
        bounds = np.zeros(shard_count + 1, np.float64)
        rc = self.lib.b2h_plan_strip(self.h, shard_count, rank, ctypes.c_float(margin), n, _ptr(ids), _ptr(ghosts),
                                     _ptr(exports), _ptr(counts), _ptr(bounds))
        if rc != 0:
            raise RuntimeError("b2h_plan_strip failed: %d" % rc)
        return ids[:counts[0]], ghosts[:counts[1]], exports[:counts[2]], int(counts[3]), bounds

    def shard(self, shard_count, margin=2.0, devices=None, grid_fraction=1.0):
        """b2CudaShardedWorld over this (unstepped) world"""
        return ShardedWorld(self, shard_count, margin, devices, grid_fraction)

    def counts(self):
        a, b, c = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
        self.lib.b2h_counts(self.h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
        return a.value, b.value, c.value

    def step(self, dt=1.0 / 60.0, vel_iters=8, pos_iters=3):
        rc = self.lib.b2h_step(self.h, dt, vel_iters, pos_iters)
        if rc != 0:
            raise RuntimeError("b2World::Step failed (%d): %s" % (rc, self.lib.b2h_last_error(self.h).decode()))

    def bodies(self):
        out = np.zeros(self.counts()[0], T.BODY)
        self.lib.b2h_get_bodies(self.h, _ptr(out))
        return out

    def proxies(self):
        out = np.zeros(self.counts()[1], T.PROXY)
        self.lib.b2h_get_proxies(self.h, _ptr(out))
        return out

    def transforms(self):
        n = self.counts()[0]
        xya = np.zeros((n, 3), np.float32)
        awake = np.zeros(n, np.int32)
        self.lib.b2h_get_transforms(self.h, _ptr(xya), _ptr(awake))
        return xya, awake

    def mass(self):
        out = np.zeros((self.counts()[0], 4), np.float32)
        self.lib.b2h_get_mass(self.h, _ptr(out))
        return out

    def contacts(self):
        n = self.counts()[2]
        keys = np.zeros(n, np.uint64)
        touching = np.zeros(n, np.int32)
        points = np.zeros(n, np.int32)
        m = self.lib.b2h_get_contacts(self.h, n, _ptr(keys), _ptr(touching), _ptr(points))
        return keys[:m], touching[:m], points[:m]

    def events(self, kind):
        cap = 1 << 16
        while True:
            out = np.zeros(cap, np.uint64)
            n = self.lib.b2h_events(self.h, kind, cap, _ptr(out))
            if n <= cap:
                return out[:n]
            cap = n

    def solver_order(self):
        cap = max(16, 4 * self.counts()[2])
        out = np.zeros(cap, np.uint64)
        n = self.lib.b2h_solver_order(self.h, cap, _ptr(out))
        return out[:n]

    def profile(self):
        out = np.zeros(13, np.float32)
        self.lib.b2h_profile(self.h, _ptr(out))
        return out

    def step_info(self):
        out = np.zeros((), T.STEP_INFO)
        self.lib.b2h_step_info(self.h, _ptr(out))
        return out

    def host_timings(self):
        """ms of the last step on the calling thread: upload, b2cuStep, body download, events"""
        out = np.zeros(4, np.float32)
        self.lib.b2h_host_timings(self.h, _ptr(out))
        return out

    def device_world(self):
        """b2cuda.World view of the device copy (diagnostics: contacts, proxies, solver order)"""
        import b2cuda
        handle = self.lib.b2h_device_handle(self.h)
        assert handle, "the world has not been stepped yet"
        nb, np_, _ = self.counts()
        return b2cuda.World.attach(handle, nb, 0, np_)

    def sum_y(self, first, count):
        return float(self.lib.b2h_sum_y(self.h, first, count))

    def hash(self):
        return self.lib.b2h_hash(self.h)

    def set_transform(self, body, x, y, angle):
        self.lib.b2h_set_transform(self.h, body, x, y, angle)

    def set_modulo_filter(self, modulus):
        self.lib.b2h_set_modulo_filter(self.h, modulus)

    def set_pre_solve_rule(self, modulus):
        self.lib.b2h_set_pre_solve_rule(self.h, modulus)

    def pre_solve_digest(self):
        d = np.zeros(1, np.uint64)
        c = np.zeros(1, np.int64)
        self.lib.b2h_pre_solve_digest(self.h, _ptr(d), _ptr(c))
        return int(d[0]), int(c[0])

    def record_post_solve(self, on=True):
        self.lib.b2h_record_post_solve(self.h, int(on))

    def post_solve_digest(self):
        d = np.zeros(1, np.uint64)
        c = np.zeros(1, np.int64)
        self.lib.b2h_post_solve_digest(self.h, _ptr(d), _ptr(c))
        return int(d[0]), int(c[0])

    def query_aabb(self, box):
        a = np.asarray(box, np.float32)
        out = np.zeros(1 << 16, np.int32)
        n = self.lib.b2h_query_aabb(self.h, _ptr(a), len(out), _ptr(out))
        return out[:n]

    def ray_cast_closest(self, p1, p2):
        a, b = np.asarray(p1, np.float32), np.asarray(p2, np.float32)
        out = np.zeros(5, np.float32)
        proxy = self.lib.b2h_ray_cast_closest(self.h, _ptr(a), _ptr(b), _ptr(out))
        return proxy, out

    def shift_origin(self, x, y):
        self.lib.b2h_shift_origin(self.h, x, y)

    def set_active(self, body, on):
        self.lib.b2h_set_active(self.h, body, int(on))

    def set_type(self, body, body_type):
        self.lib.b2h_set_type(self.h, body, body_type)

    def set_filter(self, fixture, category, mask, group):
        self.lib.b2h_set_filter(self.h, fixture, category, mask, group)

    def set_velocity(self, body, vx, vy, w):
        self.lib.b2h_set_velocity(self.h, body, vx, vy, w)

    def apply_force(self, body, fx, fy, torque):
        self.lib.b2h_apply_force(self.h, body, fx, fy, torque)

    def set_awake(self, body, awake):
        self.lib.b2h_set_awake(self.h, body, int(awake))

    def set_options(self, download_bodies, events):
        self.lib.b2h_set_options(self.h, int(download_bodies), int(events))

    def apply_force_range(self, first, count, fx, fy):
        self.lib.b2h_apply_force_range(self.h, first, count, fx, fy)

    # ---- sharding (b2CudaStepExecutor::ConfigureShard / GetShardLink / ConnectShard) ----
    def shard_configure(self, rank, rank_count, ghost_bodies, export_bodies, grid_fraction=1.0):
        g = np.ascontiguousarray(ghost_bodies, np.int32)
        e = np.ascontiguousarray(export_bodies, np.int32)
        rc = self.lib.b2h_shard_configure(self.h, rank, rank_count, _ptr(g), len(g), _ptr(e), len(e), grid_fraction)
        if rc != 0:
            raise RuntimeError("ConfigureShard failed (%d): %s" % (rc, self.lib.b2h_last_error(self.h).decode()))

    def shard_link(self):
        link = np.zeros((), T.SHARD_LINK)
        rc = self.lib.b2h_shard_link(self.h, _ptr(link))
        if rc != 0:
            raise RuntimeError("GetShardLink failed (%d)" % rc)
        return link

    def shard_connect(self, lower=None, upper=None):
        lo = None if lower is None else np.array(lower, dtype=T.SHARD_LINK, copy=True)
        up = None if upper is None else np.array(upper, dtype=T.SHARD_LINK, copy=True)
        rc = self.lib.b2h_shard_connect(self.h, None if lo is None else _ptr(lo), None if up is None else _ptr(up))
        if rc != 0:
            raise RuntimeError("ConnectShard failed (%d): %s" % (rc, self.lib.b2h_last_error(self.h).decode()))

    def destroy_body(self, body):
        self.lib.b2h_destroy_body(self.h, body)

    def destroy_fixture(self, fixture):
        self.lib.b2h_destroy_fixture(self.h, fixture)

    LINEAR_DAMPING, ANGULAR_DAMPING, GRAVITY_SCALE, BULLET, SLEEPING_ALLOWED = range(5)

    def set_body_param(self, body, which, value):
        self.lib.b2h_set_body_param(self.h, body, which, float(value))


class ShardedWorld:
    """b2CudaShardedWorld: one scene cut into x-strips, one strip per GPU, stepped concurrently."""

    def __init__(self, host, shard_count, margin=2.0, devices=None, grid_fraction=1.0):
        self.lib = load()
        self.host = host
        self.count = shard_count
        dev = None if devices is None else np.ascontiguousarray(devices, np.int32)
        self.s = self.lib.b2h_sharded_create(host.h, shard_count, ctypes.c_float(margin),
                                             None if dev is None else _ptr(dev), ctypes.c_float(grid_fraction))
        if self.lib.b2h_sharded_status(self.s) != 0:
            raise RuntimeError("b2CudaShardedWorld: %s" % self.lib.b2h_sharded_error(self.s).decode())

    def step(self, dt=1.0 / 60.0, vel_iters=8, pos_iters=3):
        rc = self.lib.b2h_sharded_step(self.s, ctypes.c_float(dt), vel_iters, pos_iters)
        if rc != 0:
            raise RuntimeError("b2CudaShardedWorld::Step: %d %s" % (rc, self.lib.b2h_sharded_error(self.s).decode()))

    def gather(self):
        """copy the stepped state back into the host world this was made from"""
        self.lib.b2h_sharded_gather(self.s, self.host.h)

    def rebalance(self, bounds=None):
        """b2CudaShardedWorld::Rebalance: new strips at the current positions (or at the given boundaries), state carried over"""
        b = None if bounds is None else np.ascontiguousarray(bounds, np.float64)
        rc = self.lib.b2h_sharded_rebalance(self.s, None if b is None else _ptr(b))
        if rc != 0:
            raise RuntimeError("b2CudaShardedWorld::Rebalance: %d %s" % (rc, self.lib.b2h_sharded_error(self.s).decode()))

    def set_transport(self, download_bodies, events):
        self.lib.b2h_sharded_set_transport(self.s, int(download_bodies), int(events))

    def set_rebalance_interval(self, steps):
        self.lib.b2h_sharded_set_rebalance_interval(self.s, steps)

    def lost_contacts(self):
        return self.lib.b2h_sharded_lost_contacts(self.s)

    def bounds(self):
        out = np.zeros(self.count + 1, np.float64)
        self.lib.b2h_sharded_bounds(self.s, _ptr(out))
        return out

    def strip_plan(self, rank):
        """(scene body ids, ghost local ids, export local ids, proxy count) of a strip as it is now"""
        counts = np.zeros(4, np.int32)
        self.lib.b2h_sharded_strip_plan(self.s, rank, _ptr(counts), None, None, None)
        ids = np.zeros(counts[0], np.int32)
        ghosts = np.zeros(counts[1], np.int32)
        exports = np.zeros(counts[2], np.int32)
        self.lib.b2h_sharded_strip_plan(self.s, rank, _ptr(counts), _ptr(ids), _ptr(ghosts), _ptr(exports))
        return ids, ghosts, exports, int(counts[3])

    def strip_bodies(self, rank):
        out = np.zeros(len(self.strip_plan(rank)[0]), T.BODY)
        self.lib.b2h_sharded_strip_bodies(self.s, rank, _ptr(out))
        return out

    def strip_solver_order(self, rank, capacity=1 << 20):
        """(keys over scene proxy ids, colours) of the strip's constraints of the last step, in solve order"""
        keys = np.zeros(capacity, np.uint64)
        colour = np.zeros(capacity, np.int32)
        n = self.lib.b2h_sharded_solver_order(self.s, rank, capacity, _ptr(keys), _ptr(colour))
        if n < 0 or n > capacity:
            raise RuntimeError("b2h_sharded_solver_order: %d" % n)
        return keys[:n], colour[:n]

    def strip_contact_keys(self, rank, capacity=1 << 20):
        keys = np.zeros(capacity, np.uint64)
        n = self.lib.b2h_sharded_contact_keys(self.s, rank, capacity, _ptr(keys))
        if n < 0 or n > capacity:
            raise RuntimeError("b2h_sharded_contact_keys: %d" % n)
        return keys[:n]

    def strip_transforms(self, rank):
        n = self.lib.b2h_sharded_strip_transforms(self.s, rank, 0, None)
        out = np.zeros((n, 3), np.float32)
        self.lib.b2h_sharded_strip_transforms(self.s, rank, n, _ptr(out))
        return out

    def __del__(self):
        if getattr(self, "s", None):
            self.lib.b2h_sharded_destroy(self.s)
            self.s = None
