"""ctypes binding of libb2cuda.so (include/b2cuda.h).  Thin: numpy host buffers in, status codes checked.

There is no CPU fallback: if the library is missing it is built with nvcc, and if there is no CUDA device every
compute call raises.
"""
import ctypes
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.dirname(_HERE)
sys.path.insert(0, _HERE)
import b2cuda_types as T  # noqa: E402

_lib = None

API = [
    "b2cuGetDeviceCount", "b2cuVersion", "b2cuCreateWorld", "b2cuDestroyWorld", "b2cuGetLastError",
    "b2cuSetWorldParams", "b2cuSetInvDt0", "b2cuSetCounts", "b2cuSetBodies", "b2cuGetBodies", "b2cuSetShapes",
    "b2cuSetProxies", "b2cuGetProxies", "b2cuSetContacts", "b2cuGetContactCount", "b2cuGetContacts", "b2cuStep",
    "b2cuGetContactsByKey", "b2cuGetEvents", "b2cuGetSolverOrder", "b2cuGetIslandLabels", "b2cuGetToiCandidates", "b2cuCollidePairs",
    "b2cuSinCos", "b2cuShardConfigure", "b2cuShardGetLink", "b2cuShardConnect", "b2cuHostAlloc", "b2cuHostFree", "b2cuSetBodyMirror", "b2cuSetPairFilter", "b2cuDistancePairs", "b2cuTimeOfImpactPairs", "b2cuQueryAABB", "b2cuRayCastCandidates", "b2cuSetPreSolveHook", "b2cuGetPreSolveContacts", "b2cuDisableContacts", "b2cuGetBodyStates", "b2cuGetEventContacts", "b2cuGetToiEvents", "b2cuSetBodyForces", "b2cuSetEventPrefetch", "b2cuGetBodySweepStarts", "b2cuSetJoints", "b2cuGetJointCount", "b2cuGetJoints", "b2cuGetJointOrder",
]


class B2cuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("b2cuda error %d: %s" % (code, msg))
        self.code = code


def lib_path():
    return os.path.join(_PKG, "libb2cuda.so")


def load():
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        sys.path.insert(0, _PKG)
        import build
        build.build()
    lib = ctypes.CDLL(path)
    vp, i32, f32, u32 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_float, ctypes.c_uint32
    lib.b2cuGetDeviceCount.restype = i32
    lib.b2cuVersion.restype = ctypes.c_char_p
    lib.b2cuCreateWorld.argtypes = [vp, ctypes.POINTER(vp)]
    lib.b2cuDestroyWorld.argtypes = [vp]
    lib.b2cuDestroyWorld.restype = None
    lib.b2cuGetLastError.argtypes = [vp]
    lib.b2cuGetLastError.restype = ctypes.c_char_p
    lib.b2cuSetWorldParams.argtypes = [vp, vp, u32]
    lib.b2cuSetInvDt0.argtypes = [vp, f32]
    lib.b2cuSetCounts.argtypes = [vp, i32, i32, i32]
    for name in ("b2cuSetBodies", "b2cuGetBodies", "b2cuSetShapes", "b2cuSetProxies", "b2cuGetProxies",
                 "b2cuGetIslandLabels"):
        getattr(lib, name).argtypes = [vp, i32, i32, vp]
    lib.b2cuSetContacts.argtypes = [vp, i32, vp]
    lib.b2cuGetContactCount.argtypes = [vp, vp]
    lib.b2cuGetContacts.argtypes = [vp, i32, vp, vp]
    lib.b2cuStep.argtypes = [vp, f32, i32, i32, vp]
    lib.b2cuGetEvents.argtypes = [vp, i32, i32, vp, vp]
    lib.b2cuGetContactsByKey.argtypes = [vp, i32, vp, vp]
    lib.b2cuGetSolverOrder.argtypes = [vp, i32, vp, vp, vp]
    lib.b2cuGetToiCandidates.argtypes = [vp, i32, vp, vp]
    lib.b2cuGetToiEvents.argtypes = [vp, i32, vp, vp, vp, vp]
    lib.b2cuCollidePairs.argtypes = [i32, i32, vp, i32, vp, vp, vp, vp, vp]
    lib.b2cuSinCos.argtypes = [i32, i32, vp, vp, vp]
    lib.b2cuSetJoints.argtypes = [vp, i32, vp]
    lib.b2cuGetJointCount.argtypes = [vp, vp]
    lib.b2cuGetJoints.argtypes = [vp, i32, i32, vp]
    lib.b2cuGetJointOrder.argtypes = [vp, i32, vp, vp]
    lib.b2cuShardConfigure.argtypes = [vp, i32, i32, i32, vp, i32, vp, f32]
    lib.b2cuShardGetLink.argtypes = [vp, vp]
    lib.b2cuShardConnect.argtypes = [vp, vp, vp]
    _lib = lib
    return lib


def device_count():
    return load().b2cuGetDeviceCount()


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class World:
    """Device-resident world behind the C ABI."""

    def __init__(self, gravity=(0.0, -10.0), flags=T.WORLD_DEFAULT, device=0, body_capacity=0, proxy_capacity=0,
                 shape_capacity=0, contact_capacity=0):
        self.lib = load()
        d = np.zeros((), T.WORLD_DEF)
        d["device"] = device
        d["gravity"] = gravity
        d["flags"] = flags
        d["bodyCapacity"] = body_capacity
        d["proxyCapacity"] = proxy_capacity
        d["shapeCapacity"] = shape_capacity
        d["contactCapacity"] = contact_capacity
        h = ctypes.c_void_p()
        rc = self.lib.b2cuCreateWorld(_ptr(d), ctypes.byref(h))
        if rc != 0:
            raise B2cuError(rc, "b2cuCreateWorld failed (no CUDA device? there is no CPU fallback)")
        self.h = h
        self.body_count = self.shape_count = self.proxy_count = 0

    @classmethod
    def attach(cls, handle, body_count, shape_count, proxy_count):
        """Non-owning view of a b2cuWorld created elsewhere (b2CudaStepExecutor::GetDeviceWorld)."""
        w = cls.__new__(cls)
        w.lib = load()
        w.h = ctypes.c_void_p(handle)
        w.owned = False
        w.body_count, w.shape_count, w.proxy_count = body_count, shape_count, proxy_count
        return w

    def close(self):
        if getattr(self, "h", None):
            if getattr(self, "owned", True):
                self.lib.b2cuDestroyWorld(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def _check(self, rc):
        if rc != 0:
            raise B2cuError(rc, self.lib.b2cuGetLastError(self.h).decode())

    # ---- state upload / download ----
    def set_params(self, gravity, flags):
        g = np.asarray(gravity, np.float32)
        self._check(self.lib.b2cuSetWorldParams(self.h, _ptr(g), flags))

    def set_inv_dt0(self, v):
        self._check(self.lib.b2cuSetInvDt0(self.h, v))

    def set_counts(self, bodies, shapes, proxies):
        self._check(self.lib.b2cuSetCounts(self.h, bodies, shapes, proxies))
        self.body_count, self.shape_count, self.proxy_count = bodies, shapes, proxies

    def set_bodies(self, bodies, first=0):
        b = np.ascontiguousarray(bodies, T.BODY)
        self._check(self.lib.b2cuSetBodies(self.h, first, len(b), _ptr(b)))

    def get_bodies(self, first=0, count=None):
        count = self.body_count - first if count is None else count
        out = np.zeros(count, T.BODY)
        self._check(self.lib.b2cuGetBodies(self.h, first, count, _ptr(out)))
        return out

    def set_shapes(self, shapes, first=0):
        s = np.ascontiguousarray(shapes, T.SHAPE)
        self._check(self.lib.b2cuSetShapes(self.h, first, len(s), _ptr(s)))

    def set_proxies(self, proxies, first=0):
        p = np.ascontiguousarray(proxies, T.PROXY)
        self._check(self.lib.b2cuSetProxies(self.h, first, len(p), _ptr(p)))

    def get_proxies(self, first=0, count=None):
        count = self.proxy_count - first if count is None else count
        out = np.zeros(count, T.PROXY)
        self._check(self.lib.b2cuGetProxies(self.h, first, count, _ptr(out)))
        return out

    def set_contacts(self, contacts):
        c = np.ascontiguousarray(contacts, T.CONTACT)
        self._check(self.lib.b2cuSetContacts(self.h, len(c), _ptr(c)))

    def contact_count(self):
        n = ctypes.c_int32()
        self._check(self.lib.b2cuGetContactCount(self.h, ctypes.byref(n)))
        return n.value

    def get_contacts(self):
        n = self.contact_count()
        out = np.zeros(n, T.CONTACT)
        m = ctypes.c_int32()
        self._check(self.lib.b2cuGetContacts(self.h, n, _ptr(out), ctypes.byref(m)))
        return out

    def load_state(self, bodies, shapes, proxies, contacts=None, inv_dt0=None):
        """Upload a complete world state (records of include/b2cuda.h)."""
        self.set_counts(len(bodies), len(shapes), len(proxies))
        self.set_bodies(bodies)
        self.set_shapes(shapes)
        self.set_proxies(proxies)
        self.set_contacts(contacts if contacts is not None else np.zeros(0, T.CONTACT))
        if inv_dt0 is not None:
            self.set_inv_dt0(inv_dt0)

    # ---- stepping ----
    def set_pair_filter(self, fn):
        """fn(keys: uint64 array) -> bool array (True = the pair may collide), or None for the default rule"""
        proto = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64), ctypes.c_int32,
                                 ctypes.POINTER(ctypes.c_uint8))
        if fn is None:
            self._pair_filter = None
            self._check(self.lib.b2cuSetPairFilter(self.h, None, None))
            return

        def thunk(user, keys, count, keep):
            try:
                k = np.ctypeslib.as_array(keys, shape=(count,))
                out = np.asarray(fn(k.copy()), dtype=np.uint8)
                np.ctypeslib.as_array(keep, shape=(count,))[:] = out
                return 0
            except Exception:  # never let an exception cross the C boundary
                return 1

        self._pair_filter = proto(thunk)  # keep the callback object alive
        self._check(self.lib.b2cuSetPairFilter(self.h, self._pair_filter, None))

    def set_joints(self, joints):
        j = np.ascontiguousarray(joints, T.JOINT)
        self._check(self.lib.b2cuSetJoints(self.h, len(j), _ptr(j) if len(j) else None))

    def get_joints(self):
        n = ctypes.c_int32()
        self._check(self.lib.b2cuGetJointCount(self.h, ctypes.byref(n)))
        out = np.zeros(n.value, T.JOINT)
        if n.value:
            self._check(self.lib.b2cuGetJoints(self.h, 0, n.value, _ptr(out)))
        return out

    def joint_order(self):
        n = ctypes.c_int32()
        self._check(self.lib.b2cuGetJointCount(self.h, ctypes.byref(n)))
        out = np.zeros(n.value, np.int32)
        if n.value:
            self._check(self.lib.b2cuGetJointOrder(self.h, n.value, _ptr(out), ctypes.byref(n)))
        return out

    def step(self, dt=1.0 / 60.0, vel_iters=8, pos_iters=3):
        info = np.zeros((), T.STEP_INFO)
        self._check(self.lib.b2cuStep(self.h, dt, vel_iters, pos_iters, _ptr(info)))
        return info

    def _keys(self, fn, *args):
        n = ctypes.c_int32()
        self._check(fn(self.h, *args, 0, None, ctypes.byref(n)))
        out = np.zeros(n.value, np.uint64)
        if n.value:
            self._check(fn(self.h, *args, n.value, _ptr(out), ctypes.byref(n)))
        return out

    def toi_events(self, records=False):
        """(keys, kinds[, contact records]) of the begin / end events raised inside the time-of-impact sub-steps of the
        last step, in call order."""
        n = ctypes.c_int32()
        self._check(self.lib.b2cuGetToiEvents(self.h, 0, None, None, None, ctypes.byref(n)))
        keys = np.zeros(n.value, np.uint64)
        kinds = np.zeros(n.value, np.int32)
        recs = np.zeros(n.value, T.CONTACT)
        if n.value:
            self._check(self.lib.b2cuGetToiEvents(self.h, n.value, _ptr(keys), _ptr(kinds),
                                                  _ptr(recs) if records else None, ctypes.byref(n)))
        return (keys, kinds, recs) if records else (keys, kinds)

    def events(self, kind):
        """Callback order of the reference for one kind of event: the deferred calls of the discrete step in key order,
        then the calls made from inside the time-of-impact sub-steps as they happened."""
        discrete = self._keys(self.lib.b2cuGetEvents, kind)
        keys, kinds = self.toi_events()
        return np.concatenate([discrete, keys[kinds == kind]])

    def contacts_by_key(self, keys):
        k = np.ascontiguousarray(keys, np.uint64)
        out = np.zeros(len(k), T.CONTACT)
        self._check(self.lib.b2cuGetContactsByKey(self.h, len(k), _ptr(k), _ptr(out)))
        return out

    def toi_candidates(self):
        return self._keys(self.lib.b2cuGetToiCandidates)

    def solver_order(self):
        n = ctypes.c_int32()
        self._check(self.lib.b2cuGetSolverOrder(self.h, 0, None, None, ctypes.byref(n)))
        keys = np.zeros(n.value, np.uint64)
        colour = np.zeros(n.value, np.int32)
        if n.value:
            self._check(self.lib.b2cuGetSolverOrder(self.h, n.value, _ptr(keys), _ptr(colour), ctypes.byref(n)))
        return keys, colour

    # ---- sharding ----
    def shard_configure(self, rank, rank_count, ghost_bodies, export_bodies, grid_fraction=1.0):
        g = np.ascontiguousarray(ghost_bodies, np.int32)
        e = np.ascontiguousarray(export_bodies, np.int32)
        self._check(self.lib.b2cuShardConfigure(self.h, rank, rank_count, len(g), _ptr(g), len(e), _ptr(e),
                                                grid_fraction))

    def shard_link(self):
        link = np.zeros((), T.SHARD_LINK)
        self._check(self.lib.b2cuShardGetLink(self.h, _ptr(link)))
        return link

    def shard_connect(self, lower=None, upper=None):
        lo = None if lower is None else np.array(lower, dtype=T.SHARD_LINK, copy=True)
        up = None if upper is None else np.array(upper, dtype=T.SHARD_LINK, copy=True)
        self._check(self.lib.b2cuShardConnect(self.h, None if lo is None else _ptr(lo), None if up is None else _ptr(up)))

    def island_labels(self):
        out = np.zeros(self.body_count, np.int32)
        self._check(self.lib.b2cuGetIslandLabels(self.h, 0, self.body_count, _ptr(out)))
        return out


def collide_pairs(shapes, shape_a, xf_a, shape_b, xf_b, device=0):
    """Batched stand-alone narrow phase on the device; returns MANIFOLD records."""
    lib = load()
    s = np.ascontiguousarray(shapes, T.SHAPE)
    ia = np.ascontiguousarray(shape_a, np.int32)
    ib = np.ascontiguousarray(shape_b, np.int32)
    xa = np.ascontiguousarray(xf_a, np.float32).reshape(-1, 4)
    xb = np.ascontiguousarray(xf_b, np.float32).reshape(-1, 4)
    out = np.zeros(len(ia), T.MANIFOLD)
    rc = lib.b2cuCollidePairs(device, len(s), _ptr(s), len(ia), _ptr(ia), _ptr(xa), _ptr(ib), _ptr(xb), _ptr(out))
    if rc != 0:
        raise B2cuError(rc, "b2cuCollidePairs")
    return out


def distance_pairs(shapes, shape_a, xf_a, shape_b, xf_b, use_radii=True, device=0):
    """Batched b2Distance with a cold cache; returns a DISTANCE_RESULT array."""
    lib = load()
    shapes = np.ascontiguousarray(shapes, T.SHAPE)
    a = np.ascontiguousarray(shape_a, np.int32)
    b = np.ascontiguousarray(shape_b, np.int32)
    xa = np.ascontiguousarray(xf_a, np.float32)
    xb = np.ascontiguousarray(xf_b, np.float32)
    out = np.zeros(len(a), T.DISTANCE_RESULT)
    rc = lib.b2cuDistancePairs(device, len(shapes), _ptr(shapes), len(a), _ptr(a), _ptr(xa), _ptr(b), _ptr(xb),
                               1 if use_radii else 0, _ptr(out))
    if rc != 0:
        raise B2cuError(rc, "b2cuDistancePairs failed")
    return out


def time_of_impact_pairs(shapes, shape_a, sweep_a, shape_b, sweep_b, t_max, device=0):
    """Batched b2TimeOfImpact; returns a TOI_RESULT array."""
    lib = load()
    shapes = np.ascontiguousarray(shapes, T.SHAPE)
    a = np.ascontiguousarray(shape_a, np.int32)
    b = np.ascontiguousarray(shape_b, np.int32)
    wa = np.ascontiguousarray(sweep_a, T.SWEEP)
    wb = np.ascontiguousarray(sweep_b, T.SWEEP)
    tm = np.ascontiguousarray(np.broadcast_to(np.asarray(t_max, np.float32), (len(a),)))
    out = np.zeros(len(a), T.TOI_RESULT)
    rc = lib.b2cuTimeOfImpactPairs(device, len(shapes), _ptr(shapes), len(a), _ptr(a), _ptr(wa), _ptr(b), _ptr(wb),
                                   _ptr(tm), _ptr(out))
    if rc != 0:
        raise B2cuError(rc, "b2cuTimeOfImpactPairs failed")
    return out


def sincos(angles, device=0):
    lib = load()
    x = np.ascontiguousarray(angles, np.float32)
    s = np.zeros_like(x)
    c = np.zeros_like(x)
    rc = lib.b2cuSinCos(device, len(x), _ptr(x), _ptr(s), _ptr(c))
    if rc != 0:
        raise B2cuError(rc, "b2cuSinCos")
    return s, c
