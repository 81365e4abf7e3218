"""Parity harness: reference world (oracle) vs device world, modelled on the reference's own A/B consistency
check (Testbed/Framework/TestMT.cpp:50-132: position, angle and awake state of every body must be equal after
every step), but comparing reference vs GPU instead of A vs B.

The coloured Gauss-Seidel legally reorders the reference's sequential impulses, so the oracle is stepped with
`step_ordered`: same phases, islands and b2Island::Solve code, contacts of every island sorted into the GPU's
solver order.
"""
import numpy as np

import b2cuda_types as T

# body flags that the device maintains exactly like the reference
BODY_FLAG_MASK = T.BODY_TYPE_MASK | T.BODY_AWAKE | T.BODY_AUTOSLEEP | T.BODY_BULLET | T.BODY_FIXED_ROTATION | T.BODY_ACTIVE
# contact flags compared (island / toi bookkeeping flags are transient or host-side)
CONTACT_FLAG_MASK = T.CONTACT_TOUCHING | T.CONTACT_ENABLED | T.CONTACT_TOI_CANDIDATE | T.CONTACT_INACTIVE

BODY_FLOAT_FIELDS = ["px", "py", "qs", "qc", "cx", "cy", "a", "c0x", "c0y", "a0", "vx", "vy", "w", "fx", "fy", "torque",
                     "sleepTime"]


def dedupe_shapes(shapes, proxies):
    """Collapse identical geometry records (most scenes share one box) and remap proxies.shape."""
    if len(shapes) == 0:
        return shapes, proxies.copy()
    raw = shapes.view(np.uint8).reshape(len(shapes), -1)
    _, first, inverse = np.unique(raw, axis=0, return_index=True, return_inverse=True)
    order = np.argsort(first)
    rank = np.empty_like(order)
    rank[order] = np.arange(len(order))
    table = shapes[first[order]]
    p = proxies.copy()
    p["shape"] = rank[inverse.reshape(-1)][proxies["shape"]]
    return table, p


def ref_state(ref):
    """(bodies, shape table, proxies, contacts) of a RefWorld in the C-ABI record formats."""
    bodies = ref.bodies()
    shapes, proxies = dedupe_shapes(ref.shapes(), ref.proxies())
    contacts = ref.contacts()
    return bodies, shapes, proxies, contacts


def gpu_world_from_ref(b2cuda, ref, **caps):
    bodies, shapes, proxies, contacts = ref_state(ref)
    w = b2cuda.World(gravity=ref.gravity, flags=ref.world_flags,
                     body_capacity=len(bodies), proxy_capacity=len(proxies), shape_capacity=len(shapes),
                     contact_capacity=max(1024, 16 * len(proxies)), **caps)
    w.load_state(bodies, shapes, proxies, contacts, inv_dt0=ref.inv_dt0())
    if ref.joint_count:
        w.set_joints(ref.joints())
    return w


def force_state(gpu, ref):
    """Teacher forcing reference -> GPU: overwrite the device state with the oracle's."""
    bodies = ref.bodies()
    proxies = gpu.get_proxies()
    rp = ref.proxies()
    for f in ("aabb", "fat"):
        proxies[f] = rp[f]
    proxies["flags"] = rp["flags"]
    gpu.set_bodies(bodies)
    gpu.set_proxies(proxies)
    gpu.set_contacts(ref.contacts())
    gpu.set_inv_dt0(ref.inv_dt0())


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def assert_floats_equal(name, got, want, tol=0.0):
    got = np.asarray(got, np.float32)
    want = np.asarray(want, np.float32)
    if tol == 0.0:
        same = (_bits(got) == _bits(want)) | ((got == 0) & (want == 0))
        if not same.all():
            idx = np.argwhere(~same)[0]
            raise AssertionError("%s: %d of %d values differ bitwise; first at %s: got %r want %r" %
                                 (name, (~same).sum(), same.size, tuple(idx), got[tuple(idx)], want[tuple(idx)]))
    else:
        scale = np.maximum(1.0, np.abs(want))
        err = np.abs(got - want) / scale
        if not (err <= tol).all():
            idx = np.unravel_index(np.argmax(err), err.shape)
            raise AssertionError("%s: max relative error %g > %g at %s: got %r want %r" %
                                 (name, err.max(), tol, idx, got[idx], want[idx]))


def compare_bodies(gb, rb, tol=0.0):
    assert len(gb) == len(rb)
    gf = gb["flags"] & BODY_FLAG_MASK
    rf = rb["flags"] & BODY_FLAG_MASK
    if not (gf == rf).all():
        i = int(np.argwhere(gf != rf)[0])
        raise AssertionError("body %d flags differ: got %#x want %#x" % (i, gb["flags"][i], rb["flags"][i]))
    for f in BODY_FLOAT_FIELDS:
        assert_floats_equal("body." + f, gb[f], rb[f], tol)


def compare_proxies(gp, rp):
    for f in ("aabb", "fat"):
        assert_floats_equal("proxy." + f, gp[f], rp[f])


def compare_contacts(gc, rc, tol=0.0, impulses=True):
    gk = T.contact_keys(gc)
    rk = T.contact_keys(rc)
    if len(gk) != len(rk) or not (gk == rk).all():
        missing = np.setdiff1d(rk, gk)
        extra = np.setdiff1d(gk, rk)
        raise AssertionError("contact key sets differ: %d missing %s, %d extra %s" %
                             (len(missing), [hex(int(k)) for k in missing[:5]], len(extra),
                              [hex(int(k)) for k in extra[:5]]))
    assert (gc["proxyA"] == rc["proxyA"]).all() and (gc["proxyB"] == rc["proxyB"]).all(), "fixture A/B order differs"
    gf = gc["flags"] & CONTACT_FLAG_MASK
    rf = rc["flags"] & CONTACT_FLAG_MASK
    if not (gf == rf).all():
        i = int(np.argwhere(gf != rf)[0])
        raise AssertionError("contact %#x flags differ: got %#x want %#x" % (int(gk[i]), gc["flags"][i], rc["flags"][i]))
    gm, rm = gc["manifold"], rc["manifold"]
    assert (gm["pointCount"] == rm["pointCount"]).all(), "manifold point counts differ"
    t = rm["pointCount"] > 0
    assert (gm["type"][t] == rm["type"][t]).all(), "manifold types differ"
    for k in range(2):
        has = rm["pointCount"] > k
        assert (gm["id"][has, k] == rm["id"][has, k]).all(), "feature ids differ"
        assert_floats_equal("manifold.point%d.localPoint" % k, gm["points"]["localPoint"][has, k],
                            rm["points"]["localPoint"][has, k], tol)
        if impulses:
            assert_floats_equal("manifold.point%d.normalImpulse" % k, gm["points"]["normalImpulse"][has, k],
                                rm["points"]["normalImpulse"][has, k], tol)
            assert_floats_equal("manifold.point%d.tangentImpulse" % k, gm["points"]["tangentImpulse"][has, k],
                                rm["points"]["tangentImpulse"][has, k], tol)
    assert_floats_equal("manifold.localNormal", gm["localNormal"][t], rm["localNormal"][t], tol)
    assert_floats_equal("manifold.localPoint", gm["localPoint"][t], rm["localPoint"][t], tol)
    for f in ("friction", "restitution"):
        assert_floats_equal("contact." + f, gc[f], rc[f])


def compare_joints(gj, rj, tol=0.0):
    """persistent joint state: accumulated impulses, motor impulse, limit state"""
    assert len(gj) == len(rj), "joint count %d != %d" % (len(gj), len(rj))
    if not (gj["limitState"] == rj["limitState"]).all():
        i = int(np.nonzero(gj["limitState"] != rj["limitState"])[0][0])
        raise AssertionError("joint %d limit state: got %d want %d" % (i, gj["limitState"][i], rj["limitState"][i]))
    assert_floats_equal("joint impulse", gj["impulse"], rj["impulse"], tol)
    assert_floats_equal("joint motorImpulse", gj["motorImpulse"], rj["motorImpulse"], tol)
    assert_floats_equal("joint lastSolve", gj["lastSolve"], rj["lastSolve"], tol)


def compare_events(gpu, ref):
    for kind, name in ((T.EVENT_BEGIN, "begin"), (T.EVENT_END, "end")):
        g = gpu.events(kind)
        r = ref.events(kind)
        assert len(g) == len(r) and (g == r).all(), "%s events differ: got %d want %d" % (name, len(g), len(r))


def lockstep(gpu, ref, steps, dt=1.0 / 60.0, vel_iters=8, pos_iters=3, teacher=False, tol=0.0, check_every=1,
             on_step=None):
    """Step both worlds `steps` times and compare everything after every step.

    GPU steps first; the oracle then repeats the step with its island contacts in the GPU's solver order.
    teacher=True re-injects the oracle state into the device before every step (single-step parity)."""
    infos = []
    for s in range(steps):
        if teacher and s > 0:
            force_state(gpu, ref)
        if ref.joint_count:
            if teacher and s > 0:
                gpu.set_joints(ref.joints())
            ref.set_joint_order(gpu.joint_order())
        info = gpu.step(dt, vel_iters, pos_iters)
        keys, _ = gpu.solver_order()
        unranked = ref.step_ordered(keys, dt, vel_iters, pos_iters)
        assert unranked == 0, "step %d: %d oracle island contacts were not in the GPU solver set" % (s, unranked)
        if on_step:
            on_step(s, gpu, ref)
        if s % check_every == 0 or s == steps - 1:
            try:
                compare_contacts(gpu.get_contacts(), ref.contacts(), tol)
                compare_bodies(gpu.get_bodies(), ref.bodies(), tol)
                compare_proxies(gpu.get_proxies(), ref.proxies())
                compare_events(gpu, ref)
                if ref.joint_count:
                    compare_joints(gpu.get_joints(), ref.joints(), tol)
            except AssertionError as e:
                raise AssertionError("step %d: %s" % (s, e))
        infos.append(info)
    return infos
