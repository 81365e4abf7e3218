"""oracle/b2o_broadphase.py -- TEST INFRASTRUCTURE (CPU oracle), numpy restatement of the broad-phase RULES.

The reference finds pairs with a dynamic AABB tree (Box2D/Collision/b2DynamicTree.cpp, b2BroadPhase.h:211-267); the
pair SET it produces does not depend on the tree shape, only on these rules, restated here by
brute force so that both the compiled reference and the device grid can be checked against an independent
statement of the semantics:

  fat AABB rule      b2DynamicTree::MoveProxy (:130-174): a proxy whose tight AABB leaves its fat AABB gets
                     fat = tight +- b2_aabbExtension (0.1), extended by b2_aabbMultiplier (2) x displacement on the
                     leading side
  overlap            b2TestOverlap(b2AABB, b2AABB) (Box2D/Collision/b2Collision.h:273-286): closed intervals
  pair filter        b2ContactManager::AddPair (Box2D/Dynamics/b2ContactManager.cpp:237-312): different bodies, at
                     least one dynamic body, default b2ContactFilter (b2WorldCallbacks.cpp:24-38), a contact class
                     exists (no edge-edge)
  contact set        after a step = contacts that survived Collide (active ones whose fat AABBs, as they were
                     BEFORE the step's proxy update, still overlapped; inactive ones unconditionally)
                     UNION new pairs of moved proxies (fat AABBs after the update)
"""
import numpy as np

AABB_EXTENSION = np.float32(0.1)
AABB_MULTIPLIER = np.float32(2.0)


def move_proxy(fat, tight, displacement):
    """New fat AABB (float32[4]) or None if the tight AABB is still contained (no move)."""
    fat = np.asarray(fat, np.float32)
    tight = np.asarray(tight, np.float32)
    if fat[0] <= tight[0] and fat[1] <= tight[1] and tight[2] <= fat[2] and tight[3] <= fat[3]:
        return None
    b = np.array([tight[0] - AABB_EXTENSION, tight[1] - AABB_EXTENSION, tight[2] + AABB_EXTENSION,
                  tight[3] + AABB_EXTENSION], np.float32)
    d = (AABB_MULTIPLIER * np.asarray(displacement, np.float32)).astype(np.float32)
    if d[0] < 0:
        b[0] = np.float32(b[0] + d[0])
    else:
        b[2] = np.float32(b[2] + d[0])
    if d[1] < 0:
        b[1] = np.float32(b[1] + d[1])
    else:
        b[3] = np.float32(b[3] + d[1])
    return b


def overlap_matrix(fat_a, fat_b):
    """b2TestOverlap for every (a, b): boolean [len(a), len(b)]."""
    a = np.asarray(fat_a, np.float32)[:, None, :]
    b = np.asarray(fat_b, np.float32)[None, :, :]
    d1x = b[..., 0] - a[..., 2]
    d1y = b[..., 1] - a[..., 3]
    d2x = a[..., 0] - b[..., 2]
    d2y = a[..., 1] - b[..., 3]
    return ~((d1x > 0) | (d1y > 0) | (d2x > 0) | (d2y > 0))


def should_collide(proxies, body_types, shape_types, i, j):
    """Vectorised pair filter for index arrays i < j."""
    bi, bj = proxies["body"][i], proxies["body"][j]
    ok = bi != bj
    ok &= (body_types[bi] == 2) | (body_types[bj] == 2)
    gi, gj = proxies["groupIndex"][i], proxies["groupIndex"][j]
    same_group = (gi == gj) & (gi != 0)
    mask_ok = ((proxies["maskBits"][i] & proxies["categoryBits"][j]) != 0) & \
              ((proxies["categoryBits"][i] & proxies["maskBits"][j]) != 0)
    ok &= np.where(same_group, gi > 0, mask_ok)
    si, sj = shape_types[proxies["shape"][i]], shape_types[proxies["shape"][j]]
    ok &= ~((si == 1) & (sj == 1))
    return ok


def candidate_pairs(proxies, body_types, shape_types, moved):
    """Keys (i << 32 | j, i < j) of every filter-passing pair with overlapping fat AABBs and a moved member."""
    fat = proxies["fat"]
    m = np.where(moved)[0]
    if len(m) == 0:
        return np.zeros(0, np.uint64)
    hit = overlap_matrix(fat[m], fat)
    qi, pj = np.nonzero(hit)
    a = m[qi]
    i = np.minimum(a, pj)
    j = np.maximum(a, pj)
    keep = i != j
    i, j = i[keep], j[keep]
    ok = should_collide(proxies, body_types, shape_types, i, j)
    keys = (i[ok].astype(np.uint64) << np.uint64(32)) | j[ok].astype(np.uint64)
    return np.unique(keys)


def expected_contact_keys(old_keys, old_inactive, proxies_before, proxies_after, body_types, shape_types, moved):
    """The contact key set after a step, from the set before it (see the module docstring)."""
    old_keys = np.asarray(old_keys, np.uint64)
    i = (old_keys >> np.uint64(32)).astype(np.int64)
    j = (old_keys & np.uint64(0xFFFFFFFF)).astype(np.int64)
    fb = proxies_before["fat"]
    fa, fbj = fb[i], fb[j]
    still = ~((fbj[:, 0] - fa[:, 2] > 0) | (fbj[:, 1] - fa[:, 3] > 0) | (fa[:, 0] - fbj[:, 2] > 0) |
              (fa[:, 1] - fbj[:, 3] > 0))
    survivors = old_keys[still | np.asarray(old_inactive, bool)]
    new = candidate_pairs(proxies_after, body_types, shape_types, moved)
    return np.union1d(survivors, new)
