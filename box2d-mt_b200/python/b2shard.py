"""Spatial sharding of a flat scene description into x-strips (include/b2cuda.h b2cuShard*).

Shard r holds every non-dynamic body, the dynamic bodies whose x lies in [bounds[r], bounds[r+1]) and GHOST copies
of the dynamic bodies of shard r+1 within `margin` of the common boundary.  Those same bodies are shard r+1's
EXPORT list, in the same order."""
import numpy as np

import b2cuda_types as T
from b2scene import BODY_DEF, FIXTURE_DEF


class ShardPlan:
    def __init__(self, rank, body_ids, fixture_ids, arrays, ghost_local, export_local):
        self.rank = rank
        self.body_ids = body_ids          # global body id of each local body
        self.fixture_ids = fixture_ids    # global fixture id of each local fixture (= proxy)
        self.arrays = arrays              # (bodies, shapes, fixtures) of this shard
        self.ghost_local = ghost_local    # local body ids that are ghosts (copies of shard rank+1's exports)
        self.export_local = export_local  # local body ids exported to shard rank-1


def strip_bounds(xs, dynamic, rank_count):
    """Equal-population strip boundaries over the dynamic bodies."""
    x = np.sort(xs[dynamic])
    cuts = [x[(len(x) * r) // rank_count] for r in range(1, rank_count)]
    return np.array([-np.inf] + cuts + [np.inf])


def split_scene(arrays, rank_count, margin=2.0, bounds=None, only_rank=None):
    bodies, shapes, fixtures = arrays
    dynamic = bodies["type"] == T.DYNAMIC_BODY
    xs = bodies["px"].astype(np.float64)
    if bounds is None:
        bounds = strip_bounds(xs, dynamic, rank_count)
    owner = np.full(len(bodies), -1)
    owner[dynamic] = np.clip(np.searchsorted(bounds, xs[dynamic], side="right") - 1, 0, rank_count - 1)
    fixture_body = fixtures["body"]
    plans = []
    for r in range(rank_count):
        if only_rank is not None and r != only_rank:
            plans.append(None)
            continue
        own = np.where(~dynamic | (owner == r))[0]
        if r + 1 < rank_count:
            ghosts = np.where(dynamic & (owner == r + 1) & (xs < bounds[r + 1] + margin))[0]
        else:
            ghosts = np.zeros(0, dtype=np.int64)
        exports = np.where(dynamic & (owner == r) & (xs < bounds[r] + margin))[0] if r > 0 else np.zeros(0, np.int64)
        # local ids keep the global relative order, so that the fixture A / fixture B roles of a contact (lower proxy id
        # first, b2ContactManager::AddPair) are the same in the shard and in the whole world
        body_ids = np.sort(np.concatenate([own, ghosts]))
        local_of = np.full(len(bodies), -1)
        local_of[body_ids] = np.arange(len(body_ids))
        b = bodies[body_ids].copy()
        keep = np.where(local_of[fixture_body] >= 0)[0]
        # fixtures in local body order (stable: creation order within a body is kept)
        order = np.argsort(local_of[fixture_body[keep]], kind="stable")
        fixture_ids = keep[order]
        f = fixtures[fixture_ids].copy()
        f["body"] = local_of[fixture_body[fixture_ids]]
        plans.append(ShardPlan(r, body_ids, fixture_ids, (b, shapes, f), local_of[ghosts].astype(np.int32),
                               local_of[exports].astype(np.int32)))
    return plans, bounds


def connect(worlds, plans, grid_fraction=1.0):
    """Configure and link b2cuda.World objects (same process) as the shards of one world."""
    n = len(worlds)
    for w, p in zip(worlds, plans):
        w.shard_configure(p.rank, n, p.ghost_local, p.export_local, grid_fraction)
    links = [w.shard_link() for w in worlds]
    for r, w in enumerate(worlds):
        w.shard_connect(links[r - 1] if r > 0 else None, links[r + 1] if r + 1 < n else None)


def global_keys(plan, local_keys):
    """Translate contact keys over local proxy ids into keys over global fixture ids."""
    k = np.asarray(local_keys, np.uint64)
    a = plan.fixture_ids[(k >> np.uint64(32)).astype(np.int64)].astype(np.uint64)
    b = plan.fixture_ids[(k & np.uint64(0xFFFFFFFF)).astype(np.int64)].astype(np.uint64)
    return (np.minimum(a, b) << np.uint64(32)) | np.maximum(a, b)


# ---- one process per shard (torchrun): the plumbing around the device calls ---------------------------------------
# torch.distributed is used for rendezvous-time exchange and for reducing timings only; the per-step halo traffic
# goes through the peer mailboxes inside the solver kernels (include/b2cuda.h, b2cuShardConnect).

def rank_plan(arrays, rank, world_size, margin):
    """The ShardPlan of this rank alone (every rank computes the same strip bounds from the same scene)."""
    plans, bounds = split_scene(arrays, world_size, margin=margin, only_rank=rank)
    return plans[rank], bounds


def exchange_links(dist, rank, world_size, link):
    """All-gather the b2cuShardLink records and return (lower neighbour's, upper neighbour's), None at the ends.
    Checks that the halo lists of neighbouring ranks have matching lengths before any device memory is touched."""
    blobs = [None] * world_size
    dist.all_gather_object(blobs, np.asarray(link).tobytes())
    links = [np.frombuffer(b, dtype=T.SHARD_LINK)[0] for b in blobs]
    for r, l in enumerate(links):
        if int(l["rank"]) != r or int(l["rankCount"]) != world_size:
            raise RuntimeError("shard link of rank %d says rank %d of %d" % (r, int(l["rank"]), int(l["rankCount"])))
    for r in range(world_size - 1):
        if int(links[r]["ghostCount"]) != int(links[r + 1]["exportCount"]):
            raise RuntimeError("rank %d holds %d ghosts but rank %d exports %d bodies" % (
                r, int(links[r]["ghostCount"]), r + 1, int(links[r + 1]["exportCount"])))
    lower = links[rank - 1] if rank > 0 else None
    upper = links[rank + 1] if rank + 1 < world_size else None
    return lower, upper


def reduce_scalar(dist, x, op, device=None):
    """max / sum of a python float over the ranks (the timings of bench.py: device time is the max over ranks)."""
    import torch
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
    return float(t.item())

