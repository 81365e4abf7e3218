#!/usr/bin/env python3
"""bench.py -- body-steps/s of b2World::Step on the BASELINE.json pile workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--bodies B] [--impl reference]

Product arm (default): the 1M-body mixed polygon/circle pile (BASELINE.json configs[4]) is built
through the reference-facing host API (b2World::CreateBody / b2Body::CreateFixture in libbox2d_b200.so), settled,
and stepped with b2World::Step(dt, 8, 3, b2CudaStepExecutor&).
  value  body-steps/s with the world resident on the device (no per-step host copies)
  e2e    the same through the same call with the host buffers in the loop: every step uploads that step's user
         input (forces on 1% of the bodies, edited through b2Body::ApplyForceToCenter) and downloads every body's
         state into the host mirror that b2Body::GetPosition() reads, plus the begin/end touch events
  roofline      the dominant kernel (SolverVelocityPersistentKernel, one cooperative launch per step) against the
                measured HBM peak; traffic = its DRAM bytes from the committed ncu capture
  cpu_baseline  the compiled reference (oracle/_ref) with its own b2ThreadPoolTaskExecutor on the host cores, on a
                bounded sample: a narrower pile of the same depth, started from the device-settled state
Reference arm (--impl reference): the reference's own CPU implementation alone, on a narrower pile of the same
depth (a bounded sample of the workload), settled by the reference itself.

N > 1 (torchrun): one pile of N x bodies cut into x-strips, one per rank, with ghost bodies and per-iteration halo
exchange through NVLink peer mailboxes inside the persistent solver kernel (weak scaling; no NCCL on the data path;
times are max-reduced over NCCL).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "box2d-mt_b200", "python"))

import numpy as np  # noqa: E402

ROWS = 100                 # pile depth (bodies per column)
DT, VEL_ITERS, POS_ITERS = 1.0 / 60.0, 8, 3
SETTLE_STEPS = 300
# algorithmic bytes per unit of work
BYTES_VELOCITY_ITER = 220  # per touching contact per velocity iteration: 180 R + 40 W
BYTES_WARM_START = 128
BYTES_STORE = 32
BYTES_POSITION_ITER = 136
BYTES_INTEGRATE_POS = 48   # per body
BYTES_PER_BODY = 212
BYTES_PER_PROXY = 120
BYTES_PER_CONTACT_NARROW = 208
BYTES_PER_CONSTRAINT_STEP = 2794


def ncu_traffic(n_constraints):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the roofline kernel, from the committed
    `ncu --set full` capture of this workload (profiles/r1_solver_traffic.json); scaled by the constraint count of
    this run when the capture had a different one.  None when there is no capture."""
    path = os.path.join(ROOT, "profiles", "r1_solver_traffic.json")
    try:
        with open(path) as f:
            cap = json.load(f)
        return float(cap["dram_bytes_per_launch"]) * float(n_constraints) / float(cap["constraints"])
    except Exception:
        return None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.stop_flag = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        self.stop_flag = True
        if self.nv is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def settle_reference(world, steps):
    for _ in range(steps):
        world.step(DT, VEL_ITERS, POS_ITERS)


def time_reference(world, warmup, steps):
    for _ in range(warmup):
        world.step(DT, VEL_ITERS, POS_ITERS)
    t0 = time.perf_counter()
    for _ in range(steps):
        world.step(DT, VEL_ITERS, POS_ITERS)
    return time.perf_counter() - t0


def reference_world_from_states(scene_arrays, states, gravity, flags, threads):
    """A reference b2World whose bodies start at the given (device-settled) states."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref
    b, s, f = scene_arrays
    b = b.copy()
    b["px"], b["py"], b["angle"] = states["px"], states["py"], states["a"]
    b["vx"], b["vy"], b["w"] = states["vx"], states["vy"], states["w"]
    return ref.RefWorld(arrays=(b, s, f), gravity=gravity, world_flags=flags, threads=threads)


def run_reference_arm(args, rank, world_size):
    """The reference's own CPU implementation (oracle/_ref built from /root/reference) on the host cores."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref
    import scenes
    cores = os.cpu_count() or 1
    threads = min(8, cores)  # b2_maxThreads = 8 is a compile-time cap of the reference (b2Settings.h:165)
    columns = max(16, args.ref_bodies // ROWS)
    scene = scenes.pile(columns, ROWS)
    w = ref.RefWorld(scene, threads=threads)
    n = w.counts()[0]
    settle_reference(w, args.ref_settle)
    elapsed = time_reference(w, args.warmup, args.steps)
    value = n * args.steps / elapsed
    sample = ("pile %d columns x %d rows (%d bodies, a narrower strip of the 1M-body pile), settled %d steps by the "
              "reference, then %d timed steps" % (columns, ROWS, n, args.ref_settle, args.steps))
    print(json.dumps({
        "impl": "reference", "metric": "body-steps/sec", "value": value, "unit": "body-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "pile_1m (1M-body mixed polygon/circle pile, 60 Hz, 8/3 iterations)",
                   "bodies_per_gpu": args.bodies, "sample_bodies": n},
        "cpu_baseline": {"value": value, "unit": "body-steps/s", "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "body-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def bind_to_gpu_cpus(local_rank):
    """Run this rank on the CPUs next to its GPU (NVML's affinity mask), so that the page-locked body mirror is
    allocated and read on the GPU's own NUMA node.  Best effort: returns a description or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return "cpus %d-%d (%d)" % (allowed[0], allowed[-1], len(allowed))
    except Exception:
        pass
    return None


def run_product_arm(args, rank, local_rank, world_size):
    import b2cuda
    import b2host
    import scenes
    import b2cuda_types as T

    if b2cuda.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")

    numa = bind_to_gpu_cpus(local_rank)
    columns = max(16, args.bodies // ROWS)
    dist = None
    if world_size > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl")

    t0 = time.perf_counter()
    if world_size == 1:
        scene = scenes.pile(columns, ROWS, seed=0)
        world = b2host.HostWorld(scene, device=local_rank, download_bodies=False, events=False)
        n_bodies = world.counts()[0]
    else:
        # one pile of world_size x bodies, cut into x-strips: every rank holds its strip, the static container and
        # ghost copies of the next strip's boundary bodies; the solver kernels exchange halo state through NVLink
        # peer mailboxes (b2cuShard*)
        import b2shard
        scene = scenes.pile(columns * world_size, ROWS, seed=0)
        plan, _ = b2shard.rank_plan(scene.arrays(), rank, world_size, args.margin)
        world = b2host.HostWorld(arrays=plan.arrays, gravity=scene.gravity, world_flags=scene.world_flags,
                                 device=local_rank, download_bodies=False, events=False)
        world.shard_configure(rank, world_size, plan.ghost_local, plan.export_local)
        lower, upper = b2shard.exchange_links(dist, rank, world_size, world.shard_link())
        world.shard_connect(lower, upper)
        n_bodies = world.counts()[0] - len(plan.ghost_local)
        dist.barrier()
    build_s = time.perf_counter() - t0

    # settle (setup, untimed)
    for _ in range(args.settle):
        world.step(DT, VEL_ITERS, POS_ITERS)

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        import b2shard
        return b2shard.reduce_scalar(dist, x, "max", device="cuda")

    def sum_over_ranks(x):
        if dist is None:
            return x
        import b2shard
        return b2shard.reduce_scalar(dist, x, "sum", device="cuda")

    # ---- value: world resident on the device ----
    for _ in range(args.warmup):
        world.step(DT, VEL_ITERS, POS_ITERS)
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    infos = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        world.step(DT, VEL_ITERS, POS_ITERS)   # returns after the device step has completed (stream synchronised)
        infos.append(world.step_info())
    elapsed = time.perf_counter() - t0
    barrier()
    clocks = sampler.result()
    elapsed = max_over_ranks(elapsed)
    device_ms = max_over_ranks(sum(float(i["step"]) for i in infos))

    # ---- e2e: host buffers in the loop ----
    world.set_options(True, True)
    block = max(1, n_bodies // 100)
    for k in range(max(1, args.warmup // 2)):
        world.apply_force_range(1 + (k * block) % max(1, n_bodies - block), block, 0.0, 0.05)
        world.step(DT, VEL_ITERS, POS_ITERS)
    barrier()
    t0 = time.perf_counter()
    checksum = 0.0
    host_ms = np.zeros(4, np.float64)
    apply_s = 0.0
    for k in range(args.steps):
        ta = time.perf_counter()
        world.apply_force_range(1 + (k * block) % max(1, n_bodies - block), block, 0.0, 0.05)
        apply_s += time.perf_counter() - ta
        world.step(DT, VEL_ITERS, POS_ITERS)
        # the step has left every body's new state in the host mirror; read back the block that was pushed
        checksum += world.sum_y(1 + (k * block) % max(1, n_bodies - block), block)
        host_ms += world.host_timings()
    e2e_elapsed = max_over_ranks(time.perf_counter() - t0)
    barrier()
    h2d = block * T.BODY.itemsize
    d2h = world.counts()[0] * T.BODY_STATE.itemsize + 8 * int(infos[-1]["beginCount"] + infos[-1]["endCount"])
    world.set_options(False, False)
    total_bodies = int(sum_over_ranks(n_bodies))

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    last = infos[-1]
    n_constraints = float(np.mean([int(i["constraintCount"]) for i in infos]))
    n_contacts = float(np.mean([int(i["contactCount"]) for i in infos]))
    # the solver kernel (one persistent cooperative launch per step: warm start, velocity iterations, impulse store,
    # position integration, position iterations) is timed by the CUDA events around it (solveVelocity+solvePosition)
    vel_ms = float(np.mean([float(i["solveVelocity"]) for i in infos]))
    colours = float(np.mean([int(i["colourCount"]) for i in infos]))
    peak, peak_src = measured_peaks()
    vel_bytes = (n_constraints * (BYTES_WARM_START + BYTES_VELOCITY_ITER * VEL_ITERS + BYTES_STORE) +
                 n_bodies * BYTES_INTEGRATE_POS)
    achieved = vel_bytes / (vel_ms * 1e-3) / 1e9 if vel_ms > 0 else 0.0
    step_bytes = (n_bodies * BYTES_PER_BODY + n_bodies * BYTES_PER_PROXY + n_contacts * BYTES_PER_CONTACT_NARROW +
                  n_constraints * BYTES_PER_CONSTRAINT_STEP)
    ms_per_step = 1e3 * elapsed / args.steps
    phases = {k: float(np.mean([float(i[k]) for i in infos])) for k in
              ("collide", "solveTraversal", "solveInit", "solveVelocity", "solvePosition", "broadphase", "solveTOI")}

    # ---- cpu_baseline: the compiled reference on a narrower strip, started from a device-settled state ----
    cpu = None
    if not args.no_cpu_baseline:
        try:
            cols_s = max(16, args.cpu_bodies // ROWS)
            scene_s = scenes.pile(cols_s, ROWS, seed=1234)
            gw = b2host.HostWorld(scene_s, device=local_rank, download_bodies=False, events=False)
            for _ in range(args.settle):
                gw.step(DT, VEL_ITERS, POS_ITERS)
            states = gw.bodies()
            cores = os.cpu_count() or 1
            threads = min(8, cores)
            rw = reference_world_from_states(scene_s.arrays(), states, scene_s.gravity, scene_s.world_flags, threads)
            ns = rw.counts()[0]
            cpu_steps = 4
            cpu_elapsed = time_reference(rw, 2, cpu_steps)
            cpu = {"value": ns * cpu_steps / cpu_elapsed, "unit": "body-steps/s", "cores": threads, "kind": "reference",
                   "sample": "pile %d columns x %d rows (%d bodies) from the device-settled state, 2 warm-up + %d timed "
                             "steps of the compiled reference with b2ThreadPoolTaskExecutor(%d threads; host has %d "
                             "cores; b2_maxThreads caps it at 8)" % (cols_s, ROWS, ns, cpu_steps, threads, cores)}
        except Exception as e:  # the oracle is only the checker: its absence must not hide the product numbers
            cpu = {"value": None, "unit": "body-steps/s", "cores": 0, "kind": "reference", "sample": "unavailable: %r" % (e,)}

    print(json.dumps({
        "metric": "body-steps/sec", "value": total_bodies * args.steps / elapsed, "unit": "body-steps/s",
        "n_gpus": world_size, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "pile_1m (BASELINE.json configs[4]: mixed polygon/circle pile in a wide static container, "
                               "60 Hz, 8 velocity / 3 position iterations, no sleeping, settled %d steps)" % args.settle,
                   "bodies_per_gpu": n_bodies, "columns": columns, "rows": ROWS,
                   "contacts": n_contacts, "constraints": n_constraints, "colours": colours,
                   "l2": "working set per step (%.1f GB algorithmic) exceeds the 126 MB L2; no flush needed"
                         % (step_bytes / 1e9),
                   "sharding": ("x-strips of one %d-body pile, ghost bodies within %.1f m of the strip boundary, halo "
                                "exchange through NVLink peer mailboxes inside the solver kernel (2 per iteration)"
                                % (total_bodies, args.margin)) if world_size > 1 else "single GPU"},
        "device_ms_per_step": device_ms / args.steps,
        "phases_ms": phases,
        "step_roofline_frac": (step_bytes / (ms_per_step * 1e-3) / 1e9) / peak,
        "e2e": {"value": total_bodies * args.steps / e2e_elapsed, "unit": "body-steps/s",
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e_elapsed / args.steps,
                "host_ms": dict(zip(("apply_forces", "upload", "step", "download", "events"),
                                    [1e3 * apply_s / args.steps] + [float(x) / args.steps for x in host_ms]))},
        "gpu_launches": int(sum(int(i["kernelLaunches"]) for i in infos)),
        "roofline": {"bound": "hbm", "kernel": "SolverVelocityPersistentKernel (warm start + 8 velocity iterations + impulse "
                                                        "store + position integration, one cooperative launch per step)",
                     "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": ncu_traffic(n_constraints),
                     "ms_per_launch": vel_ms,
                     "bytes_per_unit": "per touching contact: 128 warm start + 220 x 8 velocity + 32 store = 1920 B, "
                                       "+ 48 B per body (SURVEY.md 8d)"},
        "cpu_baseline": cpu,
        "clocks": clocks,
        "build_s": build_s, "cpu_affinity": numa, "checksum": checksum, "last_step": {k: int(last[k]) for k in
                                                               ("contactCount", "constraintCount", "colourCount",
                                                                "overflowCount", "moveCount", "kernelLaunches")},
    }))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="product", choices=["product", "reference"])
    ap.add_argument("--bodies", type=int, default=1000000, help="bodies per GPU")
    ap.add_argument("--settle", type=int, default=SETTLE_STEPS)
    ap.add_argument("--cpu-bodies", type=int, default=25000, help="bodies of the cpu_baseline sample")
    ap.add_argument("--ref-bodies", type=int, default=50000, help="bodies of the --impl reference sample")
    ap.add_argument("--ref-settle", type=int, default=150)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--margin", type=float, default=3.0, help="ghost margin of a strip boundary (m)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world_size)
    else:
        run_product_arm(args, rank, local_rank, world_size)


if __name__ == "__main__":
    main()
