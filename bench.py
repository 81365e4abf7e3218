#!/usr/bin/env python3
"""bench.py -- body-steps/s of b2World::Step on the BASELINE.json pile workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--bodies B] [--impl reference]
                    [--workload pile|add_pair|tumbler|stacks_awake|stacks_asleep] [--scaling weak|strong]

Product arm (default): the 1M-body mixed polygon/circle pile (BASELINE.json configs[4]) is built
through the reference-facing host API (b2World::CreateBody / b2Body::CreateFixture in libbox2d_b200.so), settled,
and stepped with b2World::Step(dt, 8, 3, b2CudaStepExecutor&).
  value  body-steps/s with the world resident on the device (no per-step host copies)
  e2e    the same through the same call with the host buffers in the loop: every step uploads that step's user
         input (forces on 1% of the bodies, edited through b2Body::ApplyForceToCenter) and downloads every body's
         state into the host mirror that b2Body::GetPosition() reads, plus the begin/end touch events
  roofline      the dominant kernel (SolverVelocityFlowKernel, one launch per step) against the
                measured HBM peak; traffic = its DRAM bytes from the committed ncu capture
  cpu_baseline  the compiled reference (oracle/_ref) with its own b2ThreadPoolTaskExecutor on the host cores, on a
                bounded sample: a narrower pile of the same depth, started from the device-settled state
Reference arm (--impl reference): the reference's own CPU implementation alone, on a narrower pile of the same
depth (a bounded sample of the workload), settled by the reference itself.

N > 1 (torchrun): one pile cut into x-strips, one per rank, with ghost bodies; the boundary bodies' rows travel as
versioned 16-byte stores into NVLink peer mailboxes from inside the solver kernels (no NCCL on the data path; times
are max-reduced over NCCL).  Default: strong scaling of the 1M-body pile, with the weak-scaling run (1M bodies per GPU)
measured alongside under other_scaling.
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


def make_workload(name, bodies, scenes):
    """(scene, settle steps, description) of a BASELINE.json configuration."""
    if name == "pile":
        columns = max(16, bodies // ROWS)
        return (scenes.pile(columns, ROWS, seed=0), None,
                "pile_%s (BASELINE.json configs[4]: mixed polygon/circle pile in a wide static container, 60 Hz, 8 velocity "
                "/ 3 position iterations, no sleeping" % ("1m" if bodies >= 1000000 else "%dk" % (bodies // 1000)))
    if name == "add_pair":
        return (scenes.add_pair(10000), 0,
                "add_pair_10k (BASELINE.json configs[1]: 10 000 circles + one bullet box at 150 m/s, zero gravity, continuous "
                "physics on; timed from the first step, through the impact")
    if name == "tumbler":
        return (scenes.tumbler(20000), 600,
                "tumbler_20k (BASELINE.json configs[2]: 20 000 boxes in a rotating kinematic drum, no sleeping")
    if name in ("stacks_awake", "stacks_asleep"):
        scene = scenes.pyramids(476, 20, thick_polygon_ground=True)
        if name == "stacks_awake":
            return (scene, 0, "stacks_100k awake window (BASELINE.json configs[3]: 476 pyramids of 20 rows = 99 960 boxes on a "
                              "thick static ground, sleeping on; the first steps, everything awake")
        return (scene, 700, "stacks_100k asleep window (BASELINE.json configs[3]: 476 pyramids of 20 rows = 99 960 boxes, "
                            "sleeping on; after every island has gone to sleep")
    raise SystemExit("unknown workload %r" % name)


def kernel_traffic():
    """Per-kernel DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the committed `ncu --set full`
    capture of the 1M-body pile (profiles/r2_kernel_traffic.json, made by tools/ncu_traffic.py), with the counts of that
    capture so that a run with other counts can scale them.  {} when there is no capture."""
    path = os.path.join(ROOT, "profiles", "r2_kernel_traffic.json")
    try:
        with open(path) as f:
            return json.load(f)
    except Exception:
        return {}


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


def pow2_threads(cores, cap):
    t = 1
    while 2 * t <= min(cores, cap):
        t *= 2
    return t


def run_reference_arm(args, rank, world_size):
    """The reference's own CPU implementation (oracle/_ref built from /root/reference) on the host cores: its own
    b2World::Step with its own b2ThreadPoolTaskExecutor on the workload of the product arm -- for the pile a narrower
    strip of the same depth (throughput is per body; 1M bodies would take ~3 s per step), settled by the reference
    itself.  Two rows: the reference as shipped (b2_maxThreads = 8, b2Settings.h:165) is the headline of this arm;
    `cpu_rows` adds the copy compiled with the cap raised to 32 on all the host's cores (largest power of two)."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref
    import scenes
    cores = os.cpu_count() or 1
    rows = []
    for variant, cap in ((True, 8), ("mt", 32)):
        threads = pow2_threads(cores, cap)
        if variant == "mt" and threads <= 8:
            continue
        try:
            if args.workload == "pile":
                columns = max(16, args.ref_bodies // ROWS)
                scene, settle = scenes.pile(columns, ROWS), args.ref_settle
                what = "pile %d columns x %d rows (a narrower strip of the %d-body pile)" % (columns, ROWS, args.bodies)
            else:
                scene, settle, what = make_workload(args.workload, args.bodies, scenes)
            w = ref.RefWorld(scene, threads=threads, stock_libm=variant)
            n = w.counts()[0]
            settle_reference(w, settle or 0)
            elapsed = time_reference(w, args.warmup, args.steps)
            rows.append({"value": n * args.steps / elapsed, "unit": "body-steps/s", "cores": threads, "kind": "reference",
                         "ms_per_step": 1e3 * elapsed / args.steps, "bodies": n,
                         "sample": "%s, %d bodies, settled %d steps by the reference, then %d warm-up + %d timed steps; "
                                   "b2ThreadPoolTaskExecutor with %d threads (host has %d cores; b2_maxThreads = %d%s)"
                                   % (what, n, settle or 0, args.warmup, args.steps, threads, cores, cap,
                                      "" if cap == 8 else ", patched copy, oracle/build_ref.py build_mt")})
        except Exception as e:
            rows.append({"value": None, "unit": "body-steps/s", "cores": threads, "kind": "reference",
                         "sample": "unavailable: %r" % (e,)})
    head = rows[0]
    print(json.dumps({
        "impl": "reference", "metric": "body-steps/sec", "value": head["value"], "unit": "body-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": head.get("ms_per_step"),
        "higher_is_better": True, "scaling": args.scaling or ("strong" if world_size > 1 else "weak"), "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "bodies_per_gpu": args.bodies, "sample_bodies": head.get("bodies")},
        "cpu_baseline": head, "cpu_rows": rows,
        "e2e": {"value": head["value"], "unit": "body-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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
    dist = None
    if world_size > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl")

    t0 = time.perf_counter()
    scaling = args.scaling or ("strong" if world_size > 1 else "weak")

    def sharded_pile(mode):
        """ONE pile cut into x-strips: every rank holds its strip, the static container and ghost copies of the next
        strip's boundary bodies; the halo bodies' rows cross NVLink inside the solver kernels (b2cuShard*).
        weak: world_size x bodies in total (bodies per GPU fixed); strong: `bodies` in total (the 1M-body pile cut N ways)."""
        import b2shard
        cols = max(16, args.bodies // ROWS) if mode == "weak" else max(16 * world_size, args.bodies // ROWS) // world_size
        sc = scenes.pile(cols * world_size, ROWS, seed=0)
        sc.world_flags &= ~T.WORLD_CONTINUOUS  # the container is thick-shape and nothing is a bullet: no TOI candidates
        plan, _ = b2shard.rank_plan(sc.arrays(), rank, world_size, args.margin)
        wd = b2host.HostWorld(arrays=plan.arrays, gravity=sc.gravity, world_flags=sc.world_flags,
                              device=local_rank, download_bodies=False, events=False)
        wd.shard_configure(rank, world_size, plan.ghost_local, plan.export_local)
        lower, upper = b2shard.exchange_links(dist, rank, world_size, wd.shard_link())
        wd.shard_connect(lower, upper)
        dist.barrier()
        return wd, wd.counts()[0] - len(plan.ghost_local), cols, sc

    if world_size == 1:
        scene, settle, workload = make_workload(args.workload, args.bodies, scenes)
        if settle is None:
            settle = args.settle
        workload += ", settled %d steps)" % settle
        world = b2host.HostWorld(scene, device=local_rank, download_bodies=False, events=False)
        n_bodies = world.counts()[0]
        columns = max(16, args.bodies // ROWS)
    else:
        if args.workload != "pile":
            raise SystemExit("bench.py: only the pile workload is sharded")
        settle = args.settle
        world, n_bodies, columns, scene = sharded_pile(scaling)
        workload = ("pile_%dk x %d GPUs, %s scaling (BASELINE.json configs[4]: one mixed polygon/circle pile in a wide static "
                    "container, 60 Hz, 8 velocity / 3 position iterations, no sleeping, settled %d steps)"
                    % (columns * ROWS // 1000, world_size, scaling, settle))
    build_s = time.perf_counter() - t0

    # settle (setup, untimed)
    for _ in range(settle):
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

    # ---- N > 1: the other scaling mode as a side measurement (same code path, device-resident value only) ----
    other = None
    if world_size > 1 and not args.no_other_scaling:
        other_mode = "weak" if scaling == "strong" else "strong"
        del world
        w2, n2, cols2, _ = sharded_pile(other_mode)
        for _ in range(settle):
            w2.step(DT, VEL_ITERS, POS_ITERS)
        for _ in range(args.warmup):
            w2.step(DT, VEL_ITERS, POS_ITERS)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            w2.step(DT, VEL_ITERS, POS_ITERS)
        el2 = max_over_ranks(time.perf_counter() - t0)
        barrier()
        total2 = int(sum_over_ranks(n2))
        other = {"scaling": other_mode, "bodies_total": total2, "bodies_per_gpu": n2, "ms_per_step": 1e3 * el2 / args.steps,
                 "value": total2 * args.steps / el2, "unit": "body-steps/s", "steps": args.steps}
        del w2

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

    # ---- per-phase rooflines: algorithmic bytes of SURVEY.md 8d with this run's counts over the CUDA-event time of the
    # phase; traffic = the phase's kernels' dram bytes from the committed ncu capture, scaled to this run's counts ----
    n_proxies = float(last["proxyCount"])
    n_new = float(np.mean([int(i["newContactCount"]) for i in infos]))
    phase_bytes = {
        "collide": n_contacts * BYTES_PER_CONTACT_NARROW,
        "solveTraversal": n_constraints * 50.0,
        "solveInit": n_constraints * 416.0 + n_bodies * 84.0,
        "solveVelocity": vel_bytes,
        "solvePosition": n_constraints * BYTES_POSITION_ITER * POS_ITERS + n_bodies * 80.0,
        "broadphase": n_proxies * (BYTES_PER_PROXY + 20.0) + 8.0 * n_new + 8.0 * n_contacts,
    }
    phase_units = {"collide": ("contacts", n_contacts), "solveTraversal": ("constraints", n_constraints),
                   "solveInit": ("constraints", n_constraints), "solveVelocity": ("constraints", n_constraints),
                   "solvePosition": ("constraints", n_constraints), "broadphase": ("proxies", n_proxies)}
    cap = kernel_traffic()
    phase_rooflines = []
    for name in ("collide", "solveTraversal", "solveInit", "solveVelocity", "solvePosition", "broadphase"):
        ms = phases[name]
        ach = phase_bytes[name] / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        traffic = None
        try:
            unit, count = phase_units[name]
            traffic = float(cap["phases"][name]["dram_bytes"]) * count / float(cap["counts"][unit])
        except Exception:
            pass
        phase_rooflines.append({"phase": name, "ms": ms, "bound": "hbm", "algorithmic_bytes": phase_bytes[name],
                                "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                                "kernels": (cap.get("phases", {}).get(name, {}) or {}).get("kernels")})

    # ---- cpu_baseline: the compiled reference on a bounded sample of the same workload.  For the pile: a narrower
    # strip of the same depth started from a device-settled state (settling is setup, not the thing measured) ----
    cpu = None
    cpu_rows = []
    if not args.no_cpu_baseline and world_size == 1:
        try:
            cores = os.cpu_count() or 1
            if args.workload == "pile":
                cols_s = max(16, args.cpu_bodies // ROWS)
                scene_s = scenes.pile(cols_s, ROWS, seed=1234)
                gw = b2host.HostWorld(scene_s, device=local_rank, download_bodies=False, events=False)
                for _ in range(settle):
                    gw.step(DT, VEL_ITERS, POS_ITERS)
                states = gw.bodies()
                del gw
                what = "pile %d columns x %d rows from the device-settled state" % (cols_s, ROWS)
                cpu_steps, cpu_warm = 4, 2
            else:
                scene_s, _, what = make_workload(args.workload, args.bodies, scenes)
                states = None
                cpu_steps, cpu_warm = max(4, min(args.steps, 30)), 0
            for variant, capn in ((True, 8), ("mt", 32)):
                threads = pow2_threads(cores, capn)
                if variant == "mt" and threads <= 8:
                    continue
                sys.path.insert(0, os.path.join(ROOT, "oracle"))
                import ref
                if states is not None:
                    b_, s_, f_ = scene_s.arrays()
                    b_ = b_.copy()
                    b_["px"], b_["py"], b_["angle"] = states["px"], states["py"], states["a"]
                    b_["vx"], b_["vy"], b_["w"] = states["vx"], states["vy"], states["w"]
                    rw = ref.RefWorld(arrays=(b_, s_, f_), gravity=scene_s.gravity, world_flags=scene_s.world_flags,
                                      threads=threads, stock_libm=variant)
                else:
                    rw = ref.RefWorld(scene_s, threads=threads, stock_libm=variant)
                    for _ in range(settle):
                        rw.step(DT, VEL_ITERS, POS_ITERS)
                ns = rw.counts()[0]
                cpu_elapsed = time_reference(rw, cpu_warm, cpu_steps)
                cpu_rows.append({"value": ns * cpu_steps / cpu_elapsed, "unit": "body-steps/s", "cores": threads,
                                 "kind": "reference", "ms_per_step": 1e3 * cpu_elapsed / cpu_steps, "bodies": ns,
                                 "sample": "%s (%d bodies), %d warm-up + %d timed steps of the compiled reference with "
                                           "b2ThreadPoolTaskExecutor(%d threads; host has %d cores; b2_maxThreads = %d%s)"
                                           % (what, ns, cpu_warm, cpu_steps, threads, cores, capn,
                                              "" if capn == 8 else ", patched copy")})
                del rw
            cpu = cpu_rows[0]
        except Exception as e:  # the oracle is only the checker: its absence must not hide the product numbers
            cpu = {"value": None, "unit": "body-steps/s", "cores": 0, "kind": "reference", "sample": "unavailable: %r" % (e,)}

    print(json.dumps({
        "metric": "body-steps/sec", "value": total_bodies * args.steps / elapsed, "unit": "body-steps/s",
        "n_gpus": world_size, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "other_scaling": other,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload,
                   "bodies_per_gpu": n_bodies, "bodies_total": total_bodies, "columns": columns, "rows": ROWS,
                   "contacts": n_contacts, "constraints": n_constraints, "colours": colours,
                   "continuous_physics": bool(scene.world_flags & T.WORLD_CONTINUOUS),
                   "toi_sub_steps_per_step": float(np.mean([int(i["toiSubSteps"]) for i in infos])),
                   "l2": ("working set per step (%.2f GB algorithmic) exceeds the 126 MB L2; no flush needed" % (step_bytes / 1e9))
                         if step_bytes > 2.0e8 else
                         ("working set per step %.3f GB algorithmic: the step's kernels stream more than L2 between two uses "
                          "of a row only partly; no explicit flush (every step rewrites all rows)" % (step_bytes / 1e9)),
                   "sharding": ("x-strips of one %d-body pile, ghost bodies within %.1f m of the strip boundary; boundary "
                                "rows pushed as versioned 16-byte stores into NVLink peer mailboxes by the solver kernels"
                                % (total_bodies, args.margin)) if world_size > 1 else "single GPU"},
        "device_ms_per_step": device_ms / args.steps,
        "phases_ms": phases,
        "phase_rooflines": phase_rooflines,
        "step_roofline_frac": (step_bytes / (ms_per_step * 1e-3) / 1e9) / peak,
        "e2e": {"value": total_bodies * args.steps / e2e_elapsed, "unit": "body-steps/s",
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e_elapsed / args.steps,
                "host_ms": dict(zip(("apply_forces", "upload", "step", "download", "events"),
                                    [1e3 * apply_s / args.steps] + [float(x) / args.steps for x in host_ms]))},
        "gpu_launches": int(sum(int(i["kernelLaunches"]) for i in infos)),
        "roofline": {"bound": "hbm", "kernel": "SolverVelocityFlowKernel (warm start + 8 velocity iterations + impulse store + "
                                                        "position integration; one launch per step, ordered by per-body row versions)",
                     "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": phase_rooflines[3]["traffic"] or ncu_traffic(n_constraints),
                     "ms_per_launch": vel_ms,
                     "bytes_per_unit": "per touching contact: 128 warm start + 220 x 8 velocity + 32 store = 1920 B, "
                                       "+ 48 B per body (SURVEY.md 8d)"},
        "cpu_baseline": cpu, "cpu_rows": cpu_rows,
        "clocks": clocks,
        "build_s": build_s, "cpu_affinity": numa, "checksum": checksum, "last_step": {k: int(last[k]) for k in
                                                               ("contactCount", "constraintCount", "colourCount",
                                                                "overflowCount", "moveCount", "kernelLaunches",
                                                                "toiCandidateCount", "toiSubSteps")},
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
    ap.add_argument("--cpu-bodies", type=int, default=200000, help="bodies of the cpu_baseline sample (pile)")
    ap.add_argument("--ref-bodies", type=int, default=100000, help="bodies of the --impl reference sample (pile)")
    ap.add_argument("--ref-settle", type=int, default=120)
    ap.add_argument("--workload", default="pile", choices=["pile", "add_pair", "tumbler", "stacks_awake", "stacks_asleep"])
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"],
                    help="N > 1: strong (default) = `bodies` in total, the 1M-body pile of BASELINE.json cut N ways; weak = "
                         "`bodies` per GPU.  The other mode is measured too and reported under other_scaling")
    ap.add_argument("--no-other-scaling", action="store_true")
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
