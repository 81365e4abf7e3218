#!/usr/bin/env python3
"""Per-kernel and per-phase DRAM traffic of steady-state steps, from an ncu CSV.

    # on the GPU box (gpurun): settle, then profile STEPS steps
    ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
        --clock-control none --csv --log-file gpurun_out/traffic.csv python tools/profile_step.py 1000000 300 3 > gpurun_out/traffic.txt
    # here
    python tools/ncu_traffic.py gpurun_out/traffic.csv gpurun_out/traffic.txt 3 profiles/r2_kernel_traffic.json

Writes {counts, kernels: {name: {launches_per_step, dram_bytes, ms}}, phases: {phase: {dram_bytes, ms, kernels}}}, all per
step (averaged over the profiled steps).  bench.py scales the phase traffic to its own counts."""
import csv
import json
import re
import sys

PHASE_OF = [
    (r"^Collide|^ApplyWake", "collide"),
    (r"^SolveInitBodies|^Island|^JointUnion|^SelectConstraints|^Colour", "solveTraversal"),
    (r"^IntegrateVelocities|^ConstraintSlot|^InitConstraints", "solveInit"),
    (r"^SolverVelocity|^HaloMask|^WarmStart|^SolveVelocity|^StoreImpulses", "solveVelocity"),
    (r"^SolverPosition|^FlowOverflow|^SolvePosition|^IntegratePositions|^FinalizeBodies|^SleepIslands", "solvePosition"),
    (r"^SyncProxies|^EndStepBodies|^Grid|^Query|^ClearMoved|^MergeMove|^RebuildNew|^BuildLowStart|^Iota|^PackBodyStates",
     "broadphase"),
    (r"^Toi", "solveTOI"),
]
# the scan / sort / compaction primitives (prims.cu) serve the phase of the kernel launched before them
PRIMS = r"^Scan|^Radix|^Compact|^TileSort|^MergeRuns|^Exclusive|^Reduce|^Tile"


def main():
    path, txt, steps, out = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    reader = csv.DictReader(lines)
    per_launch = {}
    order = []
    for r in reader:
        key = r["ID"]
        name = re.sub(r"^void\s+", "", re.sub(r"\(.*$", "", r["Kernel Name"])).split("::")[-1].split("<")[0].strip()
        if key not in per_launch:
            per_launch[key] = {"name": name, "bytes": 0.0, "ns": 0.0}
            order.append(key)
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        m = r["Metric Name"]
        if m.startswith("dram__bytes"):
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
            per_launch[key]["bytes"] += v * scale
        elif m.startswith("gpu__time_duration"):
            scale = {"ns": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1.0, "second": 1e9}.get(unit, 1.0)
            per_launch[key]["ns"] += v * scale
    kernels, phases = {}, {}
    last_phase = "broadphase"
    for key in order:
        e = per_launch[key]
        name = e["name"]
        phase = None
        if re.search(PRIMS, name):
            phase = last_phase
        else:
            for pat, ph in PHASE_OF:
                if re.search(pat, name):
                    phase = ph
                    break
            if phase is None:
                phase = "other"
            last_phase = phase
        k = kernels.setdefault(name, {"launches_per_step": 0.0, "dram_bytes": 0.0, "ms": 0.0, "phase": phase})
        k["launches_per_step"] += 1.0 / steps
        k["dram_bytes"] += e["bytes"] / steps
        k["ms"] += e["ns"] * 1e-6 / steps
        p = phases.setdefault(phase, {"dram_bytes": 0.0, "ms": 0.0, "kernels": []})
        p["dram_bytes"] += e["bytes"] / steps
        p["ms"] += e["ns"] * 1e-6 / steps
        if name not in p["kernels"]:
            p["kernels"].append(name)
    counts = {}
    with open(txt) as f:
        for ln in f:
            m = re.search(r"bodies (\d+) contacts (\d+) constraints (\d+)", ln)
            if m:
                counts = {"bodies": int(m.group(1)), "proxies": int(m.group(1)), "contacts": int(m.group(2)),
                          "constraints": int(m.group(3))}
    doc = {"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none, "
                     "%d steady-state steps of the 1M-body pile (tools/profile_step.py); times are ncu's serialised, cold-cache "
                     "per-launch durations: shares, not absolutes" % steps,
           "counts": counts, "phases": phases, "kernels": kernels}
    with open(out, "w") as f:
        json.dump(doc, f, indent=1)
    tot = sum(p["ms"] for p in phases.values())
    for ph, p in sorted(phases.items(), key=lambda kv: -kv[1]["ms"]):
        print("%-16s %7.3f ms %5.1f%%  %8.1f MB" % (ph, p["ms"], 100 * p["ms"] / tot, p["dram_bytes"] / 1e6))
    for name, k in sorted(kernels.items(), key=lambda kv: -kv[1]["ms"])[:25]:
        print("  %-36s n=%5.1f %7.3f ms %9.1f MB" % (name, k["launches_per_step"], k["ms"], k["dram_bytes"] / 1e6))


if __name__ == "__main__":
    main()
