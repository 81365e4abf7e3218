#!/usr/bin/env python3
"""One-line digest of bench.py JSON lines (developer aid)."""
import json, sys
for p in sys.argv[1:]:
    try:
        d = json.loads(open(p).read().strip().splitlines()[-1])
    except Exception as e:
        print(p, "unreadable", e); continue
    ph = {k: round(v, 3) for k, v in d.get("phases_ms", {}).items()}
    print("%s: N=%d %s value %.3e ms/step %.3f e2e %.3f phases %s" % (p, d["n_gpus"], d["scaling"], d["value"], d["ms_per_step"],
                                                                      d["e2e"].get("ms_per_step", 0) or 0, ph))
