#!/usr/bin/env python3
"""Developer tool: average the per-kernel lines of B2CU_TRACE=1 over the traced steps (reads stdin)."""
import re, sys
tot, cnt, steps = {}, {}, 0
for line in sys.stdin:
    m = re.match(r"\[b2cu trace\] (\S+)\s+n=\s*(\d+)\s+([0-9.]+) ms", line)
    if m:
        tot[m.group(1)] = tot.get(m.group(1), 0.0) + float(m.group(3))
        cnt[m.group(1)] = cnt.get(m.group(1), 0) + int(m.group(2))
    elif line.startswith("[b2cu trace] step:"):
        steps += 1
    elif line.startswith("{"):
        print(line.rstrip())
steps = max(steps, 1)
print("average over %d traced steps" % steps)
total = 0.0
for k in sorted(tot, key=lambda k: -tot[k]):
    print("%-34s n=%6.1f %8.3f ms" % (k, cnt[k] / steps, tot[k] / steps))
    total += tot[k] / steps
print("%-34s          %8.3f ms" % ("(sum)", total))
