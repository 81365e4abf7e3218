#!/bin/bash
# usage: variant_bench.sh "<nvcc flags>" ... : rebuild libb2cuda.so with each flag set and run the 1-GPU bench
for flags in "$@"; do
  echo "=== variant: $flags"
  B2CU_NVCC_FLAGS="$flags" python box2d-mt_b200/build.py --force > /dev/null 2>&1 || { echo build failed; continue; }
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['phases_ms'].items()}, round(d['roofline']['frac'],3))"
done
