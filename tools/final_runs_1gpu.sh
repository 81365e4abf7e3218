set -x
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2_gpu_tests.txt 2>&1; tail -3 gpurun_out/r2_gpu_tests.txt
python bench.py > gpurun_out/r2_bench_1gpu_final.json 2> gpurun_out/r2_bench_1gpu_final.err
python bench.py --impl reference > gpurun_out/r2_bench_reference_arm_final.json 2> gpurun_out/r2_bench_reference_arm_final.err
python bench.py --bodies 100000 --steps 60 --warmup 5 > gpurun_out/r2_bench_100k_final.json 2>/dev/null
for w in add_pair tumbler stacks_awake stacks_asleep; do python bench.py --workload $w --steps 30 --warmup 3 > gpurun_out/r2_bench_${w}_final.json 2>/dev/null; done
python tools/bench_line.py gpurun_out/r2_bench_1gpu_final.json gpurun_out/r2_bench_100k_final.json gpurun_out/r2_bench_add_pair_final.json gpurun_out/r2_bench_tumbler_final.json gpurun_out/r2_bench_stacks_awake_final.json gpurun_out/r2_bench_stacks_asleep_final.json
tail -c 600 gpurun_out/r2_bench_reference_arm_final.json
