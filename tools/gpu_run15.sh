set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_small_hard64.json 2> gpurun_out/bench_small_hard64.err
B200MOBY_HARD_COST=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_small_hard0.json 2> gpurun_out/bench_small_hard0.err
B200MOBY_HARD_COST=24 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_small_hard24.json 2> gpurun_out/bench_small_hard24.err
B200MOBY_PIVOT_BUDGET=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_small_hard64_nobudget.json 2> gpurun_out/bench_small_hard64_nobudget.err
B200MOBY_HARD_COST=24 B200MOBY_PIVOT_BUDGET=48 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_small_hard24_b48.json 2> gpurun_out/bench_small_hard24_b48.err
grep -h -o '"value": [0-9.]*, "unit": "env-steps/s", "n_gpus"' gpurun_out/bench_small_hard*.json
timeout 400 python bench.py --workload ur10 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ur10_b.json 2> gpurun_out/bench_ur10_b.err
timeout 600 ncu --set full --import-source on --clock-control none -k regex:impact_warp_kernel -s 604 -c 1 -o gpurun_out/ncu_ur10_impact -f python bench.py --workload ur10 --steps 2 --warmup 3 --preroll 100 --no-cpu-baseline > gpurun_out/ncu_ur10_impact.log 2>&1
timeout 600 python tools/stacks_probe.py 256 5 > gpurun_out/stacks_probe.log 2>&1
cat gpurun_out/stacks_probe.log | cut -c1-600
timeout 900 python tools/drift_report.py > gpurun_out/drift_report.json 2> gpurun_out/drift_report.err
tail -3 gpurun_out/drift_report.err
ls gpurun_out
