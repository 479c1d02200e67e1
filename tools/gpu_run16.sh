set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_small_$name.json 2> gpurun_out/bench_small_$name.err; }
run t24 A=1
run t48 B200MOBY_THREAD_BUDGET=48 B200MOBY_HARD_COST=48
run t12 B200MOBY_THREAD_BUDGET=12 B200MOBY_HARD_COST=12
run t24_s64 B200MOBY_STRAGGLER_THREADS=64
run t24_s128 B200MOBY_STRAGGLER_THREADS=128
run t24_n24 B200MOBY_THREAD_NMAX=24
run nothread B200MOBY_IMPACT_THREAD=0 B200MOBY_HARD_COST=0 B200MOBY_PIVOT_BUDGET=0
grep -h -o '"value": [0-9.]*, "unit": "env-steps/s", "n_gpus"' gpurun_out/bench_small_t*.json gpurun_out/bench_small_nothread.json
timeout 400 python bench.py --workload ur10 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ur10_t24.json 2> gpurun_out/bench_ur10_t24.err
B200MOBY_IMPACT_THREAD=0 timeout 400 python bench.py --workload ur10 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ur10_nothread.json 2> gpurun_out/bench_ur10_nothread.err
timeout 600 python tools/stacks_probe.py 256 4 > gpurun_out/stacks_probe.log 2>&1
cat gpurun_out/stacks_probe.log | cut -c1-500
ls gpurun_out | head -80
