set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_small_$name.json 2> gpurun_out/bench_small_$name.err; }
run d12 A=1
run d8 B200MOBY_THREAD_BUDGET=8 B200MOBY_HARD_COST=8
run d12_n24 B200MOBY_THREAD_NMAX=24
run d12_s256 B200MOBY_STRAGGLER_THREADS=256
run d12_r1 B200MOBY_ROUNDS=1
grep -h -o '"value": [0-9.]*, "unit": "env-steps/s", "n_gpus"' gpurun_out/bench_small_d*.json
timeout 400 python bench.py --workload ur10 --steps 20 --warmup 3 > gpurun_out/bench_ur10_final.json 2> gpurun_out/bench_ur10_final.err
ls gpurun_out | wc -l
