set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
run() { name=$1; shift; env "$@" timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/w_bench_small_$name.json 2> gpurun_out/w_bench_small_$name.err; }
run shift2 A=1
run shift1 B200MOBY_COST_DECAY_SHIFT=1
run shift3 B200MOBY_COST_DECAY_SHIFT=3
run shift0 B200MOBY_COST_DECAY_SHIFT=0
grep -h -o '"value": [0-9.]*, "unit": "env-steps/s", "n_gpus"' gpurun_out/w_bench_small_*.json
timeout 300 python bench.py --workload stacks --envs-per-gpu 256 --steps 3 --warmup 3 --preroll 2 --no-cpu-baseline > gpurun_out/w_bench_stacks256.json 2> gpurun_out/w_bench_stacks256.err
