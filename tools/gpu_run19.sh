set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/v_bench_small.json 2> gpurun_out/v_bench_small.err
B200MOBY_THREAD_BUDGET=16 B200MOBY_HARD_COST=16 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/v_bench_small_b16.json 2> gpurun_out/v_bench_small_b16.err
B200MOBY_THREAD_BUDGET=12 B200MOBY_HARD_COST=20 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/v_bench_small_b12h20.json 2> gpurun_out/v_bench_small_b12h20.err
grep -h -o '"value": [0-9.]*, "unit": "env-steps/s", "n_gpus"' gpurun_out/v_bench_small*.json
timeout 300 python bench.py --workload ur10 --steps 20 --warmup 3 > gpurun_out/v_bench_ur10.json 2> gpurun_out/v_bench_ur10.err
timeout 300 python bench.py --workload feeder --steps 20 --warmup 3 > gpurun_out/v_bench_feeder.json 2> gpurun_out/v_bench_feeder.err
timeout 500 python bench.py --workload stacks --envs-per-gpu 256 --steps 3 --warmup 3 --preroll 2 > gpurun_out/v_bench_stacks256.json 2> gpurun_out/v_bench_stacks256.err
ls -la gpurun_out/v_*
