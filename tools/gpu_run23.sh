set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/x_bench_small_split.json 2> gpurun_out/x_bench_small_split.err
grep -h -o '"value": [0-9.]*, "unit": "env-steps/s", "n_gpus"' gpurun_out/x_bench_small_split.json
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python bench.py > gpurun_out/x_bench_small_default.json 2> gpurun_out/x_bench_small_default.err
