set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err
timeout 600 python bench.py --workload ur10 --steps 20 --warmup 3 > gpurun_out/bench_ur10.json 2> gpurun_out/bench_ur10.err
timeout 900 python bench.py --workload stacks --steps 10 --warmup 3 > gpurun_out/bench_stacks.json 2> gpurun_out/bench_stacks.err
tail -3 gpurun_out/bench_*.err
ls -la gpurun_out
