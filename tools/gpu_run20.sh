set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_rc.py tests/test_noslip.py -m gpu -q > gpurun_out/pytest_gpu_rc.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_rc.log
tail -4 gpurun_out/pytest_gpu_rc.log
timeout 400 python bench.py --workload stacks --envs-per-gpu 256 --steps 3 --warmup 3 --preroll 2 --no-cpu-baseline > gpurun_out/v_bench_stacks256.json 2> gpurun_out/v_bench_stacks256.err
ls -la gpurun_out/v_bench_stacks256.json
