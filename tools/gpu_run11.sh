set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err
timeout 600 python bench.py --workload ur10 --steps 20 --warmup 3 > gpurun_out/bench_ur10.json 2> gpurun_out/bench_ur10.err
timeout 900 python bench.py --workload stacks --steps 10 --warmup 3 > gpurun_out/bench_stacks.json 2> gpurun_out/bench_stacks.err
tail -3 gpurun_out/bench_*.err
cat gpurun_out/bench_*.json
for W in small ur10 stacks; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_$W.csv python bench.py --workload $W --steps 2 --warmup 1 --preroll 20 --no-cpu-baseline > gpurun_out/ncu_$W.log 2>&1
done
ls -la gpurun_out
