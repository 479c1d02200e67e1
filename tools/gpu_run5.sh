set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rc.py tests/test_gpu_sim.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python tools/impact_profile.py > gpurun_out/impact_profile.json 2> gpurun_out/impact_profile.err
B200MOBY_CONCURRENT=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_conc1.json 2> gpurun_out/bench_conc1.err
B200MOBY_CONCURRENT=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_conc0.json 2> gpurun_out/bench_conc0.err
B200MOBY_CONCURRENT=1 B200MOBY_IMPACT_THREADS=32 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_conc1_warp.json 2> gpurun_out/bench_conc1_warp.err
B200MOBY_CONCURRENT=1 B200MOBY_PIVOT_BUDGET=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_conc1_nobudget.json 2> gpurun_out/bench_conc1_nobudget.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 4560 -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
grep -h -o '"value": [0-9.]*' gpurun_out/bench_conc*.json
ls -la gpurun_out
