set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_small_new.json 2> gpurun_out/bench_small_new.err
B200MOBY_ADV_THREAD=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_small_advwarp.json 2> gpurun_out/bench_small_advwarp.err
B200MOBY_WARP_NMAX=24 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_small_w24.json 2> gpurun_out/bench_small_w24.err
B200MOBY_PIVOT_BUDGET=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_small_nobudget.json 2> gpurun_out/bench_small_nobudget.err
B200MOBY_CONCURRENT=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_small_serial.json 2> gpurun_out/bench_small_serial.err
grep -h -o '"value": [0-9.]*, "unit": "env-steps/s", "n_gpus"' gpurun_out/bench_small_new.json gpurun_out/bench_small_advwarp.json gpurun_out/bench_small_w24.json gpurun_out/bench_small_nobudget.json gpurun_out/bench_small_serial.json
timeout 400 python bench.py --workload ur10 --steps 20 --warmup 3 > gpurun_out/bench_ur10.json 2> gpurun_out/bench_ur10.err
timeout 300 python bench.py --workload lcp --steps 10 --warmup 3 > gpurun_out/bench_lcp32.json 2> gpurun_out/bench_lcp32.err
timeout 300 python bench.py --workload lcp --lcp-n 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_lcp8.json 2> gpurun_out/bench_lcp8.err
timeout 300 python bench.py --workload lcp --lcp-n 96 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_lcp96.json 2> gpurun_out/bench_lcp96.err
cat gpurun_out/bench_lcp*.json | cut -c1-900
timeout 600 python tools/stacks_probe.py 256 6 > gpurun_out/stacks_probe.log 2>&1
cat gpurun_out/stacks_probe.log | cut -c1-700
ls gpurun_out
