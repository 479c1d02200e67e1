set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for R in 64 96; do
  B200MOBY_LIB=$PWD/moby_b200/libb200moby_r$R.so timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r$R.json 2> gpurun_out/bench_r$R.err
  B200MOBY_LIB=$PWD/moby_b200/libb200moby_r$R.so B200MOBY_IMPACT_THREADS=32 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r${R}_warp.json 2> gpurun_out/bench_r${R}_warp.err
done
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r128.json 2> gpurun_out/bench_r128.err
B200MOBY_IMPACT_THREADS=32 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r128_warp.json 2> gpurun_out/bench_r128_warp.err
B200MOBY_LIB=$PWD/moby_b200/libb200moby_r64.so B200MOBY_IMPACT_THREADS=32 B200MOBY_ADV_WPB=8 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r64_warp_adv8.json 2> gpurun_out/bench_r64_warp_adv8.err
grep -h -o '"value": [0-9.]*, "unit": "env-steps/s", "n_gpus"' gpurun_out/bench_r*.json
for f in gpurun_out/bench_r*.json; do echo $f $(head -c 60 $f); done
B200MOBY_LIB=$PWD/moby_b200/libb200moby_r64.so timeout 600 python -m pytest tests/test_gpu_sim.py -m gpu -x -q 2>&1 | tail -3
