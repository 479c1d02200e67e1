set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
# default vs lean build of the small workload
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_small_default.json 2> gpurun_out/bench_small_default.err
B200MOBY_LIB=$PWD/moby_b200/libb200moby_lean.so timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_small_lean.json 2> gpurun_out/bench_small_lean.err
grep -h -o '"value": [0-9.]*, "unit": "env-steps/s", "n_gpus"' gpurun_out/bench_small_default.json gpurun_out/bench_small_lean.json
# stacks with the block finish kernel (bounded)
timeout 900 python bench.py --workload stacks --steps 3 --warmup 3 --preroll 3 > gpurun_out/bench_stacks.json 2> gpurun_out/bench_stacks.err
cat gpurun_out/bench_stacks.json | head -c 1500
# source-level ncu captures: one launch each of the advance kernel, a warp impact class and the n<=40 block class
timeout 900 ncu --set full --import-source on --clock-control none -k regex:advance_kernel -s 330 -c 1 -o gpurun_out/ncu_advance -f python bench.py --steps 2 --warmup 3 --preroll 300 --no-cpu-baseline > gpurun_out/ncu_advance.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:impact_warp_kernel -s 992 -c 1 -o gpurun_out/ncu_impact_warp -f python bench.py --steps 2 --warmup 3 --preroll 300 --no-cpu-baseline > gpurun_out/ncu_impact_warp.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:impact_block_kernel -s 991 -c 1 -o gpurun_out/ncu_impact_block -f python bench.py --steps 2 --warmup 3 --preroll 300 --no-cpu-baseline > gpurun_out/ncu_impact_block.log 2>&1
ls -la gpurun_out
