set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 300 python bench.py > gpurun_out/z_bench_small_default.json 2> gpurun_out/z_bench_small_default.err
head -c 300 gpurun_out/z_bench_small_default.json
