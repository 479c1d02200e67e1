set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/final_bench_small.json 2> gpurun_out/final_bench_small.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/final_bench_small_reference.json 2> gpurun_out/final_bench_small_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches_small.csv python bench.py --steps 2 --warmup 1 --preroll 20 --no-cpu-baseline > gpurun_out/final_ncu_launches.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:impact_thread_kernel -s 1604 -c 1 -o gpurun_out/ncu_final_impact_thread -f python bench.py --steps 2 --warmup 3 --preroll 300 --no-cpu-baseline > gpurun_out/ncu_final_impact_thread.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:impact_block_kernel -s 640 -c 1 -o gpurun_out/ncu_final_hard_queue -f python bench.py --steps 2 --warmup 3 --preroll 300 --no-cpu-baseline > gpurun_out/ncu_final_hard_queue.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:advance_thread_kernel -s 330 -c 1 -o gpurun_out/ncu_final_advance_thread -f python bench.py --steps 2 --warmup 3 --preroll 300 --no-cpu-baseline > gpurun_out/ncu_final_advance_thread.log 2>&1
timeout 900 python bench.py --workload stacks --envs-per-gpu 512 --steps 3 --warmup 3 --preroll 2 > gpurun_out/final_bench_stacks512.json 2> gpurun_out/final_bench_stacks512.err
timeout 300 python bench.py --workload lcp --steps 10 --warmup 3 > gpurun_out/final_bench_lcp32.json 2> gpurun_out/final_bench_lcp32.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
ls -la gpurun_out | tail -15
