set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 300 gpurun_out/bench_default.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 4800 -c 100 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:impact -s 3600 -c 6 -o /tmp/impact_full -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ncu -i /tmp/impact_full.ncu-rep --page raw --csv > gpurun_out/impact_raw.csv 2>/dev/null
ncu -i /tmp/impact_full.ncu-rep --page source --csv > gpurun_out/impact_source.csv 2>gpurun_out/impact_source.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:advance -s 600 -c 1 -o /tmp/advance_full -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1
ncu -i /tmp/advance_full.ncu-rep --page raw --csv > gpurun_out/advance_raw.csv 2>/dev/null
ncu -i /tmp/advance_full.ncu-rep --page source --csv > gpurun_out/advance_source.csv 2>gpurun_out/advance_source.err
gzip -f gpurun_out/*_source.csv
du -sh gpurun_out; ls -la gpurun_out
