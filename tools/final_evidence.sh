# Round-end evidence on one B200: ncu launch list of the default bench command (plain launches, feed off: ncu serialises
# kernels, so the hard-queue launch would only wait for class launches that cannot run beside it), final bench lines, smoke.
B200MOBY_FEED=0 B200MOBY_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 5800 -c 200 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 2 --warmup 3 --preroll 300 --no-cpu-baseline --no-secondary > gpurun_out/r02_ncu_final.log 2>&1
tail -2 gpurun_out/r02_ncu_final.log | cut -c1-200; wc -l gpurun_out/r02_launches_final.csv
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_final_ref.json 2>/dev/null
timeout 300 python bench.py --workload ur10 --steps 20 --warmup 5 > gpurun_out/r02_bench_ur10.json 2>/dev/null
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python - <<P
import json
d=json.loads(open("gpurun_out/r02_bench_ur10.json").read().strip().splitlines()[-1]); print("ur10", round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]))
d=json.loads(open("gpurun_out/r02_bench_final.json").read().strip().splitlines()[-1]); print("small", round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]), d["secondary"]["stabilization_off"]["value"], d["gpu_launches"], d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"]["traffic"])
d=json.loads(open("gpurun_out/r02_bench_final_ref.json").read().strip().splitlines()[-1]); print("ref", round(d["value"]))
P
