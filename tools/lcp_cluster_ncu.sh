#!/bin/bash
# ncu --set full of the 8-CTA-cluster Lemke kernel (n = 320, 296 problems); the report is read on the box, text comes back
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lcp_cluster_kernel --launch-skip 3 --launch-count 1 -o /tmp/lcpcl -f python bench.py --workload lcp --lcp-n 320 --envs 296 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/lcp_cluster_ncu.log 2>&1
ncu -i /tmp/lcpcl.ncu-rep --page raw --csv > /tmp/lcpcl_raw.csv 2>/dev/null
ncu -i /tmp/lcpcl.ncu-rep --page source --csv --print-source cuda,sass > /tmp/lcpcl_src.csv 2>/dev/null
python tools/ncu_hotspots.py /tmp/lcpcl_src.csv > gpurun_out/lcp_cluster_ncu_hotspots.txt 2>&1
python - <<PY
import csv
rows=list(csv.reader(open('/tmp/lcpcl_raw.csv')))
hdr,units,vals=rows[0],rows[1],rows[2]
want=('gpu__time_duration.sum','launch__grid_size','launch__cluster','launch__registers_per_thread','launch__occupancy','sm__throughput.avg.pct','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','smsp__inst_executed.sum','sm__warps_active.avg.pct','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_bytes.sum','smsp__average_warp','smsp__warp_issue_stalled','sm__inst_executed_pipe_fp64','l1tex__data_pipe_lsu_wavefronts.avg.pct','sm__pipe_fp64_cycles_active.avg.pct')
with open('gpurun_out/lcp_cluster_ncu_summary.txt','w') as f:
    for i,h in enumerate(hdr):
        if any(w in h for w in want): f.write(f"{h} [{units[i]}] = {vals[i]}\n")
PY
head -30 gpurun_out/lcp_cluster_ncu_hotspots.txt | cut -c1-200
