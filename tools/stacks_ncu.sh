#!/bin/bash
# ncu capture of the longest impact_block_kernel<256> launch of a small stacks run (sequential ladder: no cross-block waits under replay)
export B200MOBY_LADDER=0 B200MOBY_GRAPH=0
NE=${1:-24}; ST=${2:-5}
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:impact_block_kernel --csv --log-file /tmp/stk_launches.csv python tools/stacks_run.py $NE $ST > gpurun_out/stacks_ncu_pass1.log 2>&1
IDX=$(python - <<PY
import csv
rows=[r for r in csv.reader(open('/tmp/stk_launches.csv')) if len(r)>5]
hdr=None; best=(-1,0); k=0
for r in rows:
    if r[0]=='ID': hdr=r; continue
    if hdr is None: continue
    try: v=float(r[hdr.index('Metric Value')].replace(',',''))
    except: continue
    if v>best[1]: best=(k,v)
    k+=1
print(best[0])
import sys
sys.stderr.write(f"longest launch {best} of {k}\n")
PY
)
echo "capturing launch $IDX" > gpurun_out/stacks_ncu_pick.txt
timeout 600 ncu --section SourceCounters --section WarpStateStats --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --clock-control none --import-source on -k regex:impact_block_kernel --launch-skip $IDX --launch-count 1 -o /tmp/stk -f python tools/stacks_run.py $NE $ST > gpurun_out/stacks_ncu_pass2.log 2>&1
ncu -i /tmp/stk.ncu-rep --page raw --csv > gpurun_out/stacks_ncu_raw.csv 2>/dev/null
ncu -i /tmp/stk.ncu-rep --page source --csv --print-source cuda,sass > /tmp/stk_src.csv 2>/dev/null
python tools/ncu_hotspots.py /tmp/stk_src.csv > gpurun_out/stacks_ncu_hotspots.txt 2>&1
tail -3 gpurun_out/stacks_ncu_pass1.log; cat gpurun_out/stacks_ncu_pick.txt; head -60 gpurun_out/stacks_ncu_hotspots.txt
