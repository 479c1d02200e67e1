for v in reg noreg; do for n in 40 20; do
  if [ $v = noreg ]; then export B200MOBY_LIB=$PWD/moby_b200/libb200moby_noreg.so; else unset B200MOBY_LIB; fi
  timeout 120 python bench.py --workload lcp --lcp-n $n --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/lcpb_${v}_$n.json 2> gpurun_out/lcpb_${v}_$n.err
  tail -1 gpurun_out/lcpb_${v}_$n.err | cut -c1-200
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/lcpb_${v}_$n.json").read().strip().splitlines()[-1]); print("$v", $n, round(d["value"]), round(d["ms_per_step"],3), {k:d[k] for k in d if "pivot" in k})
except Exception as e: print("ERR", e)
P
done; done
