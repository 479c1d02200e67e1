for v in reg noreg; do
  if [ $v = noreg ]; then export B200MOBY_LIB=$PWD/moby_b200/libb200moby_noreg.so; else unset B200MOBY_LIB; fi
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:lcp_warp_kernel --launch-skip 4 --launch-count 1 -o /tmp/lcp_$v -f python bench.py --workload lcp --lcp-n 40 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/lcp_ncu_$v.log 2>&1
  ncu -i /tmp/lcp_$v.ncu-rep --page raw --csv > gpurun_out/lcp_ncu_${v}_raw.csv 2>/dev/null
  ncu -i /tmp/lcp_$v.ncu-rep --page source --csv --print-source cuda,sass > /tmp/src_$v.csv 2>/dev/null
  python tools/ncu_hotspots.py /tmp/src_$v.csv > gpurun_out/lcp_ncu_${v}_hotspots.txt 2>&1
done
