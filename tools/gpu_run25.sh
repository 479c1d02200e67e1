cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/y_bench_small_2gpu.json 2> gpurun_out/y_bench_small_2gpu.err
wc -l gpurun_out/y_bench_small_2gpu.json; head -c 200 gpurun_out/y_bench_small_2gpu.json
