set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/v_bench_small_2gpu.json 2> gpurun_out/v_bench_small_2gpu.err
tail -c 600 gpurun_out/v_bench_small_2gpu.json; tail -3 gpurun_out/v_bench_small_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --impl reference --steps 5 --warmup 3 > gpurun_out/v_bench_ref_2gpu.json 2> gpurun_out/v_bench_ref_2gpu.err
tail -c 300 gpurun_out/v_bench_ref_2gpu.json
