set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/debug_pendulum.py 5 10 > gpurun_out/dbg_pend.log 2>&1
B200MOBY_FUSED=1 timeout 300 python tools/debug_pendulum.py 5 10 > gpurun_out/dbg_pend_fused.log 2>&1
B200MOBY_CONCURRENT=0 timeout 300 python tools/debug_pendulum.py 5 10 > gpurun_out/dbg_pend_serial.log 2>&1
head -30 gpurun_out/dbg_pend*.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 300 python tools/impact_profile.py > gpurun_out/impact_profile.json 2> gpurun_out/impact_profile.err
B200MOBY_IMPACT_THREADS=32 timeout 300 python tools/impact_profile.py > gpurun_out/impact_profile_warp.json 2> gpurun_out/impact_profile_warp.err
ls -la gpurun_out
