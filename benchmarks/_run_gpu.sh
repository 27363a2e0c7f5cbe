mkdir -p gpurun_out
timeout 200 python benchmarks/_dbg_sweep.py > gpurun_out/dbg_sweep.log 2>&1
