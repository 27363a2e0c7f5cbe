mkdir -p gpurun_out
DIC_TC_PROFILE=1 timeout 100 python benchmarks/_pw_prof1.py > gpurun_out/pwprof.log 2>&1
