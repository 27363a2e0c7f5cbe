mkdir -p gpurun_out
timeout 120 python benchmarks/_km_prof.py 2>&1 | head -3 > gpurun_out/km_hc_on.log
DIC_KMEANS_NO_HC=1 timeout 120 python benchmarks/_km_prof.py 2>&1 | head -3 > gpurun_out/km_hc_off.log
timeout 120 python benchmarks/_km_prof_small.py 2>&1 | head -3 >> gpurun_out/km_hc_on.log
DIC_KMEANS_NO_HC=1 timeout 120 python benchmarks/_km_prof_small.py 2>&1 | head -3 >> gpurun_out/km_hc_off.log
