mkdir -p gpurun_out
timeout 120 python benchmarks/_km_prof2.py > gpurun_out/kmplain.log 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:"kmeans_assign_rw" -s 2 -c 1 -o gpurun_out/prof_km3 python benchmarks/_km_prof2.py > gpurun_out/ncu_km3.log 2>&1
