mkdir -p gpurun_out
timeout 120 python benchmarks/_km_prof2.py > gpurun_out/kmplain.log 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:"kmeans_assign_tile2" -s 2 -c 4 -o gpurun_out/prof_km4 python benchmarks/_km_prof2.py > gpurun_out/ncu_km4.log 2>&1
