#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python benchmarks/_ktc_one.py > /dev/null 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kmeans_assign_tc_kernel --launch-skip 2 --launch-count 1 -o gpurun_out/ktc_f32 -f python benchmarks/_ktc_one.py > gpurun_out/ncu_ktc.log 2>&1
echo ncu rc=$?
ncu -i gpurun_out/ktc_f32.ncu-rep --page raw --csv > gpurun_out/ktc_f32_raw.csv 2>/dev/null
ncu -i gpurun_out/ktc_f32.ncu-rep --page source --csv > gpurun_out/ktc_f32_src.csv 2>/dev/null
ls -la gpurun_out | tail -5
