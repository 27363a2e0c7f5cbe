#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for t in f32 f64; do
  python benchmarks/_ktc_one.py $t > /dev/null 2>&1 || exit 1
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:kmeans_assign_tc --launch-skip 2 --launch-count 1 -o gpurun_out/ktc_$t -f python benchmarks/_ktc_one.py $t > gpurun_out/ncu_ktc_$t.log 2>&1
  echo ncu $t rc=$?
  ncu -i gpurun_out/ktc_$t.ncu-rep --page raw --csv > gpurun_out/ktc_${t}_raw.csv 2>/dev/null
  ncu -i gpurun_out/ktc_$t.ncu-rep --page source --csv > gpurun_out/ktc_${t}_src.csv 2>/dev/null
done
