#!/bin/bash
# Phase probes of the tcgen05 Lloyd pass: benchmark builds with phases removed / variants chosen at compile time.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for s in ${KTC_VARIANTS:-"-DDIC_KTC_SKIP=0"}; do
  touch deep_interpolation_clustering_b200/csrc/kmeans_tc.cu
  NVCC_EXTRA="$s" python -m deep_interpolation_clustering_b200.build > /dev/null
  echo "$s $(timeout 300 python benchmarks/_ktc_probe.py 2>&1 | tail -1)"
done | tee gpurun_out/ktc_probe.txt
