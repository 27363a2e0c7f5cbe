#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kmeans.py -m gpu -q -x 2>&1 | tail -5
timeout 600 python benchmarks/_km_small.py > gpurun_out/km_small.json 2> gpurun_out/km_small.err; echo rc=$?
