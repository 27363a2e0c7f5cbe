mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kmeans.py tests/test_final_labels.py -m gpu -x -q --timeout 100 > gpurun_out/pytest_km.log 2>&1; echo "exit $?" >> gpurun_out/pytest_km.log
timeout 200 python benchmarks/_sweep_prof.py > gpurun_out/sweep_prof.log 2>&1
