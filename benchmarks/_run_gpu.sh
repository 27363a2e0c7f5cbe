mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kmeans.py -m gpu -q --timeout 100 -k "dispatch" > gpurun_out/pytest_disp.log 2>&1; echo "exit $?" >> gpurun_out/pytest_disp.log
