mkdir -p gpurun_out
timeout 120 python benchmarks/_km_prof.py > gpurun_out/kmplain.log 2>&1
timeout 200 python -m pytest tests/test_gpu_kmeans.py -m gpu -x -q --timeout 100 -k "lloyd or plusplus or reloc or gap" > gpurun_out/pytest_km.log 2>&1; echo "exit $?" >> gpurun_out/pytest_km.log
