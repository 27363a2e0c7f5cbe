mkdir -p gpurun_out
DIC_KMEANS_T2_DB=1 timeout 300 python -m pytest tests/test_gpu_kmeans.py -m gpu -x -q --timeout 100 -k "lloyd or dispatch or config4" > gpurun_out/pytest_km.log 2>&1; echo "exit $?" >> gpurun_out/pytest_km.log
timeout 300 python benchmarks/_km_pass.py > gpurun_out/km_pass_sb.json 2> gpurun_out/km_pass_sb.err
DIC_KMEANS_T2_DB=1 timeout 300 python benchmarks/_km_pass.py > gpurun_out/km_pass_db.json 2> gpurun_out/km_pass_db.err
