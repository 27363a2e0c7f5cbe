mkdir -p gpurun_out
DIC_TC_PROFILE=1 timeout 100 python benchmarks/_pw_prof1.py > gpurun_out/pwprof.log 2>&1
timeout 60 python -m pytest tests/test_gpu_kmeans.py -m gpu -x -q --timeout 50 -k "tensor or silhouette" > gpurun_out/pytest_tc.log 2>&1; echo "exit $?" >> gpurun_out/pytest_tc.log
timeout 100 python benchmarks/_pw_prof.py > gpurun_out/pwplain.log 2>&1; echo "exit $?" >> gpurun_out/pwplain.log
