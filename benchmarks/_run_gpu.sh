mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_interp.py -m gpu -q --timeout 100 -k "allmasked" > gpurun_out/pytest_am.log 2>&1; echo "exit $?" >> gpurun_out/pytest_am.log
