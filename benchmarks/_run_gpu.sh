mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_final_labels.py -m gpu -x -q --timeout 100 > gpurun_out/pytest_final.log 2>&1; echo "exit $?" >> gpurun_out/pytest_final.log
