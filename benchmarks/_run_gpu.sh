mkdir -p gpurun_out; rm -f gpurun_out/parity_report.jsonl
timeout 600 python -m pytest tests -m gpu -x -q --timeout 100 > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/pytest_gpu.log
