mkdir -p gpurun_out; rm -f gpurun_out/parity_report.jsonl
timeout 300 python -m pytest tests/test_gpu_interp.py -m gpu -q --timeout 100 > gpurun_out/pytest_interp.log 2>&1; echo "exit $?" >> gpurun_out/pytest_interp.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench18.json 2> gpurun_out/bench18.err; echo "exit $?" >> gpurun_out/bench18.err
