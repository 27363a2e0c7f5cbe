mkdir -p gpurun_out; rm -f gpurun_out/parity_report.jsonl
timeout 120 python -m pytest tests/test_gpu_kmeans.py -m gpu -x -q --timeout 100 -k "silhouette or tensor" > gpurun_out/pytest_sil.log 2>&1; echo "exit $?" >> gpurun_out/pytest_sil.log
timeout 200 python -m pytest tests/test_gpu_kmeans.py tests/test_gpu_interp.py -m gpu -x -q --timeout 100 >> gpurun_out/pytest_sil.log 2>&1; echo "exit $?" >> gpurun_out/pytest_sil.log
timeout 300 python bench.py --workload c5 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "exit $?" >> gpurun_out/bench_c5.err
