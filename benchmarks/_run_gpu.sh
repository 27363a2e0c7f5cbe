mkdir -p gpurun_out; rm -f gpurun_out/parity_report.jsonl
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench7.json 2> gpurun_out/bench7.err; echo "exit $?" >> gpurun_out/bench7.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sci_fwd_kernel|sci_bwd_kernel|rbf_fwd_kernel|rbf_bwd_kernel" -s 4 -c 4 -o gpurun_out/prof_v5 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu5.log 2>&1; echo "exit $?" >> gpurun_out/ncu5.log
