mkdir -p gpurun_out; rm -f gpurun_out/parity_report.jsonl
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit $?" >> gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench12.json 2> gpurun_out/bench12.err; echo "exit $?" >> gpurun_out/bench12.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref3.json 2> gpurun_out/bench_ref3.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches3.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --encounters 131072 > gpurun_out/ncu_launch.log 2>&1; echo "exit $?" >> gpurun_out/ncu_launch.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sci_fwd_kernel|sci_bwd_kernel|rbf_fwd2_kernel|rbf_bwd_kernel|cci_fwd_warp|cci_bwd_warp" -s 6 -c 6 -o gpurun_out/prof_v8 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu8.log 2>&1; echo "exit $?" >> gpurun_out/ncu8.log
timeout 900 python benchmarks/bench_kselect.py --sweep-n 20000 --cpu-n 500 --full-refs 2 --full-ninit 2 > gpurun_out/kselect12.json 2> gpurun_out/kselect12.err; echo "exit $?" >> gpurun_out/kselect12.err
