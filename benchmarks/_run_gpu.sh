mkdir -p gpurun_out
timeout 300 python bench.py --workload c3 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "exit $?" >> gpurun_out/bench_c3.err
timeout 300 python bench.py --workload c1 --steps 3 --warmup 3 --no-e2e > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; echo "exit $?" >> gpurun_out/bench_c1.err
