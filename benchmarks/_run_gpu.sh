mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_N8.json 2> gpurun_out/bench_N8.err; echo "exit $?" >> gpurun_out/bench_N8.err
