mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_N8.json 2> gpurun_out/bench_N8.err; echo "exit $?" >> gpurun_out/bench_N8.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 --steps 3 --warmup 3 --no-e2e > gpurun_out/bench_N4.json 2> gpurun_out/bench_N4.err; echo "exit $?" >> gpurun_out/bench_N4.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_N2.json 2> gpurun_out/bench_N2.err; echo "exit $?" >> gpurun_out/bench_N2.err
