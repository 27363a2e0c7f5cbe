mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_final_labels.py -m gpu -x -q --timeout 100 > gpurun_out/pytest_final.log 2>&1; echo "exit $?" >> gpurun_out/pytest_final.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_N2.json 2> gpurun_out/bench_N2.err; echo "exit $?" >> gpurun_out/bench_N2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench_ref_N2.json 2> gpurun_out/bench_ref_N2.err; echo "exit $?" >> gpurun_out/bench_ref_N2.err
