mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q --timeout 600 > gpurun_out/pytest_multi.log 2>&1; echo "exit $?" >> gpurun_out/pytest_multi.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tests/dist_gpu_check.py > gpurun_out/dist_check.json 2> gpurun_out/dist_check.err; echo "exit $?" >> gpurun_out/dist_check.err
