set -x
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_multi.log 2>&1; echo "multi rc=$?"
tail -4 gpurun_out/pytest_multi.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-e2e 2>/dev/null | tail -1 > gpurun_out/bench_2gpu.json; cut -c1-260 gpurun_out/bench_2gpu.json
