set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/bench_idx16.json 2> gpurun_out/bench_idx16.err; echo "bench rc=$?"
cat gpurun_out/bench_idx16.json
timeout 300 python bench.py --steps 200 --warmup 5 --index32 --no-cpu-baseline --no-e2e > gpurun_out/bench_idx32.json 2> gpurun_out/bench_idx32.err
cat gpurun_out/bench_idx32.json
