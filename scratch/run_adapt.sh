set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_adapt.json 2> gpurun_out/bench_adapt.err; echo "bench rc=$?"
cat gpurun_out/bench_adapt.json
timeout 300 python bench.py --cells 1024 --no-cpu-baseline --no-e2e > gpurun_out/bench_adapt_1024.json 2> gpurun_out/bench_adapt_1024.err; cut -c1-1500 gpurun_out/bench_adapt_1024.json
timeout 300 python bench.py --cells 1024 --extrapolate-order 4 --no-cpu-baseline --no-e2e > gpurun_out/bench_o4_1024.json 2>/dev/null; cut -c1-1500 gpurun_out/bench_o4_1024.json
