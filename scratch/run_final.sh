set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --steps 200 --warmup 5 > gpurun_out/bench_auto.json 2> gpurun_out/bench_auto.err; echo "bench rc=$?"
cat gpurun_out/bench_auto.json
timeout 300 python bench.py --steps 200 --warmup 5 --verify-always --no-e2e --no-cpu-baseline > gpurun_out/bench_always.json 2> gpurun_out/bench_always.err
cat gpurun_out/bench_always.json
timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/b3.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_auto.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu.log 2>&1
echo "ncu rc=$?"
