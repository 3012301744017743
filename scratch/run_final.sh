set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
cat gpurun_out/bench_final.json
timeout 900 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"
cat gpurun_out/bench_reference.json
timeout 300 python bench.py --steps 30 --warmup 30 --no-e2e --no-cpu-baseline > gpurun_out/b3.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 30 --warmup 30 --no-e2e --no-cpu-baseline > gpurun_out/ncu.log 2>&1
echo "ncu rc=$?"
