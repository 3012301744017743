set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_lean.json 2> gpurun_out/bench_lean.err; echo "bench rc=$?"
cat gpurun_out/bench_lean.json
timeout 300 python bench.py --steps 30 --warmup 30 --no-e2e --no-cpu-baseline --no-graph > gpurun_out/b3.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:t_pv|t_st|k_xrp|t_init|k_s|k_extrapolate" -s 900 -c 12 -f -o gpurun_out/r01_final python bench.py --steps 30 --warmup 30 --no-e2e --no-cpu-baseline --no-graph > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"
