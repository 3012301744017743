set -x
mkdir -p gpurun_out
free -g | head -2
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 1000 --warmup 5 --no-cpu-baseline > gpurun_out/bench_o4.json 2> gpurun_out/bench_o4.err; echo "bench rc=$?"
cat gpurun_out/bench_o4.json
for q in 1 2 3; do timeout 300 python bench.py --steps 1000 --warmup 5 --extrapolate-order $q --no-cpu-baseline --no-e2e > gpurun_out/bench_o$q.json 2> gpurun_out/bench_o$q.err; cut -c1-220 gpurun_out/bench_o$q.json; done
for reg in P-T10 P-stiff; do timeout 300 python bench.py --steps 100 --warmup 5 --regime $reg --no-cpu-baseline --no-e2e > gpurun_out/bench_$reg.json 2> gpurun_out/bench_$reg.err; cut -c1-220 gpurun_out/bench_$reg.json; done
