set -x
mkdir -p gpurun_out
timeout 200 python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu-baseline --no-graph > gpurun_out/b3.log 2>&1 || exit 1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 300 --csv --log-file gpurun_out/launches_steady.csv python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu-baseline --no-graph > gpurun_out/ncu.log 2>&1
echo "ncu rc=$?"
