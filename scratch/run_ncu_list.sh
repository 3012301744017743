set -x
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 250 --csv --log-file gpurun_out/launches_steady.csv python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu-baseline --no-graph > gpurun_out/ncu.log 2>&1
echo "ncu rc=$?"
