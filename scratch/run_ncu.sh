set -x
mkdir -p gpurun_out
timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-graph > gpurun_out/b3.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:t_pv|t_st|k_xrp|t_init|k_s" -s 10 -c 6 -f -o gpurun_out/r01_idx16 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-graph > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep
