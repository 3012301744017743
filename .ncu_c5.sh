set -x
python bench.py --config5 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r02d_c5.json 2> gpurun_out/r02d_c5.err
ncu --set full --import-source on --clock-control none -k regex:k_update_system_rows -s 2 -c 1 -o gpurun_out/r02d_ncu_c5 python bench.py --config5 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r02d_ncu_c5.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"t_init_be|t_pv|t_st|k_xrp|k_s|k_extrapolate" -s 300 -c 12 -o gpurun_out/r02d_ncu_step python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-strong --no-e2e --chunk 1 --no-graph > gpurun_out/r02d_ncu_step.log 2>&1
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "reassembly or chunk or last_iteration or step_residual or time_varying" 2>&1 | tail -5
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-strong --no-e2e > gpurun_out/r02d_bench_n1.json 2> gpurun_out/r02d_bench_n1.err
ls -la gpurun_out/
