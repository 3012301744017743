for st in 2 3 4; do for pf in 0 1; do
  CRBE_LIB_PATH=/root/repo/scratch/libs/libcrbe_s${st}_p${pf}.so timeout 120 python bench.py --steps 60 --no-e2e --no-cpu-baseline 2>>gpurun_out/sweep.err | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('stages $st prefetch $pf | steps/s %.1f |' % d['value'], {k:round(v['ms_per_launch'],4) for k,v in d['kernels'].items()})
"
done; done
