for cfg in "1 1" "1 0" "0 0"; do set -- $cfg
  OCC_GS_WAVE=$1 OCC_GS_COEF=$2 python bench.py --workload occ --steps 2 --warmup 1 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
o=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('wave=$1 coef=$2 value %.1f'%o['value'], {k:round(v,1) for k,v in o['device_ms_per_step_by_kernel_group'].items()})"
done
