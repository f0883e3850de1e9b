python -m pytest tests -m gpu -x -q 2>&1 | tail -3
IFS=";" read -ra ARR <<< "${CFGS:-4 1 24 2 3 1 ;4 1 16 2 3 1 ;4 1 24 2 3 1 1;2 1 24 2 3 1 ;8 1 24 2 3 1 ;4 1 24 2 2 1 }"; for cfg in "${ARR[@]}"; do
  set -- $cfg
  RRTQX_GRID_OCCUPANCY=$1 RRTQX_GRID_ASPECT=$2 RRTQX_FUSED_NW=$3 RRTQX_FUSED_QN=$4 RRTQX_QSORT_S=$5 RRTQX_QSORT_F=$6 RRTQX_FUSED_PAIR=$7 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-sweep 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('occ=$1 asp=$2 NW=$3 QN=$4 S=$5 F=$6 PAIR=$7', 'ms', round(d['ms_per_step'],3), 'fill', round(d['roofline']['kernel_ms']['range_fill'],3), 'sort', round(d['roofline']['kernel_ms']['range_sort'],3), 'frac', round(d['roofline']['frac'],3))
    elif 'rror' in l: print(l.strip()[:200])
"
done
