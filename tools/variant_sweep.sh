for v in 0 1 2 3; do
C4A0_NET_VARIANT=$v python bench.py --steps 2 --warmup 2 --no-ablation --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('variant $v:', round(d['ms_per_step'],1), 'ms', round(d['ticks_per_step']), 'ticks kstep', round(1e3*d['roofline']['avg_launch_ms'],1), 'nn', round(1e3*d['roofline']['nn_graph_avg_ms'],1), 'value', round(d['value']))
"
done
