for mi in 2 3 4 6; do
  python bench.py --steps 1 --warmup 2 --no-ablation --no-cpu-baseline --max-inline $mi 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('max_inline $mi:', round(d['ms_per_step']), 'ms', round(d['ticks_per_step']), 'ticks', 'kstep', round(1e3*d['roofline']['avg_launch_ms'],1), 'nn', round(1e3*d['roofline']['nn_graph_avg_ms'],1), 'sims/launch', round(d['roofline']['sims_per_launch']))
"
done
