# in-kernel simulation budget while speculating = C4A0_INLINE_SPEC_MULT x max_inline_sims (default 4 x 3)
for m in ${MULTS:-1 2 3 4 6}; do
C4A0_INLINE_SPEC_MULT=$m python bench.py --steps 2 --warmup 2 --no-ablation --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('mult $m:', round(d['ms_per_step'],1), 'ms', round(d['ticks_per_step']), 'ticks kstep', round(1e3*d['roofline']['avg_launch_ms'],1), 'nn', round(1e3*d['roofline']['nn_graph_avg_ms'],1), 'hit', round(d['eval_cache']['hit_rate_of_expansions'],3), 'value', round(d['value']))
"
done
