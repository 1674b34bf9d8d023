for cfg in "1024 8192" "2048 8192" "4096 8192" "4096 16384" "8192 16384" "16384 16384"; do
  set -- $cfg
  C4A0_SPEC_THR=$1 python bench.py --steps 1 --warmup 2 --no-ablation --no-cpu-baseline --spec-rows $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('thr $1 spec $2:', round(d['ms_per_step']), 'ms', round(d['ticks_per_step']), 'ticks', 'kstep', round(1e3*d['roofline']['avg_launch_ms'],1), 'nn', round(1e3*d['roofline']['nn_graph_avg_ms'],1), 'spec rows', d['eval_cache']['speculative_rows'], 'hit', round(d['eval_cache']['hit_rate_of_expansions'],3))
"
done
