run() {
  env $1 python bench.py --steps 1 --warmup 2 --no-ablation --no-cpu-baseline $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1 $2:', round(d['ms_per_step']), 'ms', round(d['ticks_per_step']), 'ticks', 'kstep', round(1e3*d['roofline']['avg_launch_ms'],1), 'nn', round(1e3*d['roofline']['nn_graph_avg_ms'],1), 'hit', round(d['eval_cache']['hit_rate_of_expansions'],3), 'spec rows', d['eval_cache']['speculative_rows'], 'nn rows', round(d['nn_evals_per_s']*d['ms_per_step']/1e3), 'GB', round(d['engine_device_gb'],1))
"
}
run "X=0" "--max-inline 3 --eval-cache-entries 67108864"
run "X=0" "--max-inline 3 --eval-cache-entries 268435456"
run "X=0" "--max-inline 3 --eval-cache-entries 536870912"
run "C4A0_SPEC_CHAIN=1" "--max-inline 3 --eval-cache-entries 536870912"
run "X=0" "--max-inline 3 --eval-cache-entries 1073741824"
