# A/B of programmatic dependent launch in the native loop (C4A0_PDL=0 -> ordinary launches), parity tests with it on
timeout 300 python -m pytest tests/test_gpu_net.py tests/test_gpu_bigconfig.py tests/test_gpu_golden.py tests/test_gpu_next.py -x -q 2>&1 | tail -3
for v in 1 0 1 0; do
C4A0_PDL=$v python bench.py --steps 2 --warmup 2 --no-ablation --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('PDL=$v', round(d['ms_per_step'],1), 'ms', round(d['ticks_per_step']), 'ticks kstep', round(1e3*d['roofline']['avg_launch_ms'],1), 'nn', round(1e3*d['roofline']['nn_graph_avg_ms'],1), 'e2e', round(d['e2e']['value']), 'value', round(d['value']))
"
done
