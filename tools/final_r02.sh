# Final round-2 validation on one GPU: whole GPU suite, smoke, default bench line (ablations + CPU baseline), launch list
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/r02_final_tests.log; cat $O/r02_final_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 > $O/r02_final_bench.json 2> $O/r02_final_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_final_bench.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],1), 'kstep', d['roofline']['avg_launch_ms'], 'nn', d['nn_roofline']['avg_launch_ms'], 'cpu', d['cpu_baseline']['value'])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --kill 1 --csv --log-file $O/r02_launches_bench_python_loop.csv \
  python bench.py --steps 1 --warmup 0 --host-loop python --no-ablation --no-cpu-baseline > $O/ncu_launches.log 2>&1
tail -2 $O/ncu_launches.log; wc -l $O/r02_launches_bench_python_loop.csv
