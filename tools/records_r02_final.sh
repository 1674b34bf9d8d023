# records of the final round-2 build on one GPU: sims sweep, config 4, sanitizer
O=gpurun_out
for s in 400 800 1600; do
  python bench.py --sims $s --steps 2 --warmup 2 --no-ablation --no-cpu-baseline > $O/r02_final_bench_sims$s.json 2> $O/r02_final_bench_sims$s.err
done
python bench.py --preset config4 --steps 1 --warmup 1 --no-ablation --no-cpu-baseline > $O/r02_final_bench_config4.json 2> $O/r02_final_bench_config4.err
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_target.py --native > $O/r02_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/r02_sanitizer_memcheck.log
tail -4 $O/r02_sanitizer_memcheck.log
for f in sims400 sims800 sims1600 config4; do python - $O/r02_final_bench_$f.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],1), 'sims/s', round(d['sims_per_s']/1e6,1), 'kstep', round(1e3*d['roofline']['avg_launch_ms'],1), 'frac', round(d['roofline']['frac'],4), 'nn', round(1e3*d['nn_roofline']['avg_launch_ms'],1), 'nn TF', round(d['nn_roofline']['achieved']))
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
done
