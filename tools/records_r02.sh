# Round-2 records: default bench, sims sweep, config 4, config 3 on one GPU, compat path
set -x
O=gpurun_out
python bench.py --steps 2 --warmup 3 > $O/r02_bench_default.json 2> $O/r02_bench_default.err
for s in 400 800 1600; do
  python bench.py --sims $s --steps 2 --warmup 2 --no-ablation --no-cpu-baseline > $O/r02_bench_sims$s.json 2> $O/r02_bench_sims$s.err
done
python tools/compat_probe.py 1000 600 2000 > $O/r02_compat_path.jsonl 2> $O/r02_compat_path.err
rm -rf /tmp/c4a0_train && python -m c4a0_b200.training --base-dir /tmp/c4a0_train --max-gens 10 --report $O/r02_config3_1gpu.jsonl > $O/r02_config3_1gpu.log 2>&1
python bench.py --preset config4 --steps 1 --warmup 1 --no-ablation --no-cpu-baseline > $O/r02_bench_config4.json 2> $O/r02_bench_config4.err
tail -2 $O/r02_config3_1gpu.log; cat $O/r02_compat_path.jsonl
