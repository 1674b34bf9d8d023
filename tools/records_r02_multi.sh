# Round-2 multi-GPU records: N = $1 GPUs.  Sims sweep (config 5), strong scaling, config 3 (train --max-gens 10)
N=$1
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
for s in 600 400 800 1600; do
  $TR bench.py --gpus $N --sims $s --steps 2 --warmup 2 > $O/r02_bench_${N}gpu_sims$s.json 2> $O/r02_bench_${N}gpu_sims$s.err
done
$TR bench.py --gpus $N --scaling strong --total-games 16384 --steps 2 --warmup 2 > $O/r02_bench_${N}gpu_strong16k.json 2> $O/r02_bench_${N}gpu_strong16k.err
$TR bench.py --gpus $N --scaling strong --total-games 131072 --steps 2 --warmup 2 > $O/r02_bench_${N}gpu_strong131k.json 2> $O/r02_bench_${N}gpu_strong131k.err
rm -rf /tmp/c4a0_train$N
$TR -m c4a0_b200.training --base-dir /tmp/c4a0_train$N --max-gens 10 --report $O/r02_config3_${N}gpu.jsonl > $O/r02_config3_${N}gpu.log 2>&1
tail -3 $O/r02_config3_${N}gpu.log
for f in $O/r02_bench_${N}gpu_*.json; do python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step']), 'wall ms', round(d['e2e']['wall_ms_per_step']))
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
done
